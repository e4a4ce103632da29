"""GPU: energy of one step. Replays the CUDA graph of `arch batch` for ~`seconds` and reads NVML's total-energy counter and the
SM clock around it: joules per step, average board power, clock under load, throttle reasons. Under a power cap the step
time follows the energy of a step, so this — not the critical path — is what an A/B of two kernel variants has to compare
(run it once per environment, e.g. RNB_FUSE=0 against the default)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pynvml  # noqa: E402
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
dtype = sys.argv[4] if len(sys.argv) > 4 else "bf16"
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype=dtype, max_batch=B)
x = weights.synthetic_images(B).cuda()
logits, top1 = m.forward(x)
for _ in range(20):
    m.forward(x, logits, top1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
clocks, reasons, steps = [], 0, 0
j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)  # mJ
t0 = time.time()
e0.record()
while time.time() - t0 < seconds:
    for _ in range(50):
        m.forward(x, logits, top1)
    steps += 50
    torch.cuda.synchronize()
    clocks.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    reasons |= pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
e1.record()
torch.cuda.synchronize()
j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
ms = e0.elapsed_time(e1)
joules = (j1 - j0) / 1e3
print(f"{arch} {dtype} B={B}: {steps} steps, {ms / steps:.4f} ms/step, {joules / steps:.4f} J/step, {joules / (ms / 1e3):.0f} W average "
      f"(enforced limit {limit_w:.0f} W), {joules / steps / B * 1e3:.3f} mJ/image, SM clock {min(clocks)}-{max(clocks)} MHz, "
      f"reasons 0x{reasons:x}")
