"""GPU: energy of each tensor-core conv launch of one step (`arch batch`). Every launch is replayed back to back for
~`seconds` (rnb_model_repeat_launch) between two reads of NVML's total-energy counter: µs, joules, average watts, pJ per
algorithmic FLOP and per algorithmic byte for each launch. A replayed launch finds its operands in whatever state its own
previous run left L2 in, so the numbers rank the launches (where does a 1000 W cap bite) rather than add up exactly to the
energy of a step (tools/energy_ab.py measures that)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import pynvml  # noqa: E402
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
seconds = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype="bf16", max_batch=B)
x = weights.synthetic_images(B).cuda()
m.forward(x)
prof = [p for p in m.profile(x, iters=3) if p["kind"] == "conv_igemm"]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
idle_w = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
print(f"{arch} B={B}: {len(prof)} conv launches, {seconds} s each; board power before the run {idle_w:.0f} W")
print("  # |    us | in-step us |  J/launch |     W |  GFLOP |     MB | pJ/FLOP | pJ/byte | MHz")
tot_j = tot_ms = 0.0
for i, p in enumerate(prof):
    rep = max(50, int(seconds / (p["ms"] * 1e-3)))
    m.repeat_launch(B, i, 50)
    torch.cuda.synchronize()
    j0 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    e0.record()
    m.repeat_launch(B, i, rep)
    e1.record()
    torch.cuda.synchronize()
    j1 = pynvml.nvmlDeviceGetTotalEnergyConsumption(h)
    mhz = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
    ms = e0.elapsed_time(e1) / rep
    j = (j1 - j0) / 1e3 / rep
    tot_j += j
    tot_ms += ms
    print(f"{i:3d} | {ms * 1e3:5.1f} | {p['ms'] * 1e3:10.1f} | {j:9.5f} | {j / (ms * 1e-3):5.0f} | {p['flops'] / 1e9:6.1f} | "
          f"{p['bytes'] / 1e6:6.1f} | {j / p['flops'] * 1e12:7.3f} | {j / p['bytes'] * 1e12:7.1f} | {mhz}")
    time.sleep(0.05)
print(f"sum over conv launches: {tot_ms:.3f} ms, {tot_j:.3f} J ({tot_j / tot_ms * 1e3:.0f} W average)")
