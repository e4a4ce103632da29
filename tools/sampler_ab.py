"""GPU: how much does clock sampling perturb a 20-step timed region? Modes: no sampler, `nvidia-smi -lms 50`,
`nvidia-smi -lms 200`, an in-process pynvml thread every 50 ms. Same model, interleaved bursts of 20 steps."""
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, B = sys.argv[1], int(sys.argv[2])
m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0), dtype="bf16", max_batch=B)
x = weights.synthetic_images(B).cuda()
lg, t1 = m.forward(x)
for _ in range(5):
    m.forward(x, lg, t1)
torch.cuda.synchronize()
FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
          "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


def burst():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        m.forward(x, lg, t1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


class Nvml(threading.Thread):
    def __init__(self, period):
        super().__init__(daemon=True)
        import pynvml
        pynvml.nvmlInit()
        self.nv, self.h, self.period, self.stop, self.n = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0), period, False, 0

    def run(self):
        while not self.stop:
            self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
            self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            self.nv.nvmlDeviceGetPowerUsage(self.h)
            self.n += 1
            time.sleep(self.period)


res = {}
for mode in ("none", "smi50", "smi200", "nvml50", "none2"):
    proc = th = None
    if mode.startswith("smi"):
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={FIELDS}", "--format=csv,noheader,nounits", "-lms", mode[3:],
                                 "-i", "0"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(1.5)  # start-up over
    if mode == "nvml50":
        th = Nvml(0.05)
        th.start()
        time.sleep(0.2)
    ts = []
    for _ in range(8):
        time.sleep(0.2)
        ts.append(burst())
    if proc:
        proc.terminate()
        proc.wait()
    if th:
        th.stop = True
        th.join()
    res[mode] = ts
    print(f"{arch} B={B} sampler {mode}: ms/step min {min(ts):.4f} med {statistics.median(ts):.4f} max {max(ts):.4f}", flush=True)
