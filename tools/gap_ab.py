"""GPU: does the 20-step bench figure depend on what ran just before it? Same model; each trial = `pre` untimed steps
immediately followed by 20 timed steps, after `gap` seconds of idle before the trial / between pre and timed steps."""
import statistics
import sys
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


def trial(pre, gap_mid):
    for _ in range(pre):
        m.forward(x, lg, t1)
    torch.cuda.synchronize()
    if gap_mid:
        time.sleep(gap_mid)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.forward(x, lg, t1)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


for pre, gap_mid in ((5, 0.0), (5, 0.3), (200, 0.0), (200, 0.3), (200, 1.0), (5, 0.0)):
    ts = []
    for _ in range(5):
        time.sleep(0.5)
        ts.append(trial(pre, gap_mid))
    print(f"{arch} B={B} pre {pre} steps, idle {gap_mid} s before the timed 20: ms/step min {min(ts):.4f} med {statistics.median(ts):.4f} max {max(ts):.4f}", flush=True)
