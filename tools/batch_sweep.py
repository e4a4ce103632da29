"""GPU: step time and images/s against the batch size (wave quantisation on 148 SMs = 74 CTA pairs):
    python tools/batch_sweep.py resnet50 bf16 63 64 96 97 128 185 192 193 194 256"""
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, dtype = sys.argv[1], sys.argv[2]
sizes = [int(a) for a in sys.argv[3:]]
wdir = weights.cached_weights_dir(arch, 0)
ms = {}
for B in sizes:
    m = engine.ResNet(arch, wdir, dtype=dtype, max_batch=B)
    x = weights.synthetic_images(B).cuda()
    lg, t1 = m.forward(x)
    for _ in range(5):
        m.forward(x, lg, t1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(6):
        time.sleep(0.25)
        m.forward(x, lg, t1)
        e0.record()
        for _ in range(10):
            m.forward(x, lg, t1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10)
    t = statistics.median(ts)
    print(f"{arch} {dtype} B={B}: {t:.4f} ms per step = {B / t:.2f} k images/s, {t / B * 1e3:.3f} us per image, launches {m.launches_per_forward(B)}", flush=True)
    m.close()
