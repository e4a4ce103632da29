"""GPU: interleaved A/B of plan-time environment switches on one box.

    python tools/ab.py resnet50 256 "" "RNB_NO_SPLIT=1" "RNB_C3N1=1 RNB_NO_SPLIT=0"

One model per variant (the switches are read when a model is created / planned), all alive at once. Per variant:
logits checksum (bit-identity across variants), BURST step time (10-step CUDA-graph bursts separated by idle gaps, so
the board's power controller does not set the clock: this is the critical path) and SUSTAINED step time (1.5 s of
back-to-back replays: this is what the 1000 W cap allows, i.e. the energy of a step). Variants are interleaved
burst by burst so that box-to-box and minute-to-minute drift cancels."""
import hashlib
import os
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, B = sys.argv[1], int(sys.argv[2])
variants = sys.argv[3:] or [""]
dtype = os.environ.get("AB_DTYPE", "bf16")
x = weights.synthetic_images(B).cuda()
models, outs = [], []
for v in variants:
    kv = dict(p.split("=", 1) for p in v.split())
    old = {k: os.environ.get(k) for k in kv}
    os.environ.update(kv)
    m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype=dtype, max_batch=B)
    lg, t1 = m.forward(x)
    for _ in range(5):
        m.forward(x, lg, t1)
    torch.cuda.synchronize()
    for k, o in old.items():
        if o is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = o
    models.append(m)
    outs.append((lg, t1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
burst = [[] for _ in variants]
for rep in range(8):
    for i, m in enumerate(models):
        time.sleep(0.25)
        lg, t1 = outs[i]
        m.forward(x, lg, t1)
        e0.record()
        for _ in range(10):
            m.forward(x, lg, t1)
        e1.record()
        torch.cuda.synchronize()
        burst[i].append(e0.elapsed_time(e1) / 10)
for i, m in enumerate(models):
    lg, t1 = outs[i]
    time.sleep(1.0)
    t0 = time.perf_counter()
    n = 0
    e0.record()
    while time.perf_counter() - t0 < 1.5:
        for _ in range(25):
            m.forward(x, lg, t1)
        n += 25
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sus = e0.elapsed_time(e1) / n
    sha = hashlib.sha256(lg.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"{arch} B={B} [{variants[i] or 'default'}] sha {sha} burst min {min(burst[i]):.4f} med "
          f"{statistics.median(burst[i]):.4f} ms | sustained {sus:.4f} ms | launches {m.launches_per_forward(B)}",
          flush=True)
