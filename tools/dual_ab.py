"""GPU experiment: do two half-batch replicas on two streams fill each other's wave tails?

    python tools/dual_ab.py resnet50 256            (RNB_NO_PDL=1 in the environment for the no-PDL variant)

Times (a) one model forwarding B images per step, (b) two models forwarding B/2 images each per step on two CUDA
streams (both graphs in flight at once), (c) four models with B/4. Bursts of 10 steps separated by idle gaps."""
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, B = sys.argv[1], int(sys.argv[2])
wdir = weights.cached_weights_dir(arch, 0, True)


def setup(parts):
    n = B // parts
    ms = [engine.ResNet(arch, wdir, dtype="bf16", max_batch=n) for _ in range(parts)]
    xs = [weights.synthetic_images(n, seed=10 + i).cuda() for i in range(parts)]
    ss = [torch.cuda.Stream() for _ in range(parts)]
    outs = []
    for m, x, s in zip(ms, xs, ss):
        with torch.cuda.stream(s):
            lg, t1 = m.forward(x)
            for _ in range(3):
                m.forward(x, lg, t1)
        outs.append((lg, t1))
    torch.cuda.synchronize()
    return ms, xs, ss, outs


def burst(cfg, steps=10):
    ms, xs, ss, outs = cfg
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for m, x, s, (lg, t1) in zip(ms, xs, ss, outs):
            with torch.cuda.stream(s):
                m.forward(x, lg, t1)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


cfgs = {p: setup(p) for p in (1, 2, 4)}
res = {p: [] for p in cfgs}
for rep in range(8):
    for p, cfg in cfgs.items():
        time.sleep(0.25)
        burst(cfg, 2)
        res[p].append(burst(cfg))
for p in cfgs:
    print(f"{arch} B={B} as {p} x {B // p} on {p} stream(s): burst min {min(res[p]):.4f} med {statistics.median(res[p]):.4f} ms per step "
          f"(host wall clock around 10 steps)", flush=True)
