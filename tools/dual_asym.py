"""GPU experiment: asymmetric two-lane splits of one batch (main lane + small helper lane on a second stream)."""
import statistics
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os  # noqa: E402

os.environ["RNB_LANES"] = "1"
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, B = sys.argv[1], int(sys.argv[2])
splits = [tuple(int(v) for v in a.split("+")) for a in sys.argv[3:]]
wdir = weights.cached_weights_dir(arch, 0, True)


def setup(parts):
    ms = [engine.ResNet(arch, wdir, dtype="bf16", max_batch=n) for n in parts]
    xs = [weights.synthetic_images(n, seed=10 + i).cuda() for i, n in enumerate(parts)]
    ss = [torch.cuda.Stream() for _ in parts]
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


cfgs = {p: setup(p) for p in [(B,)] + splits}
res = {p: [] for p in cfgs}
for rep in range(6):
    for p, cfg in cfgs.items():
        time.sleep(0.25)
        burst(cfg, 2)
        res[p].append(burst(cfg))
for p in cfgs:
    print(f"{arch} B={B} as {'+'.join(map(str, p))}: burst med {statistics.median(res[p]):.4f} ms per step (min {min(res[p]):.4f})", flush=True)
