"""GPU: A/B of a run-time kernel switch. Prints, for the current environment, the per-kind event profile of one ResNet-50
BF16 B=256 step, the CUDA-graph replay time of 30 steps and a checksum of the logits (bit-identity check between runs)."""
import hashlib
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype="bf16", max_batch=B)
x = weights.synthetic_images(B).cuda()
logits, top1 = m.forward(x)
torch.cuda.synchronize()
print("logits sha", hashlib.sha256(logits.cpu().numpy().tobytes()).hexdigest()[:16], "top1 sum", int(top1.sum()))
for _ in range(5):
    m.forward(x, logits, top1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(30):
        m.forward(x, logits, top1)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 30)
print(f"graph replay {best:.4f} ms/step  {B / best:.1f} k img/s")
prof = m.profile(x, iters=3)
agg = {}
for p in prof:
    agg[p["kind"]] = agg.get(p["kind"], 0.0) + p["ms"]
print("first conv launches us:", [round(p["ms"] * 1e3, 1) for p in prof if p["kind"] == "conv_igemm"][:6])
print("profile us:", {k: round(v * 1e3, 1) for k, v in agg.items()}, "sum", round(sum(agg.values()) * 1e3, 1))
