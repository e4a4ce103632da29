"""GPU: A/B of the shared-memory split of bneck_c3n1s_kernel (RNB_C3N1S_RINGS, see bneck_c3n1.cuh). For the current
environment prints the logits checksum (bit-identity between variants), the CUDA-graph replay time of a ResNet-50 BF16
B=256 step (min and median of 8 x 40 steps) and the event-timed launches of the layer3 fused kernel (conv3 256 -> 1024 +
conv1' 1024 -> 256 at 14 x 14: 52.6 GFLOP per launch)."""
import hashlib
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

B = 256
m = engine.ResNet("resnet50", weights.cached_weights_dir("resnet50", 0, True), dtype="bf16", max_batch=B)
x = weights.synthetic_images(B).cuda()
logits, top1 = m.forward(x)
torch.cuda.synchronize()
sha = hashlib.sha256(logits.cpu().numpy().tobytes()).hexdigest()[:16]
for _ in range(10):
    m.forward(x, logits, top1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(8):
    e0.record()
    for _ in range(40):
        m.forward(x, logits, top1)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 40)
prof = m.profile(x, iters=5)
want = 2.0 * B * 196 * 256 * 1024 * 2
l3 = [round(p["ms"] * 1e3, 1) for p in prof if p["kind"] == "conv_igemm" and abs(p["flops"] - want) < 1e6]
print(f"sha {sha}  replay min {min(ts):.4f} med {statistics.median(ts):.4f} ms  c3n1s us {l3} sum {sum(l3):.1f}")
