"""GPU: A/B of RNB_BALANCE (balanced persistent pair grids, conv_plan.cu::pair_count). Prints the logits checksum and the
CUDA-graph replay time (min / median of 8 x 30 steps) for `arch batch`."""
import hashlib
import statistics
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
sha = hashlib.sha256(logits.cpu().numpy().tobytes()).hexdigest()[:16]
for _ in range(10):
    m.forward(x, logits, top1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for rep in range(8):
    e0.record()
    for _ in range(30):
        m.forward(x, logits, top1)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 30)
print(f"{arch} B={B} sha {sha}  replay min {min(ts):.4f} med {statistics.median(ts):.4f} ms")
