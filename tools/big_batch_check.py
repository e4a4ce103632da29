"""GPU check: a large per-GPU batch (BASELINE configs[3]: 2048 images over 2 GPUs = 1024 per GPU) gives, image by
image, bit-identical logits to the same images run at batch 256 (fused layer1/2/3 kernels, tensor-core FC)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
big = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
wdir = weights.cached_weights_dir(arch, 0, True)
x = weights.synthetic_images(big).cuda()
m_big = engine.ResNet(arch, wdir, dtype="bf16", max_batch=big)
lb, tb = m_big.forward(x)
torch.cuda.synchronize()
m_small = engine.ResNet(arch, wdir, dtype="bf16", max_batch=256)
ok = True
for off in range(0, big, 256):
    ls, ts = m_small.forward(x[off:off + 256].contiguous())
    torch.cuda.synchronize()
    same = torch.equal(ls, lb[off:off + 256]) and torch.equal(ts, tb[off:off + 256])
    ok = ok and same
    print(f"images {off}..{off + 255}: identical={same}")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    m_big.forward(x, lb, tb)
e0.record()
for _ in range(10):
    m_big.forward(x, lb, tb)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"{arch} bf16 B={big}: {ms:.3f} ms/step, {big / ms * 1e3:.0f} img/s")
print("BIG_BATCH", "OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
