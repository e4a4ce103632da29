"""Small end-to-end pass over the round-2 code paths for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
BF16 / TF32 / FP8 models at batch 3, the tail-split and small-grid split conv, a planned Bottleneck block, the GPU
resize + crop, two replicas on one GPU. Prints a checksum per path; the sanitizer's own summary is the result."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import os  # noqa: E402

import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

os.environ["RNB_AUTOTUNE"] = "0"
x = weights.synthetic_images(3).cuda()
for arch, dtype in (("resnet18", "tf32"), ("resnet50", "bf16"), ("resnet50", "fp8")):
    m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0, True), dtype=dtype, max_batch=3)
    lg, t1 = m.forward(x)
    lg2, _ = m.forward(x[:2].contiguous())
    torch.cuda.synchronize()
    print(arch, dtype, float(lg.abs().sum()), t1.tolist(), bool(torch.equal(lg[:2], lg2)))
    m.close()
g = torch.Generator().manual_seed(0)
os.environ["RNB_FORCE_TILE"] = "1256"
xc = torch.randn(8, 256, 14, 14, generator=g).cuda()
wc = (torch.randn(256, 256, 3, 3, generator=g) * 0.03).cuda()
print("split conv", float(engine.conv_bn_act_forward(xc, wc, None, None, True, 1, 1, "bf16").abs().sum()))
print("fp8 conv", float(engine.conv_fp8_forward(xc, wc, None, None, True, 1, 1, 0.02, 1.0, 0.05).abs().sum()))
del os.environ["RNB_FORCE_TILE"]
convs = [dict(w=(torch.randn(64, 64, 1, 1, generator=g) * 0.1).cuda(), bn=None, stride=1, pad=0),
         dict(w=(torch.randn(64, 64, 3, 3, generator=g) * 0.05).cuda(), bn=None, stride=1, pad=1),
         dict(w=(torch.randn(256, 64, 1, 1, generator=g) * 0.1).cuda(), bn=None, stride=1, pad=0),
         dict(w=(torch.randn(256, 64, 1, 1, generator=g) * 0.1).cuda(), bn=None, stride=1, pad=0)]
blk = engine.Block("bottleneck", convs, "bf16")
print("block", float(blk.forward(torch.randn(2, 64, 56, 56, generator=g).cuda()).abs().sum()), blk.num_launches(2, 56, 56))
blk.close()
img = torch.randint(0, 256, (2, 375, 500, 3), generator=g, dtype=torch.uint8).cuda()
print("resize", int(engine.resize_crop_u8(img).sum()))
grp = engine.ResNetGroup("resnet18", weights.cached_weights_dir("resnet18", 0, True), [0, 0], max_batch_per_device=2)
lg, t1 = grp.forward([x[:2].contiguous(), x[2:3].contiguous()])
grp.synchronize()
print("group", float(lg.abs().sum()), t1.tolist())
grp.close()
print("done")
