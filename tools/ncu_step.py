"""One forward step between cudaProfilerStart / Stop, for `ncu --profile-from-start off`:

    python tools/ncu_step.py resnet50 256 bf16            # plain run first (exit 0), then the same line under ncu

Three warm-up forwards (planning, autotune off via RNB_AUTOTUNE=0 in the caller's environment if wanted, FP8
calibration), then exactly one profiled forward replayed from the CUDA graph."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

arch, B, dtype = sys.argv[1], int(sys.argv[2]), sys.argv[3]
m = engine.ResNet(arch, weights.cached_weights_dir(arch, 0), dtype=dtype, max_batch=B)
x = weights.synthetic_images(B).cuda()
logits, top1 = m.forward(x)
for _ in range(3):
    m.forward(x, logits, top1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m.forward(x, logits, top1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(arch, B, dtype, "top1[0]", int(top1[0]), "launches", m.launches_per_forward(B))
