"""GPU: A/B harness used for the stem kernel's epilogue variants (an RNB_STEM_EPI switch read per launch; the variants
themselves were measured slower and removed from stem_tc.cu — profiles/stem_r2.md section 3 keeps the numbers): per-launch
time of the stem from the un-graphed event profile and a checksum of the logits (all variants agreed bit for bit).
Run with RNB_NO_GRAPH=1 so that forward() launches the kernels directly; with one variant it is a stem timer."""
import hashlib
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
from resnet_c_b200 import engine, weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1"]
switch = sys.argv[3] if len(sys.argv) > 3 else "RNB_STEM_FORM"   # a switch the library reads per launch
m = engine.ResNet("resnet50", weights.cached_weights_dir("resnet50", 0, True), dtype="bf16", max_batch=B)
x = weights.synthetic_images(B).cuda()
logits, top1 = m.forward(x)
for rep in range(2):
    for v in variants:
        os.environ[switch] = v
        m.forward(x, logits, top1)
        torch.cuda.synchronize()
        sha = hashlib.sha256(logits.cpu().numpy().tobytes()).hexdigest()[:12]
        prof = m.profile(x, iters=5)
        print(f"{switch}={v}: stem launches {prof[0]['ms'] * 1e3:.1f} + {prof[1]['ms'] * 1e3:.1f} us, logits {sha}", flush=True)
