"""Golden vectors for the BASELINE configs at their STATED sizes (round 2): the oracle cannot run 256 ResNet-50 or
128 ResNet-152 images in seconds, so a fixed sample of the very batch `bench.py` forwards (synthetic_images(B, seed
1234), seed-0 DEFAULT-init weights — the bench's weights, not the randomised-BN test weights) is pinned here:

    python tests/golden/make_golden_fullsize.py        (build container; needs no /root/reference)

Per case: the sampled image indices, oracle logits fp32 (the reference path, pytorch_inference.py:29-162 restated in
oracle/torch_model.py) and fp64 (arbiter for near-ties), top-1 and the fp64 top-1 / top-2 margin.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import torch_model  # noqa: E402
from resnet_c_b200 import weights  # noqa: E402

# (arch, batch, sampled indices) — configs[2] (ResNet-50 B=256 per GPU), configs[4] (ResNet-152, 128 per GPU),
# configs[1] (ResNet-18 B=256)
CASES = [
    ("resnet50", 256, [0, 36, 73, 109, 146, 182, 219, 255]),
    ("resnet152", 128, [0, 42, 85, 127]),
    ("resnet18", 256, [0, 36, 73, 109, 146, 182, 219, 255]),
]


def main():
    torch.set_num_threads(1)  # fixed summation order for the committed numbers
    for arch, batch, idx in CASES:
        sd = weights.make_state_dict(arch, 0)
        x = weights.synthetic_images(batch)[idx]
        y32 = torch_model.run(arch, sd, x)
        y64 = torch_model.run(arch, sd, x, torch.float64)
        srt = y64.sort(dim=1, descending=True).values
        margin = ((srt[:, 0] - srt[:, 1]) / y64.abs().amax(1)).numpy()
        name = f"{arch}_default_synth_b{batch}_sample{len(idx)}"
        np.savez(HERE / f"{name}.npz", index=np.asarray(idx, dtype=np.int32), logits_fp32=y32.numpy(),
                 logits_fp64=y64.numpy(), top1=y64.argmax(1).numpy().astype(np.int32), margin_rel=margin)
        print(name, "top1", y64.argmax(1).tolist(), "margin_rel", margin.round(5).tolist(),
              "fp32-vs-fp64 rel", float((y32.double() - y64).abs().max() / y64.abs().max()))


if __name__ == "__main__":
    main()
