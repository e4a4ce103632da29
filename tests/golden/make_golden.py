"""Generates the committed golden fixtures under tests/golden/. Run in the BUILD container, where
/root/reference is mounted:

    python tests/golden/make_golden.py

What it pins
  1. ILSVRC2012_val_00004749.bin — the reference's only test image pushed through the preprocessing
     of /root/reference/convert_imgs_to_bin.py:12,18 (the [1,3,224,224] float32 file main.cu:236 reads).
  2. ref_class_resnet152.npz — logits of the REFERENCE'S OWN PyTorch model: the class definitions of
     /root/reference/pytorch_inference.py:29-162 are exec'd from that file (nothing is copied) and run
     on CPU with the seeded ResNet-152 weights on the image above. This is "output of the reference
     itself run here"; oracle/torch_model.py must reproduce it bit for bit.
  3. <arch>_<tag>.npz — oracle logits (fp32 CPU = the reference path, fp64 = arbiter for near-ties),
     top-1 and top-1/top-2 margins for the parity configs of BASELINE.json, plus a checksum of the
     seeded weights so a test can tell "the RNG changed" from "the kernel is wrong".
The reference ships no golden vectors of its own, so these are the pin.
"""
import re
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))

from oracle import torch_model  # noqa: E402
from resnet_c_b200 import weights  # noqa: E402

REFERENCE = Path("/root/reference")

# (arch, randomize_bn, input tag, batch)
CASES = [
    ("resnet18", False, "jpeg", 1),     # BASELINE.json configs[0]
    ("resnet18", True, "jpeg", 1),
    ("resnet18", True, "synth", 4),
    ("resnet50", True, "synth", 4),
    ("resnet50", False, "synth", 2),
    ("resnet152", True, "synth", 2),
    ("resnet152", False, "jpeg", 1),    # what the reference's main.cu / pytorch_inference.py run
]


def weights_checksum(sd) -> float:
    return float(sum(v.double().sum().item() for k, v in sd.items() if v.dtype.is_floating_point))


def reference_model_classes():
    """exec the class/def section of the reference's pytorch_inference.py (no module-level script)."""
    src = (REFERENCE / "pytorch_inference.py").read_text()
    start = src.index("class ResnetBlock")
    end = src.index("\nC = 3")
    header = "import torch\nimport torch.nn as nn\nimport torch.nn.functional as F\n"
    ns = {}
    exec(compile(header + src[start:end], "pytorch_inference_classes", "exec"), ns)
    return ns


def inputs(tag, batch, jpeg):
    if tag == "jpeg":
        return jpeg.repeat(batch, 1, 1, 1)
    return weights.synthetic_images(batch)


def main():
    torch.set_num_threads(1)  # fixed summation order for the committed numbers
    jpeg_path = REFERENCE / "test_imgs" / "ILSVRC2012_val_00004749.jpeg"
    jpeg = weights.preprocess_jpeg(jpeg_path)
    weights.save_image_bin(jpeg, HERE / "ILSVRC2012_val_00004749.bin")
    # the decoded / resized / cropped uint8 HWC image before normalisation (input of rnb_model_forward_u8)
    weights.decode_jpeg_u8(jpeg_path).numpy().tofile(HERE / "ILSVRC2012_val_00004749_u8hwc.bin")
    print("image", tuple(jpeg.shape), float(jpeg.min()), float(jpeg.max()), float(jpeg.mean()))

    # 2. the reference's own class
    ns = reference_model_classes()
    sd152 = weights.make_state_dict("resnet152", 0)
    ref = ns["Resnet152"](1000)
    ref.load_state_dict(sd152, strict=True)
    ref.eval()
    with torch.no_grad():
        y_ref = ref(jpeg)
    np.savez(HERE / "ref_class_resnet152.npz", logits=y_ref.numpy(), top1=y_ref.argmax(1).numpy(),
             weights_checksum=weights_checksum(sd152))
    print("reference Resnet152 class: top1", y_ref.argmax(1).tolist())

    for arch, rbn, tag, batch in CASES:
        sd = weights.make_state_dict(arch, 0, randomize_bn=rbn)
        x = inputs(tag, batch, jpeg)
        y32 = torch_model.run(arch, sd, x)
        y64 = torch_model.run(arch, sd, x, torch.float64)
        srt = y64.sort(dim=1, descending=True).values
        margin = ((srt[:, 0] - srt[:, 1]) / y64.abs().amax(1)).numpy()
        name = f"{arch}_{'rbn' if rbn else 'default'}_{tag}_b{batch}"
        np.savez(HERE / f"{name}.npz", logits_fp32=y32.numpy(), logits_fp64=y64.numpy(),
                 top1=y64.argmax(1).numpy().astype(np.int32), margin_rel=margin,
                 weights_checksum=weights_checksum(sd))
        print(name, "top1", y64.argmax(1).tolist(), "margin_rel", margin.round(5).tolist(),
              "fp32-vs-fp64 rel", float((y32.double() - y64).abs().max() / y64.abs().max()))


if __name__ == "__main__":
    main()
