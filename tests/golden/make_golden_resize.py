"""Golden fixtures for the resize 256 + centre-crop 224 step. Run in the BUILD container:

    python tests/golden/make_golden_resize.py

Writes
  ILSVRC2012_val_00004749_decoded.png — the reference's test JPEG decoded by Pillow (lossless container of the decoded
      pixels, so the GPU box — which has no /root/reference — can run the resize on the reference's own image; its
      resize + crop must equal the committed ILSVRC2012_val_00004749_u8hwc.bin)
  resize_crop_cases.npz — for seeded synthetic images of several sizes (down-scaling, up-scaling, identity axis,
      odd crop offsets): SHA-256 of what torchvision's F.resize + F.center_crop (Pillow underneath — the call chain of
      /root/reference/convert_imgs_to_bin.py:12,18) produce, and the full output of two of them.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
from PIL import Image
from torchvision.transforms import functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
REFERENCE = Path("/root/reference")

# (H, W) of the decoded image
SIZES = [(500, 492), (375, 500), (256, 256), (224, 224), (300, 200), (200, 333), (257, 1000), (1024, 768), (231, 229),
         (480, 640), (1080, 1920), (64, 48)]


def synthetic(h, w, seed):
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3), dtype=np.uint8)


def pillow_resize_crop(a):
    img = F.center_crop(F.resize(Image.fromarray(a), [256], interpolation=F.InterpolationMode.BILINEAR, antialias=True),
                        [224])
    return np.asarray(img).copy()


def main():
    Image.open(REFERENCE / "test_imgs" / "ILSVRC2012_val_00004749.jpeg").convert("RGB").save(
        HERE / "ILSVRC2012_val_00004749_decoded.png", optimize=True)
    out = {"sizes": np.array(SIZES, np.int32)}
    digests = []
    for i, (h, w) in enumerate(SIZES):
        got = pillow_resize_crop(synthetic(h, w, 100 + i))
        digests.append(hashlib.sha256(got.tobytes()).hexdigest())
        if (h, w) in ((375, 500), (64, 48)):
            out[f"full_{h}x{w}"] = got
    out["sha256"] = np.array(digests)
    np.savez_compressed(HERE / "resize_crop_cases.npz", **out)
    print("wrote", len(SIZES), "cases")


if __name__ == "__main__":
    main()
