"""On-disk formats of the reference, plus seeded synthetic fixtures.

Weight directory  — one raw native-endian float32 file per state_dict key, named by the key
                    (/root/reference/save_weights.py:8-12; readers: cuda/nn.cuh:21,58-61,113-117).
                    `*.num_batches_tracked` (int64 scalars) are written as a single 4-byte float, as
                    the reference's `struct.pack('f', x.item())` does.
Image file        — the preprocessed image [1,3,224,224] as raw float32 NCHW
                    (/root/reference/convert_imgs_to_bin.py:20-23; reader: cuda/inference/main.cu:236).

The reference scripts need the network (pretrained weights) and a GPU; here weights are seeded
random-init torchvision models (north_star: "random-init weights"), optionally with randomised
BatchNorm statistics so that BN folding is actually exercised (default-init BN is the identity).
"""
from __future__ import annotations

import os
from pathlib import Path

import numpy as np
import torch

ARCHS = ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")

# (bottleneck?, blocks per layer) — resnet152 row is cuda/inference/main.cu:116-119
ARCH_SPECS = {
    "resnet18": (False, (2, 2, 2, 2)),
    "resnet34": (False, (3, 4, 6, 3)),
    "resnet50": (True, (3, 4, 6, 3)),
    "resnet101": (True, (3, 4, 23, 3)),
    "resnet152": (True, (3, 8, 36, 3)),
}

# Algorithmic FLOPs per 224x224 image (2*MAC, convs + fc), SURVEY.md section 8(d) / BASELINE.md.
FLOPS_PER_IMAGE = {
    "resnet18": 3_628_146_688,
    "resnet50": 8_178_368_512,
    "resnet152": 23_027_253_248,
}


def make_state_dict(arch: str, seed: int = 0, randomize_bn: bool = False, num_classes: int = 1000):
    """Seeded random-init torchvision ResNet state_dict (CPU, fp32)."""
    import torchvision

    if arch not in ARCHS:
        raise ValueError(f"unknown arch {arch!r}")
    torch.manual_seed(seed)
    model = getattr(torchvision.models, arch)(weights=None, num_classes=num_classes)
    sd = model.state_dict()
    if randomize_bn:
        g = torch.Generator().manual_seed(seed + 1000)
        for key in list(sd.keys()):
            t = sd[key]
            if key.endswith("running_mean"):
                sd[key] = torch.randn(t.shape, generator=g) * 0.1
            elif key.endswith("running_var"):
                sd[key] = torch.rand(t.shape, generator=g) * 0.5 + 0.75
            elif (".bn" in key or key.startswith("bn") or "downsample.1" in key) and key.endswith(".weight"):
                sd[key] = torch.rand(t.shape, generator=g) * 0.5 + 0.75
            elif (".bn" in key or key.startswith("bn") or "downsample.1" in key) and key.endswith(".bias"):
                sd[key] = torch.randn(t.shape, generator=g) * 0.1
    return sd


def save_weights_dir(state_dict, out_dir) -> int:
    """Write `state_dict` in the save_weights.py format. Returns the number of files written."""
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    n = 0
    for name, tensor in state_dict.items():
        arr = tensor.detach().cpu().to(torch.float32).contiguous().numpy().reshape(-1)
        arr.astype("<f4" if np.little_endian else "=f4").tofile(out / name)
        n += 1
    return n


def load_tensor(path) -> torch.Tensor:
    """Raw float32 file -> 1-D tensor (pytorch_inference.py:15-28 / Tensor::loadToCpu, tensor.cuh:126-147)."""
    return torch.from_numpy(np.fromfile(path, dtype=np.float32).copy())


def load_weights_dir(arch: str, weights_dir, num_classes: int = 1000):
    """Read a weight directory back into a state_dict shaped for torchvision's `arch`."""
    import torchvision

    model = getattr(torchvision.models, arch)(weights=None, num_classes=num_classes)
    sd = model.state_dict()
    out = {}
    for name, t in sd.items():
        flat = load_tensor(Path(weights_dir) / name)
        if name.endswith("num_batches_tracked"):
            out[name] = flat.to(torch.int64).reshape(t.shape)
        else:
            out[name] = flat.reshape(t.shape)
    return out


def cached_weights_dir(arch: str, seed: int = 0, randomize_bn: bool = False, root=None) -> Path:
    """Materialise (once per machine) the seeded weight directory for `arch` and return its path."""
    root = Path(root or os.environ.get("RNB_CACHE", "/tmp/rnb_cache"))
    d = root / f"{arch}_seed{seed}{'_rbn' if randomize_bn else ''}" / "weights_bin"
    marker = d / ".complete"
    if not marker.exists():
        sd = make_state_dict(arch, seed, randomize_bn)
        save_weights_dir(sd, d)
        marker.write_text("ok")
    return d


def synthetic_images(batch: int, seed: int = 1234, size: int = 224) -> torch.Tensor:
    """Synthetic normalised-image-like input [B,3,size,size] fp32."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g)


def preprocess_jpeg(path) -> torch.Tensor:
    """JPEG -> [1,3,224,224] fp32, the transform of convert_imgs_to_bin.py:12,18 (resize 256,
    centre-crop 224, /255, normalise with the ImageNet mean/std)."""
    import torchvision
    from PIL import Image

    preprocess = torchvision.models.ResNet152_Weights.IMAGENET1K_V1.transforms()
    with open(path, "rb") as f:
        img = Image.open(f)
        img = img.convert("RGB")
        return preprocess(img).unsqueeze(0).contiguous()


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def decode_jpeg_u8(path) -> torch.Tensor:
    """JPEG -> [1,224,224,3] uint8 HWC: the first half of convert_imgs_to_bin.py:12 (resize 256 with
    antialiased bilinear, centre-crop 224) — the input of ResNet.forward_u8."""
    import numpy as np
    from PIL import Image
    from torchvision.transforms import functional as F

    with open(path, "rb") as f:
        img = Image.open(f).convert("RGB")
        img = F.center_crop(F.resize(img, [256], interpolation=F.InterpolationMode.BILINEAR, antialias=True), [224])
        return torch.from_numpy(np.asarray(img).copy()).unsqueeze(0)


def load_u8_image_bin(path, size: int = 224) -> torch.Tensor:
    import numpy as np
    return torch.from_numpy(np.fromfile(path, dtype=np.uint8).copy()).reshape(-1, size, size, 3)


def synthetic_images_u8(batch: int, seed: int = 1234, size: int = 224) -> torch.Tensor:
    """Synthetic decoded images [B,size,size,3] uint8."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8)


def save_image_bin(tensor: torch.Tensor, path) -> None:
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    tensor.detach().cpu().to(torch.float32).contiguous().numpy().reshape(-1).tofile(path)


def load_image_bin(path, size: int = 224) -> torch.Tensor:
    return load_tensor(path).reshape(-1, 3, size, size)
