"""TEST ORACLE (not part of the product path; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this). CPU restatement of the normalisation step of the reference's fixture generator:

    /root/reference/convert_imgs_to_bin.py:12   preprocess = ResNet152_Weights.IMAGENET1K_V1.transforms()
    /root/reference/convert_imgs_to_bin.py:18   img = preprocess(img)

i.e. torchvision's ImageClassification preset: resize 256 (antialiased bilinear) -> centre-crop 224 ->
pil_to_tensor -> convert_image_dtype(float) [= u8 / 255] -> normalize(mean, std) [= (x - mean) / std], all
in FP32. `normalize_u8` restates the last two steps on the decoded uint8 HWC crop; pinned bit-exactly against
the committed golden tensor the full preset produced from the reference's test JPEG
(tests/test_oracle_pin.py::test_u8_normalisation_reproduces_the_reference_preprocessing).
"""
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def normalize_u8(x_u8_hwc: torch.Tensor, mean=IMAGENET_MEAN, std=IMAGENET_STD) -> torch.Tensor:
    """[B,H,W,3] uint8 -> [B,3,H,W] fp32 normalised, rounding exactly as torchvision does."""
    assert x_u8_hwc.dtype == torch.uint8 and x_u8_hwc.shape[-1] == 3
    x = x_u8_hwc.permute(0, 3, 1, 2).to(torch.float32).div(255.0)
    m = torch.tensor(mean, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=torch.float32).view(1, 3, 1, 1)
    return ((x - m) / s).contiguous()


# ---------------------------------------------------------------------------------------------------------------
# resize 256 + centre-crop 224: the FIRST half of the preset (convert_imgs_to_bin.py:12,18 call it on a PIL image, so
# torchvision's F.resize / F.center_crop dispatch to Pillow). The arithmetic lives in a third-party dependency that
# is not vendored in the reference: Pillow (12.2.0 installed here; the reference pins no version),
# src/libImaging/Resample.c — ImagingResample with the bilinear filter, 8 bits per channel:
#   precompute_coeffs():  scale = in / out, filterscale = max(scale, 1), support = 1.0 * filterscale,
#                         window [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image,
#                         weights bilinear((x - center + 0.5) / filterscale) normalised to sum 1 (double precision)
#   normalize_coeffs_8bpc(): fixed point, 22 fractional bits, round half away from zero
#   ImagingResampleHorizontal_8bpc / Vertical_8bpc: acc = 2^21 + sum(pixel * k); out = clamp(acc >> 22, 0, 255)
#   horizontal pass first, its uint8 result feeds the vertical pass.
# torchvision: output size (_compute_resized_output_size: short side -> 256, long side int(256 * long / short)),
# crop offsets int(round((size - 224) / 2.0)) with Python's round-half-even. Pinned against Pillow / torchvision
# themselves in tests/test_oracle_pin.py and against the committed golden crop of the reference's test JPEG.
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resized_size(h: int, w: int, resize: int = 256):
    """(new_h, new_w) of torchvision F.resize(img, [resize])."""
    short, long_ = (w, h) if w <= h else (h, w)
    new_short, new_long = resize, int(resize * long_ / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_offset(size: int, crop: int) -> int:
    """torchvision F.center_crop: int(round((size - crop) / 2.0)), Python round = half to even."""
    return int(round((size - crop) / 2.0))


def resample_coeffs(in_size: int, out_size: int):
    """Pillow precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter over the whole axis.
    Returns (bounds [out,2] int32 = (first input index, tap count), kk [out,ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        if ww != 0.0:
            w[:xmax] /= ww
        for x in range(ksize):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One 8bpc pass along axis 0 of a uint8 array."""
    if img.shape[0] == out_size:
        return img
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        lo, n = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += src[lo + x] * int(kk[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_crop_u8(img_hwc: np.ndarray, resize: int = 256, crop: int = 224) -> np.ndarray:
    """[H,W,3] uint8 decoded image -> [crop,crop,3] uint8: resize (short side -> `resize`, antialiased bilinear,
    Pillow arithmetic) then centre crop. Integer exact."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    h, w = img_hwc.shape[:2]
    nh, nw = resized_size(h, w, resize)
    hor = _resample_axis0(img_hwc.transpose(1, 0, 2), nw).transpose(1, 0, 2)  # horizontal pass first
    full = _resample_axis0(hor, nh)
    top, left = crop_offset(nh, crop), crop_offset(nw, crop)
    return np.ascontiguousarray(full[top:top + crop, left:left + crop])
