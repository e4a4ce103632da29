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
