"""Resize 256 + centre-crop 224 (/root/reference/convert_imgs_to_bin.py:12,18).

CPU: the numpy oracle (oracle/preprocess.py::resize_crop_u8, a restatement of Pillow's Resample.c) against Pillow /
torchvision live, against the committed digests, and on the reference's own test image against the committed crop.
GPU (-m gpu): rnb_resize_crop_u8 against the oracle, bit for bit, and chained into the network."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from make_golden_resize import SIZES, synthetic  # noqa: E402

from oracle import preprocess  # noqa: E402


@pytest.fixture(scope="module")
def cases(golden_dir):
    return np.load(golden_dir / "resize_crop_cases.npz")


def _decoded(golden_dir):
    from PIL import Image
    return np.asarray(Image.open(golden_dir / "ILSVRC2012_val_00004749_decoded.png").convert("RGB")).copy()


def test_oracle_reproduces_committed_pillow_digests(cases):
    assert [tuple(s) for s in cases["sizes"]] == SIZES
    for i, (h, w) in enumerate(SIZES):
        got = preprocess.resize_crop_u8(synthetic(h, w, 100 + i))
        assert hashlib.sha256(got.tobytes()).hexdigest() == str(cases["sha256"][i]), (h, w)
    assert np.array_equal(preprocess.resize_crop_u8(synthetic(375, 500, 101)), cases["full_375x500"])


def test_oracle_equals_pillow_live():
    from make_golden_resize import pillow_resize_crop
    for i, (h, w) in enumerate([(333, 517), (224, 224), (256, 300), (719, 1280)]):
        a = synthetic(h, w, 7 + i)
        assert np.array_equal(preprocess.resize_crop_u8(a), pillow_resize_crop(a)), (h, w)


def test_oracle_on_the_reference_image_gives_the_committed_crop(golden_dir):
    gold = np.fromfile(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin", dtype=np.uint8).reshape(224, 224, 3)
    assert np.array_equal(preprocess.resize_crop_u8(_decoded(golden_dir)), gold)


def test_size_and_offset_rules():
    assert preprocess.resized_size(500, 492) == (260, 256)
    assert preprocess.resized_size(375, 500) == (256, 341)
    # Python's round: half to even
    assert [preprocess.crop_offset(s, 224) for s in (256, 257, 259, 341)] == [16, 16, 18, 58]


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(SIZES)))
def test_gpu_resize_crop_matches_oracle(idx):
    from resnet_c_b200 import engine
    h, w = SIZES[idx]
    imgs = np.stack([synthetic(h, w, 100 + idx), synthetic(h, w, 900 + idx), np.zeros((h, w, 3), np.uint8),
                     np.full((h, w, 3), 255, np.uint8)])
    got = engine.resize_crop_u8(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i in range(len(imgs)):
        assert np.array_equal(got[i], preprocess.resize_crop_u8(imgs[i])), (h, w, i)


@pytest.mark.gpu
def test_gpu_reference_image_to_top1(golden_dir):
    """decoded reference image -> GPU resize + crop (== committed crop) -> uint8 network input -> the index the
    reference prints for this image with seed-0 ResNet-18 weights (BASELINE configs[0]: 238)."""
    from resnet_c_b200 import engine, weights
    dec = torch.from_numpy(_decoded(golden_dir)).unsqueeze(0).cuda()
    crop = engine.resize_crop_u8(dec)
    gold = weights.load_u8_image_bin(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin")
    assert torch.equal(crop.cpu(), gold)
    m = engine.ResNet("resnet18", weights.cached_weights_dir("resnet18", 0), dtype="tf32", max_batch=1)
    _, top1 = m.forward_u8(crop)
    torch.cuda.synchronize()
    assert top1.cpu().tolist() == [238]
    m.close()
