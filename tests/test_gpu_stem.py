"""-m gpu: the fused stem (conv 7x7/2 + BN + ReLU + max-pool 3x3/2; main.cu:176-192) against the
plain-C oracle's unfused chain. BF16 at 224x224 runs the tcgen05 "Hankel descriptor" kernel
(stem_tc.cu); TF32 and other sizes run the CUDA-core kernel (stem.cu)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _oracle_stem(oracle_lib, x, w, bn):
    y = oracle_lib.relu(oracle_lib.batchnorm2d(oracle_lib.conv2d(x, w, 2, 3), *bn))
    return oracle_lib.maxpool2d(y, 3, 2, 1)


def _params(seed, rbn=True):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(64, 3, 7, 7, generator=g) * (2.0 / (64 * 49)) ** 0.5  # kaiming fan_out, as torchvision
    if rbn:
        bn = (torch.rand(64, generator=g) * 0.5 + 0.75, torch.randn(64, generator=g) * 0.1,
              torch.randn(64, generator=g) * 0.1, torch.rand(64, generator=g) * 0.5 + 0.75)
    else:
        bn = (torch.ones(64), torch.zeros(64), torch.zeros(64), torch.ones(64))
    return w, bn


@pytest.mark.parametrize("dtype,B,size,tol", [
    ("bf16", 1, 224, 2e-2),    # tensor-core stem
    ("bf16", 3, 224, 2e-2),
    ("tf32", 2, 224, 1e-3),    # CUDA-core stem (fp32 math, tf32-rounded output)
    ("bf16", 2, 64, 2e-2),     # CUDA-core stem, other size
])
def test_stem_against_oracle(oracle_lib, jpeg_tensor, dtype, B, size, tol):
    from resnet_c_b200 import engine, weights
    w, bn = _params(11)
    if size == 224:
        x = torch.cat([jpeg_tensor, weights.synthetic_images(B - 1, seed=3)]) if B > 1 else jpeg_tensor
    else:
        x = weights.synthetic_images(B, seed=4, size=size)
    want = _oracle_stem(oracle_lib, x, w, bn)
    got = engine.stem_forward(x.cuda(), w.cuda(), tuple(t.cuda() for t in bn), dtype).cpu().numpy()
    assert got.shape == want.shape
    e = rel_err(got.reshape(B, -1), want.reshape(B, -1))
    assert e < tol, f"stem {dtype} B={B} size={size}: rel err {e:.3e}"


def test_stem_tc_borders_and_channels(oracle_lib):
    """Impulse images: every output a single tap produces must land where the reference puts it
    (catches an off-by-one in the padded layout, the tap order or the pooling window)."""
    from resnet_c_b200 import engine
    w, bn = _params(12, rbn=False)
    w = w.abs() + 0.01                      # positive weights: ReLU never hides a misplaced tap
    x = torch.zeros(6, 3, 224, 224)
    for i, (c, r, col) in enumerate([(0, 0, 0), (1, 0, 223), (2, 223, 0), (0, 223, 223), (1, 111, 112), (2, 5, 6)]):
        x[i, c, r, col] = 1.0
    want = _oracle_stem(oracle_lib, x, w, bn)
    got = engine.stem_forward(x.cuda(), w.cuda(), tuple(t.cuda() for t in bn), "bf16").cpu().numpy()
    wq = w.to(torch.bfloat16).float()       # the kernel rounds weights to bf16; compare like with like
    want_q = _oracle_stem(oracle_lib, x, wq, bn)
    np.testing.assert_allclose(got, want_q, rtol=1e-2, atol=1e-6)
    assert (np.abs(got) > 0).sum() == (np.abs(want) > 0).sum()


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_stem_tc_unit_partitioning_is_bit_exact(dtype):
    """The tensor-core stems give each CTA a contiguous range of work units and carry the conv row two consecutive
    units of an image share in registers (stem_tc.cu / stem_tc_split.cu). A batch of 20 images (380 units on 148 CTAs:
    ranges of 2-3 units, most starting in the middle of an image) must equal, bit for bit, the same images pushed
    through one at a time (19 units on 19 CTAs: no unit continues another) and a batch of 9 (171 units: ranges of 1-2)."""
    from resnet_c_b200 import engine, weights
    w, bn = _params(13)
    x = weights.synthetic_images(20, seed=21).cuda()
    wc, bnc = w.cuda(), tuple(t.cuda() for t in bn)
    full = engine.stem_forward(x, wc, bnc, dtype).cpu()
    singles = torch.cat([engine.stem_forward(x[i:i + 1].contiguous(), wc, bnc, dtype).cpu() for i in range(20)])
    assert torch.equal(full, singles)
    nine = engine.stem_forward(x[:9].contiguous(), wc, bnc, dtype).cpu()
    assert torch.equal(full[:9], nine)
