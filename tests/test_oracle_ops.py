"""The plain-C oracle (oracle/ref_ops.c) against torch.nn.functional on CPU — the oracle must be
right before anything is compared with it. Also covers the edge cases the reference kernels have
(clipped windows, divisor of the average pool, in-place batch-norm, arg-max tie rule)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,stride,pad", [
    (2, 1, 7, 7, 2, 2, 1, 0),      # the shape of the reference's conv2dTest (cuda/test.cu:4-96)
    (1, 3, 20, 20, 8, 7, 2, 3),    # stem-like
    (2, 8, 9, 11, 4, 3, 1, 1),     # ragged, non-square
    (2, 8, 9, 11, 4, 3, 2, 1),
    (1, 16, 8, 8, 16, 1, 2, 0),    # downsample-like
    (1, 4, 5, 5, 3, 5, 1, 2),      # window as large as the image
])
def test_conv2d(oracle_lib, B, Cin, H, W, Cout, k, stride, pad):
    x, w = _rand(B, Cin, H, W, seed=1), _rand(Cout, Cin, k, k, seed=2)
    got = oracle_lib.conv2d(x, w, stride, pad)
    ref = F.conv2d(x, w, stride=stride, padding=pad)
    assert got.shape == tuple(ref.shape)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=1e-5)


def test_conv_output_size(oracle_lib):
    for x, k, s, p in [(224, 7, 2, 3), (112, 3, 2, 1), (56, 3, 1, 1), (56, 1, 2, 0), (7, 7, 1, 0), (15, 3, 2, 1)]:
        assert oracle_lib.conv_output_size(x, k, s, p) == (2 * p + x - k) // s + 1


def test_batchnorm_matches_torch_and_is_double_inside(oracle_lib):
    x = _rand(2, 5, 4, 3, seed=3)
    w, b, m, v = _rand(5, seed=4), _rand(5, seed=5), _rand(5, seed=6), torch.rand(5) + 0.5
    got = oracle_lib.batchnorm2d(x, w, b, m, v)
    ref = F.batch_norm(x, m, v, w, b, training=False, eps=1e-5)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-6, atol=1e-6)
    # the reference evaluates in double after a float subtraction (ops.cu:149-150)
    xd = (x - m.view(1, -1, 1, 1)).double()
    exact = (xd / torch.sqrt(v.double() + 1e-5).view(1, -1, 1, 1) * w.double().view(1, -1, 1, 1)
             + b.double().view(1, -1, 1, 1)).float()
    assert np.abs(got - exact.numpy()).max() <= 1.2e-7 * float(exact.abs().max())


def test_relu_add(oracle_lib):
    a, b = _rand(3, 17, seed=7), _rand(3, 17, seed=8)
    np.testing.assert_array_equal(oracle_lib.relu(a), F.relu(a).numpy())
    np.testing.assert_array_equal(oracle_lib.add(a, b), (a + b).numpy())
    # alternating signs, N = 17: the reference's reluTest input (cuda/test.cu:176-215)
    alt = torch.tensor([(-1.0) ** i * i for i in range(17)])
    np.testing.assert_array_equal(oracle_lib.relu(alt), torch.clamp(alt, min=0).numpy())


@pytest.mark.parametrize("H,W,k,stride,pad", [(12, 12, 3, 2, 1), (9, 7, 3, 2, 1), (8, 8, 2, 2, 0), (5, 5, 3, 1, 1)])
def test_maxpool(oracle_lib, H, W, k, stride, pad):
    x = _rand(2, 3, H, W, seed=9)
    got = oracle_lib.maxpool2d(x, k, stride, pad)
    ref = F.max_pool2d(x, k, stride, pad)
    np.testing.assert_array_equal(got, ref.numpy())


def test_avgpool_global_and_padded(oracle_lib):
    x = _rand(2, 6, 7, 7, seed=10)
    got = oracle_lib.avgpool2d(x, 7)
    ref = F.adaptive_avg_pool2d(x, 1)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-6, atol=1e-7)
    # clipped windows still divide by k*k (ops.cu:107) == count_include_pad=True
    got = oracle_lib.avgpool2d(x, 3, 2, 1)
    ref = F.avg_pool2d(x, 3, 2, 1, count_include_pad=True)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-6, atol=1e-7)


def test_linear(oracle_lib):
    # B=3, 16 -> 8: the reference's linearTest shape (cuda/test.cu:98-174)
    x, w, b = _rand(3, 16, seed=11), _rand(8, 16, seed=12), _rand(8, seed=13)
    np.testing.assert_allclose(oracle_lib.linear(x, w, b), F.linear(x, w, b).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(oracle_lib.linear(x, w, None), F.linear(x, w).numpy(), rtol=1e-5, atol=1e-5)


def test_argmax_first_maximum_wins(oracle_lib):
    x = np.array([[1, 5, 5, 2], [7, 7, 7, 7], [-3, -1, -2, -1], [0, 0, 1, 0]], np.float32)
    np.testing.assert_array_equal(oracle_lib.argmax_rows(x), [1, 0, 1, 2])


def test_softmax_topk_oracle_pinned_against_torch():
    """oracle/postprocess.py::softmax_topk == torch.softmax / torch.topk (float64) incl. the tie rule."""
    import torch
    from oracle import postprocess
    g = torch.Generator().manual_seed(5)
    x = torch.randn(16, 1000, generator=g, dtype=torch.float64) * 4
    p, i, full = postprocess.softmax_topk(x.numpy(), 5)
    tp = torch.softmax(x, dim=1)
    np.testing.assert_allclose(full, tp.numpy(), rtol=1e-12, atol=1e-15)
    tv, ti = torch.topk(tp, 5, dim=1)
    np.testing.assert_array_equal(i, ti.numpy())
    np.testing.assert_allclose(p, tv.numpy(), rtol=1e-12)
    y = np.zeros((1, 8)); y[0, 2] = y[0, 6] = 1.0
    assert postprocess.softmax_topk(y, 3)[1].tolist() == [[2, 6, 0]]
