"""-m gpu: the FP32 NCHW per-op entry points (what cuda/nn.cu's module forwards call) against the
plain-C oracle. These keep the reference's arithmetic order, so the bar is BIT-EXACT, not a tolerance."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g)


@pytest.fixture(scope="module")
def eng():
    from resnet_c_b200 import engine
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return engine


@pytest.mark.parametrize("B,Cin,H,W,Cout,k,stride,pad", [
    (2, 1, 7, 7, 2, 2, 1, 0),
    (2, 3, 32, 32, 16, 7, 2, 3),
    (2, 8, 9, 11, 4, 3, 1, 1),
    (2, 8, 9, 11, 4, 3, 2, 1),
    (1, 64, 14, 14, 32, 1, 2, 0),
    (1, 4, 5, 5, 3, 5, 1, 2),
])
def test_conv2d_bit_exact(eng, oracle_lib, B, Cin, H, W, Cout, k, stride, pad):
    x, w = _rand(B, Cin, H, W, seed=1), _rand(Cout, Cin, k, k, seed=2)
    got = eng.conv2d_forward(x.cuda(), w.cuda(), stride, pad).cpu().numpy()
    np.testing.assert_array_equal(got, oracle_lib.conv2d(x, w, stride, pad))


def test_batchnorm_bit_exact_and_in_place(eng, oracle_lib):
    x = _rand(3, 6, 5, 7, seed=3)
    w, b, m, v = _rand(6, seed=4), _rand(6, seed=5), _rand(6, seed=6), torch.rand(6) + 0.5
    ref = oracle_lib.batchnorm2d(x, w, b, m, v)
    xd = x.cuda()
    got = eng.batchnorm2d_forward(xd, w.cuda(), b.cuda(), m.cuda(), v.cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    eng.batchnorm2d_forward(xd, w.cuda(), b.cuda(), m.cuda(), v.cuda(), out=xd)  # x == out, main.cu:145
    np.testing.assert_array_equal(xd.cpu().numpy(), ref)


def test_relu_add_bit_exact(eng, oracle_lib):
    a, b = _rand(1000, 37, seed=7), _rand(1000, 37, seed=8)
    np.testing.assert_array_equal(eng.relu_forward(a.cuda()).cpu().numpy(), oracle_lib.relu(a))
    np.testing.assert_array_equal(eng.add_forward(a.cuda(), b.cuda()).cpu().numpy(), oracle_lib.add(a, b))
    ad = a.cuda()
    eng.add_forward(ad, b.cuda(), out=ad)  # in place on `a`, main.cu:162
    np.testing.assert_array_equal(ad.cpu().numpy(), oracle_lib.add(a, b))


@pytest.mark.parametrize("H,W,k,stride,pad", [(112, 112, 3, 2, 1), (9, 7, 3, 2, 1), (8, 8, 2, 2, 0)])
def test_maxpool_bit_exact(eng, oracle_lib, H, W, k, stride, pad):
    x = _rand(2, 5, H, W, seed=9)
    np.testing.assert_array_equal(eng.maxpool2d_forward(x.cuda(), k, stride, pad).cpu().numpy(),
                                  oracle_lib.maxpool2d(x, k, stride, pad))


def test_avgpool_bit_exact(eng, oracle_lib):
    x = _rand(2, 64, 7, 7, seed=10)
    np.testing.assert_array_equal(eng.avgpool2d_forward(x.cuda(), 7).cpu().numpy(), oracle_lib.avgpool2d(x, 7))
    np.testing.assert_array_equal(eng.avgpool2d_forward(x.cuda(), 3, 2, 1).cpu().numpy(),
                                  oracle_lib.avgpool2d(x, 3, 2, 1))


def test_linear_bit_exact(eng, oracle_lib):
    x, w, b = _rand(3, 512, seed=11), _rand(100, 512, seed=12), _rand(100, seed=13)
    np.testing.assert_array_equal(eng.linear_forward(x.cuda(), w.cuda(), b.cuda()).cpu().numpy(),
                                  oracle_lib.linear(x, w, b))
    np.testing.assert_array_equal(eng.linear_forward(x.cuda(), w.cuda(), None).cpu().numpy(),
                                  oracle_lib.linear(x, w, None))


def test_argmax_tie_rule(eng, oracle_lib):
    x = _rand(64, 1000, seed=14)
    x[3, 10] = x[3, 700] = 99.0   # tie: the first maximum must win (main.cu:246 uses '<')
    x[5, :] = 0.0
    np.testing.assert_array_equal(eng.argmax_forward(x.cuda()).cpu().numpy(), oracle_lib.argmax_rows(x.numpy()))
    assert int(eng.argmax_forward(x.cuda())[3]) == 10


def test_softmax_topk_matches_oracle(eng, tmp_path):
    """rnb_softmax_topk_forward against oracle/postprocess.py (float64): probabilities to 1e-6, indices exact
    (value descending, lowest index first on ties); rnb_save_f32 writes Tensor::save's raw format."""
    from oracle import postprocess
    x = _rand(37, 1000, seed=21) * 3.0
    x[4, 17] = x[4, 903] = x[4].max() + 1.0      # a tie for first place
    x[9, :] = 0.25                                # all equal: indices 0..k-1
    for k in (1, 5, 32):
        top_p, top_i, probs = eng.softmax_topk_forward(x.cuda(), k=k, full=True)
        ref_p, ref_i, ref_full = postprocess.softmax_topk(x.numpy(), k)
        np.testing.assert_array_equal(top_i.cpu().numpy(), ref_i)
        np.testing.assert_allclose(top_p.cpu().numpy(), ref_p, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(probs.cpu().numpy(), ref_full, rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(probs.cpu().numpy().sum(1), 1.0, atol=1e-5)
    assert top_i[4, :2].cpu().tolist() == [17, 903] and top_i[9, :4].cpu().tolist() == [0, 1, 2, 3]
    out = tmp_path / "cuda_out.bin"
    eng.save_f32(x.cuda(), out)
    np.testing.assert_array_equal(postprocess.load_f32(out), x.numpy().reshape(-1))


def test_tail_matches_oracle(eng, oracle_lib):
    x, w, b = _rand(5, 512, 7, 7, seed=15), _rand(1000, 512, seed=16) * 0.05, _rand(1000, seed=17)
    logits, top1 = eng.tail_forward(x.cuda(), w.cuda(), b.cuda())
    pooled = oracle_lib.avgpool2d(x, 7).reshape(5, 512)
    ref = oracle_lib.linear(pooled, w, b)
    # different (tiled) summation order than the sequential reference loop: fp32 tolerance
    np.testing.assert_allclose(logits.cpu().numpy(), ref, rtol=2e-5, atol=2e-5)
    np.testing.assert_array_equal(top1.cpu().numpy(), oracle_lib.argmax_rows(ref))


def test_errors_are_reported_not_fatal(eng):
    from resnet_c_b200._lib import RnbError
    with pytest.raises(RnbError):
        eng.conv_bn_act_forward(torch.zeros(1, 3, 8, 8).cuda(), torch.zeros(64, 3, 3, 3).cuda())  # Cin % 64
    with pytest.raises(RnbError):
        eng.ResNet("resnet50", "/nonexistent_dir", max_batch=1)
    with pytest.raises(RnbError):
        eng.ResNet("resnet51", "/tmp", max_batch=1)
