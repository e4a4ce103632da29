"""-m gpu: the module classes' planned path — rnb_block_* (include/rnb.h) and, through it, Bottleneck::forward /
BasicBlock::forward / Conv2d::forward of cuda/nn.cuh — against the plain-C oracle running the reference's unfused
chain (layerForward, /root/reference/cuda/inference/main.cu:127-166: [downsample conv -> bn] conv1 -> bn -> relu ->
conv2 -> bn -> relu -> conv3 -> bn -> add -> relu). Tolerances: BASELINE.json north_star (2e-2 BF16, 1e-3 TF32)."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu
TOL = {"bf16": 2e-2, "tf32": 1e-3}


def _conv(g, cin, cout, k, stride, pad):
    w = torch.randn(cout, cin, k, k, generator=g) * (2.0 / (cin * k * k)) ** 0.5
    bn = (torch.rand(cout, generator=g) * 0.5 + 0.75, torch.randn(cout, generator=g) * 0.1,
          torch.randn(cout, generator=g) * 0.1, torch.rand(cout, generator=g) * 0.5 + 0.75)
    return dict(w=w, bn=bn, stride=stride, pad=pad)


def _oracle_cbn(o, c, x, relu, res=None):
    y = o.batchnorm2d(o.conv2d(x, c["w"], c["stride"], c["pad"]), *c["bn"])
    if res is not None:
        y = o.add(y, res)
    return o.relu(y) if relu else y


def _bottleneck(g, cin, mid, cout, stride, ds):
    convs = [_conv(g, cin, mid, 1, 1, 0), _conv(g, mid, mid, 3, stride, 1), _conv(g, mid, cout, 1, 1, 0)]
    if ds:
        convs.append(_conv(g, cin, cout, 1, stride, 0))
    return convs


def _oracle_bottleneck(o, convs, x):
    x = x.numpy() if isinstance(x, torch.Tensor) else x
    short = _oracle_cbn(o, convs[3], x, False) if len(convs) == 4 else x
    t = _oracle_cbn(o, convs[0], x, True)
    t = _oracle_cbn(o, convs[1], t, True)
    return _oracle_cbn(o, convs[2], t, True, short)


def _oracle_basic(o, convs, x):
    short = _oracle_cbn(o, convs[2], x, False) if len(convs) == 3 else x
    t = _oracle_cbn(o, convs[0], x, True)
    return _oracle_cbn(o, convs[1], t, True, short)


def _cuda(convs):
    return [dict(w=c["w"].cuda(), bn=tuple(t.cuda() for t in c["bn"]), stride=c["stride"], pad=c["pad"]) for c in convs]


@pytest.mark.parametrize("case", [
    # (cin, mid, cout, stride, downsample, H, B, dtype, expected tensor-core launches)
    (64, 64, 256, 1, True, 56, 2, "bf16", 2),     # layer1.0: conv1 + ONE fused launch (conv2 + conv3 + folded downsample)
    (256, 64, 256, 1, False, 56, 2, "bf16", 2),   # layer1.1: conv1 + fused tail with identity shortcut
    (256, 128, 512, 2, True, 56, 2, "bf16", 4),   # layer2.0: downsample, conv1, conv2 (stride), conv3 + shortcut
    (1024, 256, 1024, 1, False, 14, 3, "bf16", 3),
    (64, 64, 256, 1, True, 56, 1, "tf32", 4),     # TF32: layer by layer
    (2048, 512, 2048, 1, False, 7, 2, "tf32", 3),
], ids=lambda c: "-".join(map(str, c[:8])))
def test_bottleneck_block_object(oracle_lib, case):
    from resnet_c_b200 import engine
    cin, mid, cout, stride, ds, H, B, dtype, nlaunch = case
    g = torch.Generator().manual_seed(cin + mid + H)
    convs = _bottleneck(g, cin, mid, cout, stride, ds)
    x = torch.randn(B, cin, H, H, generator=g)
    want = _oracle_bottleneck(oracle_lib, convs, x)
    blk = engine.Block("bottleneck", _cuda(convs), dtype)
    got = blk.forward(x.cuda())
    again = blk.forward(x.cuda())          # cached plan, same result
    torch.cuda.synchronize()
    assert blk.num_launches(B, H, H) == nlaunch
    assert torch.equal(got, again)
    e = rel_err(got.cpu().numpy().reshape(B, -1), np.asarray(want).reshape(B, -1))
    assert e < TOL[dtype], f"{case}: rel err {e:.3e}"
    blk.close()


@pytest.mark.parametrize("case", [
    (64, 64, 1, False, 56, 2, "tf32"), (64, 128, 2, True, 56, 2, "tf32"), (256, 512, 2, True, 14, 3, "bf16"),
    (512, 512, 1, False, 7, 3, "bf16"),
], ids=lambda c: "-".join(map(str, c)))
def test_basic_block_object(oracle_lib, case):
    from resnet_c_b200 import engine
    cin, cout, stride, ds, H, B, dtype = case
    g = torch.Generator().manual_seed(cin + cout + H)
    convs = [_conv(g, cin, cout, 3, stride, 1), _conv(g, cout, cout, 3, 1, 1)]
    if ds:
        convs.append(_conv(g, cin, cout, 1, stride, 0))
    x = torch.randn(B, cin, H, H, generator=g)
    want = _oracle_basic(oracle_lib, convs, x)
    blk = engine.Block("basic", _cuda(convs), dtype)
    got = blk.forward(x.cuda())
    torch.cuda.synchronize()
    e = rel_err(got.cpu().numpy().reshape(B, -1), np.asarray(want).reshape(B, -1))
    assert e < TOL[dtype], f"{case}: rel err {e:.3e}"
    blk.close()


def test_conv_block_equals_per_call_path():
    """The planned single conv is the same arithmetic as rnb_conv_bn_act_forward: bit-identical, shapes re-planned."""
    from resnet_c_b200 import engine
    g = torch.Generator().manual_seed(9)
    c = _conv(g, 128, 128, 3, 1, 1)
    blk = engine.Block("conv", _cuda([c]), "bf16")
    for B, H in ((2, 28), (3, 14), (2, 28)):
        x = torch.randn(B, 128, H, H, generator=g).cuda()
        res = torch.randn(B, 128, H, H, generator=g).cuda()
        got = blk.forward(x, res, relu=True)
        want = engine.conv_bn_act_forward(x, c["w"].cuda(), tuple(t.cuda() for t in c["bn"]), res, True, 1, 1, "bf16")
        assert torch.equal(got, want)
    blk.close()


def _write_block_files(d: Path, prefix: str, convs, names):
    wd = d / "weights_bin"
    wd.mkdir(parents=True, exist_ok=True)
    for c, (cn, bn) in zip(convs, names):
        c["w"].numpy().astype(np.float32).tofile(wd / f"{prefix}{cn}.weight")
        for suffix, t in zip(("weight", "bias", "running_mean", "running_var"), c["bn"]):
            t.numpy().astype(np.float32).tofile(wd / f"{prefix}{bn}.{suffix}")


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_cxx_bottleneck_module_class(oracle_lib, tmp_path, dtype):
    """Bottleneck::loadWeightToCuda + forward of cuda/nn.cuh, as a C++ user calls them (build/block_check)."""
    exe = ROOT / "build" / "block_check"
    if not exe.exists():
        pytest.skip("build/block_check not built")
    g = torch.Generator().manual_seed(5)
    convs = _bottleneck(g, 64, 64, 256, 1, True)
    _write_block_files(tmp_path, "layer1.0.", convs,
                       [("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"), ("downsample.0", "downsample.1")])
    x = torch.randn(4, 64, 56, 56, generator=g)
    x.numpy().tofile(tmp_path / "x.bin")
    r = subprocess.run([str(exe), "bottleneck", "layer1.0.", "64", "64", "256", "1", "1", dtype, "4", "56", "x.bin",
                        "y.bin"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.fromfile(tmp_path / "y.bin", dtype=np.float32).reshape(4, -1)
    want = np.asarray(_oracle_bottleneck(oracle_lib, convs, x)).reshape(4, -1)
    assert rel_err(got, want) < TOL[dtype]
    ms = float(r.stdout.split("forward_ms")[1].split()[0])
    assert ms < 50, f"module forward took {ms} ms: weights are being re-folded per call?"


def test_cxx_conv2d_module_tensor_core_switch(oracle_lib, tmp_path):
    """Conv2d::forward: FP32 CUDA-core kernel by default (bit-exact vs the oracle), tensor cores with
    RNB_MODULE_TC=tf32 (<= 1e-3)."""
    exe = ROOT / "build" / "block_check"
    if not exe.exists():
        pytest.skip("build/block_check not built")
    g = torch.Generator().manual_seed(6)
    w = torch.randn(128, 64, 3, 3, generator=g) * 0.05
    (tmp_path / "weights_bin").mkdir()
    w.numpy().tofile(tmp_path / "weights_bin" / "c.weight")
    x = torch.randn(2, 64, 28, 28, generator=g)
    x.numpy().tofile(tmp_path / "x.bin")
    want = np.asarray(oracle_lib.conv2d(x, w, 1, 1)).reshape(2, -1)
    for mode, tol in ((None, 0.0), ("tf32", 1e-3)):
        env = dict(os.environ)
        env.pop("RNB_MODULE_TC", None)
        if mode:
            env["RNB_MODULE_TC"] = mode
        r = subprocess.run([str(exe), "conv", "c", "64", "3", "128", "1", "1", "bf16", "2", "28", "x.bin", "y.bin"],
                           cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        got = np.fromfile(tmp_path / "y.bin", dtype=np.float32).reshape(2, -1)
        if tol == 0.0:
            assert np.array_equal(got, want)
        else:
            e = rel_err(got, want)
            assert 0 < e < tol, e


def test_block_and_preprocess_argument_errors():
    """Bad requests come back as error codes with a message (nothing aborts, nothing is launched)."""
    import ctypes as C
    from resnet_c_b200 import _lib, engine
    from resnet_c_b200._lib import RnbError
    g = torch.Generator().manual_seed(1)
    with pytest.raises(RnbError, match="unsupported conv"):      # 3 input channels: not a tensor-core shape
        engine.Block("conv", [dict(w=torch.randn(64, 3, 3, 3, generator=g).cuda(), bn=None, stride=1, pad=1)], "bf16")
    with pytest.raises(RnbError, match="Bottleneck 3"):          # wrong conv count for the kind
        engine.Block("bottleneck", [dict(w=torch.randn(64, 64, 1, 1, generator=g).cuda(), bn=None)], "bf16")
    blk = engine.Block("conv", [dict(w=torch.randn(64, 64, 1, 1, generator=g).cuda(), bn=None)], "bf16")
    with pytest.raises(RnbError):                                # NULL output
        _lib.check(_lib.lib().rnb_block_forward(blk._h, C.c_void_p(1), 1, 8, 8, None, 1, None, None))
    blk.close()
    small = torch.zeros(1, 100, 300, 3, dtype=torch.uint8).cuda()
    with pytest.raises(RnbError, match="bad argument"):
        engine.resize_crop_u8(small, resize=200, crop=224)       # crop larger than the resize target
    out = engine.resize_crop_u8(small)                           # 100 x 300 -> 256 x 768 -> centre 224 x 224: fine
    assert out.shape == (1, 224, 224, 3) and int(out.sum()) == 0
