"""-m gpu: the tcgen05 implicit-GEMM convolution (conv + folded BN + residual + ReLU in one launch)
against the plain-C oracle running the reference's unfused chain conv2d -> batchnorm -> add -> relu
(layerForward, cuda/inference/main.cu:138-163), on every distinct layer shape of ResNet-18/50/152
 at small batch.

Tolerance = BASELINE.json north_star, applied per layer: max|d| / max|y| <= 2e-2 (BF16 operands,
FP32 accumulate) and <= 1e-3 (TF32, round-to-nearest operands, FP32 accumulate)."""
import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = {"bf16": 2e-2, "tf32": 1e-3}

# (Cin, Cout, H, k, stride, pad, residual, relu, B)
BOTTLENECK_SHAPES = [
    (64, 64, 56, 1, 1, 0, False, True, 1),
    (64, 64, 56, 3, 1, 1, False, True, 1),
    (64, 256, 56, 1, 1, 0, True, True, 1),      # conv3 + identity shortcut
    (64, 256, 56, 1, 1, 0, False, False, 1),    # layer1.0 downsample
    (256, 64, 56, 1, 1, 0, False, True, 1),
    (256, 128, 56, 1, 1, 0, False, True, 1),
    (128, 128, 56, 3, 2, 1, False, True, 2),
    (128, 512, 28, 1, 1, 0, True, True, 2),
    (256, 512, 56, 1, 2, 0, False, False, 2),   # strided downsample
    (512, 128, 28, 1, 1, 0, False, True, 2),
    (128, 128, 28, 3, 1, 1, False, True, 2),
    (512, 256, 28, 1, 1, 0, False, True, 2),
    (256, 256, 28, 3, 2, 1, False, True, 3),
    (256, 1024, 14, 1, 1, 0, True, True, 3),
    (512, 1024, 28, 1, 2, 0, False, False, 3),
    (1024, 256, 14, 1, 1, 0, False, True, 3),
    (256, 256, 14, 3, 1, 1, False, True, 3),
    (1024, 512, 14, 1, 1, 0, False, True, 3),
    (512, 512, 14, 3, 2, 1, False, True, 3),
    (512, 2048, 7, 1, 1, 0, True, True, 3),
    (1024, 2048, 14, 1, 2, 0, False, False, 3),
    (2048, 512, 7, 1, 1, 0, False, True, 3),
    (512, 512, 7, 3, 1, 1, False, True, 3),
]
BASIC_SHAPES = [
    (64, 64, 56, 3, 1, 1, True, True, 1),       # BasicBlock conv2 + identity shortcut
    (64, 128, 56, 3, 2, 1, False, True, 2),
    (128, 128, 28, 3, 1, 1, True, True, 2),
    (64, 128, 56, 1, 2, 0, False, False, 2),
    (128, 256, 28, 3, 2, 1, False, True, 3),
    (256, 256, 14, 3, 1, 1, True, True, 3),
    (128, 256, 28, 1, 2, 0, False, False, 3),
    (256, 512, 14, 3, 2, 1, False, True, 3),
    (512, 512, 7, 3, 1, 1, True, True, 3),
    (256, 512, 14, 1, 2, 0, False, False, 3),
]


def _case(oracle_lib, Cin, Cout, H, k, stride, pad, residual, relu, B, dtype, seed):
    from resnet_c_b200 import engine
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5
    bn = (torch.rand(Cout, generator=g) * 0.5 + 0.75, torch.randn(Cout, generator=g) * 0.1,
          torch.randn(Cout, generator=g) * 0.1, torch.rand(Cout, generator=g) * 0.5 + 0.75)
    OH = (2 * pad + H - k) // stride + 1
    res = torch.randn(B, Cout, OH, OH, generator=g) if residual else None
    y = oracle_lib.batchnorm2d(oracle_lib.conv2d(x, w, stride, pad), *bn)
    if residual:
        y = oracle_lib.add(y, res)
    if relu:
        y = oracle_lib.relu(y)
    got = engine.conv_bn_act_forward(x.cuda(), w.cuda(), tuple(t.cuda() for t in bn),
                                     None if res is None else res.cuda(), relu, stride, pad, dtype)
    return rel_err(got.cpu().numpy().reshape(B, -1), y.reshape(B, -1))


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
@pytest.mark.parametrize("shape", BOTTLENECK_SHAPES, ids=lambda s: "x".join(map(str, s[:6])))
def test_bottleneck_layer_shapes(oracle_lib, shape, dtype):
    e = _case(oracle_lib, *shape, dtype, seed=hash(shape) % 1000)
    assert e < TOL[dtype], f"{shape} {dtype}: rel err {e:.3e}"


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
@pytest.mark.parametrize("shape", BASIC_SHAPES, ids=lambda s: "x".join(map(str, s[:6])))
def test_basicblock_layer_shapes(oracle_lib, shape, dtype):
    e = _case(oracle_lib, *shape, dtype, seed=hash(shape) % 1000)
    assert e < TOL[dtype], f"{shape} {dtype}: rel err {e:.3e}"


def test_no_bn_no_relu_plain_conv(oracle_lib):
    from resnet_c_b200 import engine
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(2, 64, 10, 10, generator=g), torch.randn(128, 64, 3, 3, generator=g) * 0.05
    got = engine.conv_bn_act_forward(x.cuda(), w.cuda(), None, None, False, 1, 1, "tf32")
    assert rel_err(got.cpu().numpy().reshape(2, -1), oracle_lib.conv2d(x, w, 1, 1).reshape(2, -1)) < 1e-3


def test_linearity_at_full_size():
    """Size-independent property at the BASELINE batch (256 x 56 x 56 x 64 -> 256): with bias-free,
    ReLU-free conv, f(2x) == 2 f(x) exactly (power-of-two scaling commutes with every rounding), and
    f(0) == 0."""
    from resnet_c_b200 import engine
    g = torch.Generator().manual_seed(4)
    x = torch.randn(256, 64, 56, 56, generator=g).cuda()
    w = (torch.randn(256, 64, 1, 1, generator=g) * 0.1).cuda()
    y1 = engine.conv_bn_act_forward(x, w, None, None, False, 1, 0, "bf16")
    y2 = engine.conv_bn_act_forward(x * 2, w, None, None, False, 1, 0, "bf16")
    assert torch.equal(y2, y1 * 2)
    y0 = engine.conv_bn_act_forward(torch.zeros_like(x), w, None, None, False, 1, 0, "bf16")
    assert float(y0.abs().max()) == 0.0


def test_batch_rows_are_independent():
    """Tile boundaries must not leak between images: the first image of a batch of 5 equals the same
    image run alone, bit for bit (3x3 pad 1, 14x14: 196 pixels per image straddle the 128-row tiles)."""
    from resnet_c_b200 import engine
    g = torch.Generator().manual_seed(5)
    x = torch.randn(5, 128, 14, 14, generator=g).cuda()
    w = (torch.randn(128, 128, 3, 3, generator=g) * 0.03).cuda()
    full = engine.conv_bn_act_forward(x, w, None, None, True, 1, 1, "bf16")
    for i in (0, 2, 4):
        one = engine.conv_bn_act_forward(x[i:i + 1].contiguous(), w, None, None, True, 1, 1, "bf16")
        assert torch.equal(full[i:i + 1], one)


@pytest.mark.parametrize("shape", [
    # (Cin, Cout, H, k, stride, pad, B): activation tensors under 128 KiB, where the driver encodes im2col descriptors
    # with a flag the hardware mis-handles unless tensormap.cu's work-around is on (rnb_init's self-test decides)
    (64, 64, 8, 3, 1, 1, 1),      # 8 KiB (bf16)
    (512, 512, 7, 3, 1, 1, 1),    # layer4 conv2 at batch 1: 49 KiB
    (256, 256, 14, 3, 2, 1, 1),   # 98 KiB, stride 2
    (512, 128, 7, 1, 1, 0, 2),    # 1x1 through the im2col map
])
@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_im2col_on_tensors_below_128_kib(oracle_lib, shape, dtype):
    Cin, Cout, H, k, stride, pad, B = shape
    e = _case(oracle_lib, Cin, Cout, H, k, stride, pad, False, True, B, dtype, seed=11)
    assert e < TOL[dtype], f"{shape} {dtype}: rel err {e:.3e}"


@pytest.mark.parametrize("case", [
    # (Cin, Cout, H, k, stride, pad, residual, B, dtype, tile): pair tiles whose last wave is split into N halves
    (256, 256, 14, 3, 1, 1, False, 128, "bf16", 1256),   # 98 tiles on 74 pairs: 24 tiles as 48 half units
    (512, 512, 7, 1, 1, 0, True, 400, "bf16", 1256),      # 77 x 2 tiles, residual prefetch across the switch
    (256, 256, 14, 3, 1, 1, False, 8, "bf16", 1256),      # 7 tiles: all split (14 half units)
    (128, 128, 28, 3, 1, 1, True, 40, "tf32", 1128),      # tf32, BN = 128 (two 64-column sub-tiles)
])
def test_tail_split_is_bit_identical_to_whole_tiles(case, monkeypatch):
    """conv_igemm2's tail split (ConvGeom::split_from) changes the schedule, not the arithmetic: the launch
    with the last wave issued as N-half units equals the launch with whole tiles bit for bit."""
    from resnet_c_b200 import engine
    Cin, Cout, H, k, stride, pad, residual, B, dtype, tile = case
    g = torch.Generator().manual_seed(21)
    x = torch.randn(B, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).cuda()
    OH = (2 * pad + H - k) // stride + 1
    res = torch.randn(B, Cout, OH, OH, generator=g).cuda() if residual else None
    monkeypatch.setenv("RNB_FORCE_TILE", str(tile))
    monkeypatch.setenv("RNB_NO_SPLIT", "1")
    whole = engine.conv_bn_act_forward(x, w, None, res, True, stride, pad, dtype)
    monkeypatch.setenv("RNB_NO_SPLIT", "0")
    split = engine.conv_bn_act_forward(x, w, None, res, True, stride, pad, dtype)
    torch.cuda.synchronize()
    assert torch.equal(whole, split)
    assert float(whole.abs().max()) > 0


@pytest.mark.parametrize("case", [
    # (Cin, Cout, H, k, stride, pad, residual, B)
    (512, 2048, 7, 1, 1, 0, True, 16),     # layer4 conv3 + shortcut: the epilogue-bound shape the variant is for
    (128, 128, 14, 3, 1, 1, False, 5),
])
def test_sixteen_epilogue_warps_equal_eight(case, monkeypatch):
    """ConvCfg<..., EPI_WARPS = 16> (force code 20128) splits the same 32-column chunks over twice as many warps:
    bit-identical to the 8-warp kernel, BF16 and FP8."""
    from resnet_c_b200 import engine
    Cin, Cout, H, k, stride, pad, residual, B = case
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * k * k)) ** 0.5).cuda()
    OH = (2 * pad + H - k) // stride + 1
    res = torch.randn(B, Cout, OH, OH, generator=g).cuda() if residual else None
    outs = {}
    for tile in ("128", "20128"):
        monkeypatch.setenv("RNB_FORCE_TILE", tile)
        outs[tile] = (engine.conv_bn_act_forward(x, w, None, res, True, stride, pad, "bf16"),
                      engine.conv_fp8_forward(x, w, None, res, True, stride, pad, 2.0 ** -6, 2.0 ** -6, 2.0 ** -5))
    torch.cuda.synchronize()
    assert torch.equal(outs["128"][0], outs["20128"][0])
    assert torch.equal(outs["128"][1], outs["20128"][1])
    assert float(outs["128"][0].abs().max()) > 0 and float(outs["128"][1].abs().max()) > 0


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_standalone_conv_selftest_binary(dtype):
    """build/conv_selftest (tools/conv_selftest.cu): the conv kernel against a host loop with no Python in between —
    kept green so that a C++-only bring-up of the kernel stays possible."""
    import subprocess
    from conftest import ROOT
    exe = ROOT / "build" / "conv_selftest"
    if not exe.exists():
        pytest.skip("build/conv_selftest not built")
    r = subprocess.run([str(exe), dtype], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-500:])


@pytest.mark.parametrize("case", [
    # (H, stride, residual, B): 3x3, 128 -> 128 (layer2 conv2 of the Bottleneck nets)
    (28, 1, False, 40),    # 123 pair tiles on 74 pairs
    (56, 2, False, 9),     # stride 2 (layer2.0), ragged last tile
    (28, 1, True, 3),      # with a residual through the single staging buffer
])
def test_resident_weight_kernel_equals_streaming_kernel(case, monkeypatch):
    """Conv2Cfg<128, 2, 3, 1, 2, 18> (force code 31128): the weight matrix of a 128-channel 3x3 conv stays in shared
    memory for the whole launch, only the im2col operand streams. Same K order, same epilogue: bit-identical to the
    streaming CTA-pair kernel (1128)."""
    from resnet_c_b200 import engine
    H, stride, residual, B = case
    g = torch.Generator().manual_seed(41)
    x = torch.randn(B, 128, H, H, generator=g).cuda()
    w = (torch.randn(128, 128, 3, 3, generator=g) * 0.04).cuda()
    bn = tuple(t.cuda() for t in (torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g) * 0.1,
                                  torch.randn(128, generator=g) * 0.1, torch.rand(128, generator=g) + 0.5))
    OH = (2 + H - 3) // stride + 1
    res = torch.randn(B, 128, OH, OH, generator=g).cuda() if residual else None
    outs = {}
    for tile in ("1128", "31128"):
        monkeypatch.setenv("RNB_FORCE_TILE", tile)
        outs[tile] = engine.conv_bn_act_forward(x, w, bn, res, True, stride, 1, "bf16")
    torch.cuda.synchronize()
    assert torch.equal(outs["1128"], outs["31128"])
    assert float(outs["1128"].abs().max()) > 0


@pytest.mark.parametrize("case", [
    # (C, H, B, tile codes): every tile family a 3x3 layer admits computes the same bits
    (128, 28, 40, ("128", "1128", "11128", "12128", "31128")),
    (256, 14, 64, ("128", "1128", "1256", "11256", "12256")),
])
def test_tile_families_are_bit_identical(case, monkeypatch):
    """Single-CTA / CTA-pair tiles, deep (more ring stages, two staging buffers), deepest (12128 / 12256: eight / six
    stages, ONE staging buffer) and resident-weight variants differ in scheduling only."""
    from resnet_c_b200 import engine
    C, H, B, tiles = case
    g = torch.Generator().manual_seed(43)
    x = torch.randn(B, C, H, H, generator=g).cuda()
    w = (torch.randn(C, C, 3, 3, generator=g) * 0.03).cuda()
    outs = []
    for tile in tiles:
        monkeypatch.setenv("RNB_FORCE_TILE", tile)
        outs.append(engine.conv_bn_act_forward(x, w, None, None, True, 1, 1, "bf16"))
    torch.cuda.synchronize()
    for t, o in zip(tiles[1:], outs[1:]):
        assert torch.equal(outs[0], o), f"tile code {t} differs from {tiles[0]}"
