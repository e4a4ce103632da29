"""-m gpu: whole-network parity through rnb_model_* (include/rnb.h) against the committed goldens
and the oracle, plus size-independent properties at the BASELINE batch sizes.

Bars (BASELINE.json north_star): logits max|d|/max|y| <= 1e-3 (TF32) / 2e-2 (BF16); top-1 bit-exact.
Top-1 is asserted wherever the oracle's own top-1/top-2 margin (FP64, recorded in the golden)
exceeds twice the logit tolerance of the path — under random init some images are near-ties
(SURVEY.md section 7), and a "mismatch" inside the tolerance band is not an error of the kernel."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = {"bf16": 2e-2, "tf32": 1e-3}


def _model(arch, rbn, dtype, max_batch, chunk=0):
    from resnet_c_b200 import engine, weights
    return engine.ResNet(arch, weights.cached_weights_dir(arch, 0, rbn), dtype=dtype, max_batch=max_batch,
                         chunk=chunk)


def _inputs(tag, batch, jpeg_tensor):
    from resnet_c_b200 import weights
    return jpeg_tensor.repeat(batch, 1, 1, 1) if tag == "jpeg" else weights.synthetic_images(batch)


@pytest.mark.parametrize("name,arch,rbn,tag,batch,dtype", [
    ("resnet18_default_jpeg_b1", "resnet18", False, "jpeg", 1, "tf32"),   # BASELINE configs[0]
    ("resnet18_default_jpeg_b1", "resnet18", False, "jpeg", 1, "bf16"),
    ("resnet18_rbn_jpeg_b1", "resnet18", True, "jpeg", 1, "tf32"),
    ("resnet18_rbn_synth_b4", "resnet18", True, "synth", 4, "tf32"),
    ("resnet18_rbn_synth_b4", "resnet18", True, "synth", 4, "bf16"),
    ("resnet50_rbn_synth_b4", "resnet50", True, "synth", 4, "bf16"),
    ("resnet50_rbn_synth_b4", "resnet50", True, "synth", 4, "tf32"),
    ("resnet50_default_synth_b2", "resnet50", False, "synth", 2, "bf16"),
    ("resnet152_rbn_synth_b2", "resnet152", True, "synth", 2, "bf16"),
    ("resnet152_default_jpeg_b1", "resnet152", False, "jpeg", 1, "bf16"),  # the reference's own run
    ("resnet152_default_jpeg_b1", "resnet152", False, "jpeg", 1, "tf32"),
])
def test_logits_and_top1_against_goldens(name, arch, rbn, tag, batch, dtype, jpeg_tensor):
    g = load_golden(name)
    model = _model(arch, rbn, dtype, batch)
    logits, top1 = model.forward(_inputs(tag, batch, jpeg_tensor).cuda())
    torch.cuda.synchronize()
    got = logits.cpu().numpy()
    e = rel_err(got, g["logits_fp32"])
    assert e < TOL[dtype], f"{name} {dtype}: logits rel err {e:.3e}"
    decided = g["margin_rel"] > 2 * TOL[dtype]
    np.testing.assert_array_equal(top1.cpu().numpy()[decided], g["top1"][decided])
    # the GPU arg-max itself is exact on the GPU's own logits (first maximum wins)
    np.testing.assert_array_equal(top1.cpu().numpy(), got.argmax(1))
    model.close()


def test_reference_image_top1_is_bit_exact(jpeg_tensor):
    """BASELINE configs[0] + the reference's own scenario: margins are 8% / 14% of max|y|, far above
    any tolerance, so the index must match on every path."""
    for name, arch, dtype in [("resnet18_default_jpeg_b1", "resnet18", "tf32"),
                              ("resnet18_default_jpeg_b1", "resnet18", "bf16"),
                              ("resnet152_default_jpeg_b1", "resnet152", "bf16")]:
        g = load_golden(name)
        model = _model(arch, False, dtype, 1)
        _, top1 = model.forward(jpeg_tensor.cuda())
        assert top1.cpu().tolist() == g["top1"].tolist()
        model.close()


def test_intermediate_activations_resnet50(monkeypatch):
    """Layer-by-layer against the oracle's taps: localises an error to a block."""
    from oracle import torch_model
    from resnet_c_b200 import weights
    monkeypatch.setenv("RNB_KEEP_ACTIVATIONS", "1")
    arch, batch = "resnet50", 2
    sd = weights.make_state_dict(arch, 0, randomize_bn=True)
    x = weights.synthetic_images(batch)
    taps = {}
    torch_model.run(arch, sd, x, taps=taps)
    model = _model(arch, True, "bf16", batch)
    model.forward(x.cuda())
    for name, ref in taps.items():
        if name == "stem":
            continue  # the BF16 path fuses conv1+bn+relu+maxpool: the 112x112 map never exists
        got = model.activation(name).cpu().numpy().reshape(batch, -1)
        e = rel_err(got, ref.numpy().reshape(batch, -1))
        assert e < 2e-2, f"{name}: rel err {e:.3e}"
    model.close()


def test_forward_host_equals_device_path():
    from resnet_c_b200 import weights
    model = _model("resnet18", True, "bf16", 8, chunk=4)
    x = weights.synthetic_images(8)
    logits, top1 = model.forward(x.cuda())
    lh, th = model.forward_host(x.pin_memory())
    assert torch.equal(lh, logits.cpu()) and torch.equal(th, top1.cpu())
    lp, tp = model.forward_host(x)  # pageable host memory is accepted too
    assert torch.equal(lp, logits.cpu()) and torch.equal(tp, top1.cpu())
    model.close()


def test_pipelined_host_path_equals_device_path():
    """submit_host / wait_host with two batches in flight must return each batch's own results."""
    from resnet_c_b200 import weights
    model = _model("resnet18", True, "bf16", 16)
    xs = [weights.synthetic_images(16, seed=s).pin_memory() for s in (1, 2, 3)]
    want = []
    for x in xs:
        l, t = model.forward(x.cuda())
        want.append((l.cpu(), t.cpu()))
    outs = [(torch.empty(16, model.num_classes).pin_memory(), torch.empty(16, dtype=torch.int32).pin_memory())
            for _ in xs]
    model.submit_host(0, xs[0], *outs[0])
    model.submit_host(1, xs[1], *outs[1])
    model.wait_host(0)
    model.submit_host(0, xs[2], *outs[2])
    model.wait_host(1)
    model.wait_host(0)
    for (l, t), (wl, wt) in zip(outs, want):
        assert torch.equal(l, wl) and torch.equal(t, wt)
    model.close()


@pytest.mark.parametrize("arch,dtype", [("resnet50", "bf16"), ("resnet18", "tf32")])
def test_full_batch_properties(arch, dtype):
    """BASELINE configs[1]/[2] sizes (batch 256): the oracle cannot run these in seconds, so check
    properties that do not depend on size — every image's result is independent of its position in
    the batch, of the batch size and of the chunking, bit for bit; and it matches the oracle on a
    sampled subset within tolerance."""
    from oracle import torch_model
    from resnet_c_b200 import weights
    B = 256
    sd = weights.make_state_dict(arch, 0, randomize_bn=True)
    x = weights.synthetic_images(B)
    model = _model(arch, True, dtype, B)
    logits, top1 = model.forward(x.cuda())
    torch.cuda.synchronize()
    # permutation equivariance
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    lp, tp = model.forward(x[perm].cuda())
    assert torch.equal(lp.cpu(), logits.cpu()[perm]) and torch.equal(tp.cpu(), top1.cpu()[perm])
    # batch-size and chunk independence
    small = _model(arch, True, dtype, B, chunk=32)
    ls, ts = small.forward(x.cuda())
    assert torch.equal(ls.cpu(), logits.cpu()) and torch.equal(ts.cpu(), top1.cpu())
    l4, _ = small.forward(x[100:104].cuda())
    assert torch.equal(l4.cpu(), logits.cpu()[100:104])
    # sampled oracle comparison
    idx = [0, 37, 128, 255]
    ref = torch_model.run(arch, sd, x[idx])
    assert rel_err(logits.cpu().numpy()[idx], ref.numpy()) < TOL[dtype]
    np.testing.assert_array_equal(top1.cpu().numpy(), logits.cpu().numpy().argmax(1))
    model.close()
    small.close()


@pytest.mark.parametrize("arch,dtype", [("resnet50", "bf16"), ("resnet18", "bf16"), ("resnet18", "tf32")])
def test_uint8_input_path_is_bit_identical_to_float_input(arch, dtype, golden_dir):
    """rnb_model_forward_u8 (decoded uint8 HWC in, /255 + mean/std fused into the stem pre-pass — or the
    generic FP32 conversion kernel on the TF32 path) against rnb_model_forward fed with the oracle's
    normalised tensor: identical logits; and through the pipelined host path."""
    from oracle import preprocess
    from resnet_c_b200 import weights
    B = 5
    u8 = torch.cat([weights.load_u8_image_bin(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin"),
                    weights.synthetic_images_u8(B - 1)])
    model = _model(arch, True, dtype, B)
    want, want_top1 = model.forward(preprocess.normalize_u8(u8).cuda())
    got, got_top1 = model.forward_u8(u8.cuda())
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got_top1, want_top1)
    lh, th = torch.empty(B, model.num_classes).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory()
    model.submit_host_u8(0, u8.pin_memory(), lh, th)
    model.wait_host(0)
    assert torch.equal(lh, want.cpu()) and torch.equal(th, want_top1.cpu())
    # other constants are honoured (and cached graphs dropped)
    model.set_normalization((0.5, 0.5, 0.5), (0.25, 0.25, 0.25))
    want2, _ = model.forward(preprocess.normalize_u8(u8, (0.5, 0.5, 0.5), (0.25, 0.25, 0.25)).cuda())
    got2, _ = model.forward_u8(u8.cuda())
    assert torch.equal(got2, want2) and not torch.equal(got2, got)
    model.close()


@pytest.mark.parametrize("arch,dtype,B", [("resnet50", "bf16", 37), ("resnet18", "bf16", 70), ("resnet50", "fp8", 40)])
def test_host_packed_input_is_bit_identical_to_fp32_input(arch, dtype, B):
    """The host paths' packed form (csrc/host_pack.cpp: the host cores round the FP32 image to BF16, half the bytes
    cross PCIe, the stem loads BF16 NCHW — stem_tc_kernel<2>) against the plain FP32 form: identical logits and top-1
    from rnb_model_forward_bf16 on a device tensor, from forward_host and from both slots of submit_host, with
    pinned and pageable input and a batch that is not a multiple of the 16-image upload pieces."""
    from resnet_c_b200 import weights
    x = weights.synthetic_images(B, seed=7)
    x[0, 0, 0, :4] = torch.tensor([-0.0, 1e-40, 1.00390625, -2.51171875])   # signed zero, a denormal, two ties
    model = _model(arch, True, dtype, B)
    want, want_top1 = model.forward(x.cuda())
    got, got_top1 = model.forward_bf16(x.cuda().to(torch.bfloat16))
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got_top1, want_top1)
    want, want_top1 = want.cpu(), want_top1.cpu()
    for mode in (1, 0.5, 0):
        # all images through the host cores / a mixed batch (the leading 16 or 32 images as BF16, the rest as FP32 in
        # ONE stem launch) / plain FP32 copies
        if mode == 0.5:
            model.set_host_pack_fraction(0.5)
        else:
            model.set_host_pack(mode)
        for xin in (x.pin_memory(), x.clone()):
            lh, th = model.forward_host(xin)
            assert torch.equal(lh, want) and torch.equal(th, want_top1), (mode, xin.is_pinned())
            outs = [(torch.empty(B, model.num_classes).pin_memory(), torch.empty(B, dtype=torch.int32).pin_memory())
                    for _ in range(2)]
            for slot in (0, 1):
                model.submit_host(slot, xin, *outs[slot])
            for slot in (0, 1):
                model.wait_host(slot)
                assert torch.equal(outs[slot][0], want) and torch.equal(outs[slot][1], want_top1), (mode, slot)
        info = model.host_pack()
        assert info["choice"] == (1 if mode else 0)
        assert info["fraction"] == (mode if mode != 0.5 else {37: 16 / 37, 70: 32 / 70, 40: 16 / 40}[B])
    # left to itself the model times both forms on the first host call and keeps one of them
    model.set_host_pack(-1)
    lh, th = model.forward_host(x.pin_memory())
    assert torch.equal(lh, want)
    info = model.host_pack()
    assert info["choice"] in (0, 1) and info["convert_gbps"] > 0 and info["h2d_f32_gbps"] > 0
    model.close()


def test_host_pack_is_refused_where_the_stem_takes_no_bf16_input():
    from resnet_c_b200 import weights
    from resnet_c_b200._lib import RnbError
    model = _model("resnet18", True, "tf32", 2)
    x = weights.synthetic_images(2)
    with pytest.raises(RnbError):
        model.set_host_pack(1)
    with pytest.raises(RnbError):
        model.forward_bf16(x.cuda().to(torch.bfloat16))
    with pytest.raises(RnbError):
        model.set_host_pack(2)
    want, _ = model.forward(x.cuda())
    lh, _ = model.forward_host(x.pin_memory())     # the plain path is what runs
    assert torch.equal(lh, want.cpu()) and model.host_pack()["choice"] in (-1, 0)
    model.close()


def test_uint8_reference_image_top1(golden_dir):
    """The reference's own scenario from the decoded JPEG: ResNet-152 -> 176, ResNet-18 -> 238."""
    from resnet_c_b200 import weights
    u8 = weights.load_u8_image_bin(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin").cuda()
    for name, arch in [("resnet18_default_jpeg_b1", "resnet18"), ("resnet152_default_jpeg_b1", "resnet152")]:
        g = load_golden(name)
        model = _model(arch, False, "bf16", 1)
        _, top1 = model.forward_u8(u8)
        assert top1.cpu().tolist() == g["top1"].tolist()
        model.close()


@pytest.mark.parametrize("arch,dtype", [("resnet50", "bf16"), ("resnet18", "tf32")])
def test_packed_weight_blob_round_trip(arch, dtype, tmp_path):
    """rnb_model_save_packed -> rnb_model_create_packed: bit-identical logits, flops and class count; a corrupt
    byte, a truncated file and a wrong magic are refused."""
    from resnet_c_b200 import engine, weights
    from resnet_c_b200._lib import RnbError
    B = 4
    x = weights.synthetic_images(B).cuda()
    model = _model(arch, True, dtype, B)
    want, want_top1 = model.forward(x)
    torch.cuda.synchronize()
    blob = tmp_path / f"{arch}_{dtype}.rnbw"
    model.save_packed(blob)
    packed = engine.ResNet.from_packed(blob, max_batch=B)
    assert packed.num_classes == model.num_classes and packed.flops_per_image == model.flops_per_image
    got, got_top1 = packed.forward(x)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(got_top1, want_top1)
    packed.close()
    model.close()
    raw = bytearray(blob.read_bytes())
    bad = tmp_path / "bad.rnbw"
    corrupt = bytearray(raw)
    corrupt[len(corrupt) // 2] ^= 0x40
    bad.write_bytes(corrupt)
    with pytest.raises(RnbError, match="checksum"):
        engine.ResNet.from_packed(bad, max_batch=B)
    bad.write_bytes(raw[:len(raw) - 1000])
    with pytest.raises(RnbError, match="size"):
        engine.ResNet.from_packed(bad, max_batch=B)
    bad.write_bytes(b"NOTRNBWT" + bytes(raw[8:]))
    with pytest.raises(RnbError, match="magic"):
        engine.ResNet.from_packed(bad, max_batch=B)


def _run_with_env(monkeypatch, env, arch, batch, names):
    from resnet_c_b200 import weights
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    monkeypatch.setenv("RNB_KEEP_ACTIVATIONS", "1")
    model = _model(arch, True, "bf16", batch)
    logits, top1 = model.forward(weights.synthetic_images(batch).cuda())
    torch.cuda.synchronize()
    acts = {n: model.activation(n).clone() for n in names}
    launches = model.launches_per_forward(batch)
    model.close()
    return logits.clone(), top1.clone(), acts, launches


@pytest.mark.parametrize("arch,batch", [("resnet50", 3), ("resnet50", 37), ("resnet152", 2)])
def test_fused_bottleneck_tail_equals_layer_by_layer(monkeypatch, arch, batch):
    """csrc/bneck_l1.cuh (layer1: conv2 + conv3 + residual (+ next conv1) in one launch) and
    csrc/bneck_c3n1.cuh (layer2: conv3 + residual + next conv1, active with RNB_FUSE_NEXT=1) round at the same
    points as the separate kernels, so RNB_FUSE=1 must be BIT-IDENTICAL to RNB_FUSE=0; RNB_FUSE=2 also
    folds the downsample conv into the conv3 accumulator (one BF16 rounding fewer on the shortcut), so
    it is compared within the BF16 bar. Odd batch sizes leave CTA pairs with unequal tile counts."""
    names = ("layer1.0", "layer1.1", "layer1.2", "layer2.0", "layer2.1", "layer2.2", "layer2.3", "layer3.0")
    base = _run_with_env(monkeypatch, {"RNB_FUSE": "0"}, arch, batch, names)
    # RNB_C3N1_AUTO=0: the conv3 + next-conv1 fusion wherever the shape allows (the default times fused against plain
    # per shape and batch size and may keep the plain launches at these small batches: this test is about the kernels)
    monkeypatch.setenv("RNB_C3N1_AUTO", "0")
    for nxt in ("0", "1"):
        got = _run_with_env(monkeypatch, {"RNB_FUSE": "1", "RNB_FUSE_NEXT": nxt}, arch, batch, names)
        assert got[3] < base[3]
        for n in names:
            assert torch.equal(got[2][n], base[2][n]), f"{n} differs (fuse=1 next={nxt})"
        assert torch.equal(got[0], base[0]) and torch.equal(got[1], base[1])
        got = _run_with_env(monkeypatch, {"RNB_FUSE": "2", "RNB_FUSE_NEXT": nxt}, arch, batch, names)
        for n in names:
            e = rel_err(got[2][n].cpu().numpy().reshape(batch, -1), base[2][n].cpu().numpy().reshape(batch, -1))
            assert e < 2e-2, f"{n}: rel err {e:.3e} (fuse=2 next={nxt})"
        assert rel_err(got[0].cpu().numpy(), base[0].cpu().numpy()) < 2e-2
    # opt-in variant: the last layer1 block also computes the next LAYER's conv1 (BneckCfg<false, 128>)
    got = _run_with_env(monkeypatch, {"RNB_FUSE": "1", "RNB_FUSE_NEXT": "1", "RNB_L1L2": "1"}, arch, batch, names)
    assert got[3] == base[3] - 14 if arch == "resnet50" else got[3] < base[3]
    for n in names:
        assert torch.equal(got[2][n], base[2][n]), f"{n} differs (RNB_L1L2=1)"
    assert torch.equal(got[0], base[0]) and torch.equal(got[1], base[1])
    monkeypatch.delenv("RNB_L1L2")


@pytest.mark.parametrize("arch,batch", [("resnet18", 3), ("resnet18", 37), ("resnet34", 2)])
def test_tf32_halo_pair_kernel_equals_generic_kernel(monkeypatch, arch, batch):
    """csrc/conv3x3_halo2.cuh (layer1 of the BasicBlock nets on the TF32 path: CTA pair, resident weights, halo
    ring) keeps the K order and rounding of the generic im2col kernel: block outputs and logits bit-identical."""
    from resnet_c_b200 import weights
    monkeypatch.setenv("RNB_AUTOTUNE", "0")
    monkeypatch.setenv("RNB_KEEP_ACTIVATIONS", "1")
    names = ("layer1.0", "layer1.1", "layer2.0")
    outs = []
    for no_halo in ("1", ""):
        if no_halo:
            monkeypatch.setenv("RNB_NO_HALO", no_halo)
        else:
            monkeypatch.delenv("RNB_NO_HALO", raising=False)
        model = _model(arch, True, "tf32", batch)
        logits, top1 = model.forward(weights.synthetic_images(batch).cuda())
        torch.cuda.synchronize()
        outs.append((logits.clone(), top1.clone(), {n: model.activation(n).clone() for n in names}))
        model.close()
    for n in names:
        assert torch.equal(outs[0][2][n], outs[1][2][n]), f"{n} differs"
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("arch,batch", [("resnet50", 128), ("resnet50", 101)])
def test_scheduling_switches_are_bit_identical(monkeypatch, arch, batch):
    """RNB_C3N1S_RINGS (how bneck_c3n1s_kernel splits its 224 KB of shared memory between weight rings and staging
    boxes) and RNB_BALANCE (persistent pair grids shrunk to equal tile counts per pair, conv_plan.cu::pair_count)
    change scheduling only: logits, top-1 and block outputs must be BIT-IDENTICAL. The batch is large enough for
    layer3 to run more pair tiles (98 / 78) than there are SM pairs, so rings wrap and the balanced grid differs."""
    names = ("layer2.3", "layer3.0", "layer3.5", "layer4.2")
    base = _run_with_env(monkeypatch, {"RNB_FUSE": "1", "RNB_FUSE_NEXT": "1"}, arch, batch, names)
    # round 2: RNB_C3N1_AUTO=0 (layer3 fused although its last wave is short: the default un-fuses it at these batch
    # sizes), RNB_C3N1_HYBRID=1 (fused launch over the whole waves, plain launches over the remaining rows) and
    # RNB_NO_SPLIT=1 (whole tiles in the last wave of the CTA-pair convs)
    for env in ({"RNB_C3N1S_RINGS": "1"}, {"RNB_C3N1S_RINGS": "2"}, {"RNB_C3N1S_RINGS": "3"}, {"RNB_BALANCE": "15"},
                {"RNB_C3N1_AUTO": "0"}, {"RNB_C3N1_AUTO": "0", "RNB_C3N1_HYBRID": "1"}, {"RNB_NO_SPLIT": "1"},
                {"RNB_C3N1_AUTO": "0", "RNB_NO_SPLIT": "1"}):
        got = _run_with_env(monkeypatch, {"RNB_FUSE": "1", "RNB_FUSE_NEXT": "1", **env}, arch, batch, names)
        for n in names:
            assert torch.equal(got[2][n], base[2][n]), f"{n} differs ({env})"
        assert torch.equal(got[0], base[0]) and torch.equal(got[1], base[1]), env
        for k in env:
            monkeypatch.delenv(k)


def test_launch_accounting(monkeypatch):
    monkeypatch.setenv("RNB_FUSE", "0")
    model = _model("resnet50", True, "bf16", 64, chunk=16)
    # fused stem (conv + BN + ReLU + max-pool straight from the FP32 image: ONE launch) + 52 tensor-core convs +
    # avgpool + fc + argmax per chunk
    assert model.launches_per_forward(16) == 56
    assert model.launches_per_forward(64) == 4 * 56
    model.close()
    monkeypatch.setenv("RNB_FUSE", "2")
    monkeypatch.setenv("RNB_C3N1_AUTO", "0")   # fused wherever the shape allows (the default times it against plain)
    model = _model("resnet50", True, "bf16", 64, chunk=16)
    # layer1: downsample + conv2 + conv3 of block 0 and conv1 + conv2 + conv3 of blocks 1, 2 -> 3 launches;
    # layer2: conv3 of blocks 0..2 absorbs conv1 of blocks 1..3 (3 launches fewer); layer3: blocks 0..4 (5 fewer)
    assert model.launches_per_forward(16) == 42
    assert model.flops_per_image == pytest.approx(8_178_368_512, rel=1e-9)
    model.close()
    m18 = _model("resnet18", True, "tf32", 4)
    assert m18.flops_per_image == pytest.approx(3_628_146_688, rel=1e-9)
    m18.close()


@pytest.mark.parametrize("arch,batch,dtype,nsample", [
    ("resnet50", 256, "bf16", 8),    # BASELINE configs[2] as bench.py runs it: seed-0 DEFAULT-init weights
    ("resnet152", 128, "bf16", 4),   # configs[4]: 1024 images over 8 GPUs = 128 per GPU
    ("resnet18", 256, "tf32", 8),    # configs[1]
])
def test_stated_config_sizes_against_committed_goldens(arch, batch, dtype, nsample):
    """The configs at their STATED batch sizes (other tile families, ragged pair grids and wave counts than the
    B <= 4 golden cases) on the weights bench.py really uses, against the committed oracle sample
    (tests/golden/make_golden_fullsize.py): logits within the bar on every sampled image, top-1 equal wherever the
    oracle's fp64 margin exceeds twice the tolerance — with the evidence printed: images decided, agreement, margin."""
    from resnet_c_b200 import weights
    g = load_golden(f"{arch}_default_synth_b{batch}_sample{nsample}")
    idx = g["index"]
    model = _model(arch, False, dtype, batch)
    logits, top1 = model.forward(weights.synthetic_images(batch).cuda())
    torch.cuda.synchronize()
    got, got_top1 = logits.cpu().numpy(), top1.cpu().numpy()
    e = rel_err(got[idx], g["logits_fp32"])
    assert e < TOL[dtype], f"{arch} B={batch} {dtype}: logits rel err {e:.3e}"
    decided = g["margin_rel"] > 2 * TOL[dtype]
    print(f"{arch} B={batch} {dtype}: rel err {e:.2e}; top-1 decided on {int(decided.sum())}/{len(idx)} sampled "
          f"images (min fp64 margin {g['margin_rel'].min():.2e}), agree {int((got_top1[idx] == g['top1'])[decided].sum())}")
    np.testing.assert_array_equal(got_top1[idx][decided], g["top1"][decided])
    np.testing.assert_array_equal(got_top1, got.argmax(1))
    model.close()


@pytest.mark.parametrize("fuse", ["0", "1", "2"])
def test_every_fusion_level_against_the_fp64_golden(monkeypatch, fuse):
    """RNB_FUSE=2 (the DEFAULT plan: layer1 downsample folded into conv3's accumulator, i.e. one BF16 rounding of the
    shortcut skipped) is not bit-identical to the layer-by-layer plan; pin each level's logits against the fp64 oracle
    golden explicitly. Skipping a rounding must not make the fp64 distance worse than the bar."""
    from resnet_c_b200 import weights
    monkeypatch.setenv("RNB_FUSE", fuse)
    g = load_golden("resnet50_rbn_synth_b4")
    model = _model("resnet50", True, "bf16", 4)
    logits, top1 = model.forward(weights.synthetic_images(4).cuda())
    torch.cuda.synchronize()
    e64 = rel_err(logits.cpu().numpy(), g["logits_fp64"])
    assert e64 < TOL["bf16"], f"RNB_FUSE={fuse}: rel err vs fp64 oracle {e64:.3e}"
    decided = g["margin_rel"] > 2 * TOL["bf16"]
    np.testing.assert_array_equal(top1.cpu().numpy()[decided], g["top1"][decided])
    model.close()


def test_forwards_on_different_streams_share_the_arena_safely():
    """All forwards of one model use ONE activation arena: calls enqueued on different streams (the caller's, the
    pipelined host path's, forward_host's) must be ordered by the library, not overlap on the arena."""
    from resnet_c_b200 import weights
    B = 32
    model = _model("resnet50", True, "bf16", B)
    xs = [weights.synthetic_images(B, seed=s) for s in (1, 2, 3, 4)]
    want = []
    for x in xs:
        l, t = model.forward(x.cuda())
        torch.cuda.synchronize()
        want.append((l.cpu(), t.cpu()))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    xd = [x.cuda() for x in xs]
    lh = torch.empty(B, model.num_classes).pin_memory()
    th = torch.empty(B, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(s1):
            a = model.forward(xd[0])
        with torch.cuda.stream(s2):
            b = model.forward(xd[1])
        model.submit_host(0, xs[2].pin_memory(), lh, th)     # third stream (pipe_compute), no wait in between
        with torch.cuda.stream(s1):
            c = model.forward(xd[3])
        model.wait_host(0)
        torch.cuda.synchronize()
        for got, (wl, wt) in zip((a, b, (lh, th), c), (want[0], want[1], want[2], want[3])):
            assert torch.equal(got[0].cpu(), wl) and torch.equal(got[1].cpu(), wt)
    model.close()


def test_fresh_output_buffers_every_call_and_warmup():
    """Callers that allocate new outputs per call (cuda/nn.cu ResNet::predict, engine.ResNet.forward) re-target a
    bounded set of graph executables instead of instantiating a graph per pointer set; rnb_model_warmup plans and
    captures ahead of the first forward."""
    from resnet_c_b200 import _lib, weights
    from resnet_c_b200._lib import RnbError, check
    B = 8
    model = _model("resnet18", True, "bf16", B)
    check(_lib.lib().rnb_model_warmup(model._h, B, 1))
    assert _lib.lib().rnb_model_device(model._h) == torch.cuda.current_device()
    x = weights.synthetic_images(B).cuda()
    first, first_top1 = model.forward(x)
    keep = []
    for i in range(12):  # 12 distinct (logits, top1) pointer pairs: more than the 4 executables kept per shape
        l, t = model.forward(x.clone() if i % 3 == 0 else x)
        keep.append((l, t))
    torch.cuda.synchronize()
    for l, t in keep:
        assert torch.equal(l, first) and torch.equal(t, first_top1)
    with pytest.raises(RnbError, match="RNB_KEEP_ACTIVATIONS"):
        model.activation("layer1.0")
    model.close()


@pytest.mark.parametrize("arch,batch", [("resnet50", 37), ("resnet18", 6)])
def test_two_lanes_equal_one_lane(monkeypatch, arch, batch):
    """RNB_LANES=2: the batch runs as two half batches on two streams (a second arena / plan / graph set sharing the
    weights, csrc/model.cu forward_two). Per-image results do not depend on the batch an image is in, so logits and
    top-1 are BIT-IDENTICAL to the one-lane forward — device, uint8 and host paths."""
    from resnet_c_b200 import weights
    x = weights.synthetic_images(batch)
    xu = weights.synthetic_images_u8(batch, seed=9)
    outs = {}
    for lanes in ("1", "2"):
        monkeypatch.setenv("RNB_LANES", lanes)
        m = _model(arch, True, "bf16", batch)
        lg, t1 = m.forward(x.cuda())
        lg2, _ = m.forward(x.cuda())                    # graph replay
        lu, tu = m.forward_u8(xu.cuda())
        hl, ht = m.forward_host(x.pin_memory())
        torch.cuda.synchronize()
        assert torch.equal(lg, lg2)
        outs[lanes] = (lg.cpu(), t1.cpu(), lu.cpu(), tu.cpu(), hl.clone(), ht.clone(), m.launches_per_forward(batch))
        m.close()
    for a, b in zip(outs["1"][:6], outs["2"][:6]):
        assert torch.equal(a, b)
    assert outs["2"][6] > outs["1"][6]
