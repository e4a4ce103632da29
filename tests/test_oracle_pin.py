"""Pins the oracle. The reference ships no golden vectors, so the pin is:
  * the reference's own PyTorch class run in the build container (golden committed, and re-run live
    when /root/reference is mounted),
  * torchvision (bit-identical model definition),
  * committed oracle logits for the BASELINE.json parity configs,
  * agreement between the two independent restatements (C kernels+graph vs PyTorch)."""
import numpy as np
import pytest
import torch

from conftest import REFERENCE, load_golden, rel_err
from oracle import torch_model
from resnet_c_b200 import weights


def _checksum(sd):
    return float(sum(v.double().sum().item() for v in sd.values() if v.dtype.is_floating_point))


@pytest.fixture(scope="module")
def sd152():
    return weights.make_state_dict("resnet152", 0)


def test_seeded_weights_are_reproducible(sd152):
    g = load_golden("ref_class_resnet152")
    assert _checksum(sd152) == pytest.approx(float(g["weights_checksum"]), rel=1e-12)


def test_torch_oracle_reproduces_reference_class_golden(sd152, jpeg_tensor):
    """Golden = the reference's own Resnet152 class (pytorch_inference.py:113-162) run on CPU."""
    torch.set_num_threads(1)
    g = load_golden("ref_class_resnet152")
    y = torch_model.run("resnet152", sd152, jpeg_tensor)
    assert int(y.argmax(1)) == int(g["top1"][0]) == 176
    # same model definition, same thread count as the generator: bit-identical in practice; allow the
    # last-ulp drift of a different oneDNN code path on another CPU
    assert rel_err(y.numpy(), g["logits"]) < 2e-6


@pytest.mark.skipif(not (REFERENCE / "pytorch_inference.py").exists(), reason="reference tree not mounted")
def test_torch_oracle_bit_identical_to_live_reference_class(sd152, jpeg_tensor):
    import importlib.util
    from conftest import GOLDEN
    spec = importlib.util.spec_from_file_location("make_golden", GOLDEN / "make_golden.py")
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    torch.set_num_threads(1)
    ns = mg.reference_model_classes()
    ref = ns["Resnet152"](1000)
    ref.load_state_dict(sd152, strict=True)
    ref.eval()
    with torch.no_grad():
        y_ref = ref(jpeg_tensor)
    y = torch_model.run("resnet152", sd152, jpeg_tensor)
    assert torch.equal(y, y_ref)


@pytest.mark.parametrize("arch", ["resnet18", "resnet34", "resnet50"])
def test_torch_oracle_bit_identical_to_torchvision(arch):
    import torchvision
    torch.set_num_threads(1)
    sd = weights.make_state_dict(arch, 0, randomize_bn=True)
    x = weights.synthetic_images(1)
    tv = getattr(torchvision.models, arch)(weights=None)
    tv.load_state_dict(sd)
    tv.eval()
    with torch.no_grad():
        y_tv = tv(x)
    assert torch.equal(torch_model.run(arch, sd, x), y_tv)


@pytest.mark.parametrize("name,arch,rbn,tag,batch", [
    ("resnet18_default_jpeg_b1", "resnet18", False, "jpeg", 1),
    ("resnet18_rbn_jpeg_b1", "resnet18", True, "jpeg", 1),
    ("resnet18_rbn_synth_b4", "resnet18", True, "synth", 4),
    ("resnet50_rbn_synth_b4", "resnet50", True, "synth", 4),
])
def test_torch_oracle_matches_committed_goldens(name, arch, rbn, tag, batch, jpeg_tensor):
    torch.set_num_threads(1)
    g = load_golden(name)
    sd = weights.make_state_dict(arch, 0, randomize_bn=rbn)
    assert _checksum(sd) == pytest.approx(float(g["weights_checksum"]), rel=1e-12)
    x = jpeg_tensor.repeat(batch, 1, 1, 1) if tag == "jpeg" else weights.synthetic_images(batch)
    y = torch_model.run(arch, sd, x)
    assert rel_err(y.numpy(), g["logits_fp32"]) < 2e-6
    assert rel_err(y.numpy(), g["logits_fp64"]) < 5e-6
    np.testing.assert_array_equal(y.argmax(1).numpy(), g["top1"])


@pytest.mark.parametrize("arch,rbn", [("resnet18", True), ("resnet50", True)])
def test_c_oracle_graph_agrees_with_torch_oracle(oracle_lib, arch, rbn):
    """Two independent restatements (ops.cu+main.cu in C vs pytorch_inference.py in torch)."""
    sd = weights.make_state_dict(arch, 0, randomize_bn=rbn)
    x = weights.synthetic_images(1)
    taps_c, taps_t = {}, {}
    y_c, top_c = oracle_lib.resnet_forward(arch, sd, x.numpy(), taps_c)
    y_t = torch_model.run(arch, sd, x, taps=taps_t)
    assert rel_err(y_c, y_t.numpy()) < 1e-5
    assert top_c.tolist() == y_t.argmax(1).tolist()
    for name in ("stem", "maxpool", "layer1.0", "layer4.1" if arch == "resnet18" else "layer4.2", "avgpool"):
        assert rel_err(taps_c[name].reshape(1, -1), taps_t[name].numpy().reshape(1, -1)) < 1e-5, name


def test_c_oracle_matches_golden_on_the_reference_image(oracle_lib, jpeg_tensor):
    """BASELINE.json configs[0]: ResNet-18 FP32, batch 1, the test image."""
    g = load_golden("resnet18_default_jpeg_b1")
    sd = weights.make_state_dict("resnet18", 0)
    y, top = oracle_lib.resnet_forward("resnet18", sd, jpeg_tensor.numpy())
    assert rel_err(y, g["logits_fp64"]) < 1e-5
    assert top.tolist() == g["top1"].tolist() == [238]


def test_u8_normalisation_reproduces_the_reference_preprocessing(jpeg_tensor, golden_dir):
    """oracle/preprocess.py::normalize_u8 on the committed uint8 crop of the reference's test JPEG must give,
    bit for bit, the tensor the reference's convert_imgs_to_bin.py preset produced (the committed .bin)."""
    from oracle import preprocess
    from resnet_c_b200 import weights
    u8 = weights.load_u8_image_bin(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin")
    assert u8.shape == (1, 224, 224, 3)
    assert torch.equal(preprocess.normalize_u8(u8), jpeg_tensor)


@pytest.mark.skipif(not REFERENCE.exists(), reason="reference tree not mounted (GPU box)")
def test_u8_crop_golden_matches_live_decode(golden_dir):
    from resnet_c_b200 import weights
    live = weights.decode_jpeg_u8(REFERENCE / "test_imgs" / "ILSVRC2012_val_00004749.jpeg")
    assert torch.equal(live, weights.load_u8_image_bin(golden_dir / "ILSVRC2012_val_00004749_u8hwc.bin"))
