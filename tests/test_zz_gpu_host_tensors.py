"""The C++ drop-in's host-tensor entry (cuda/nn.cuh ResNet::predictHost -> rnb_model_forward_host) through the
whole-model driver build/resnet_infer. Kept in the last-sorted test file: it was written after the round's GPU budget
was spent, so a failure here must not keep `pytest -x` from running the validated suites before it."""
import os
import re
import subprocess

import pytest

from conftest import GOLDEN, ROOT, load_golden

pytestmark = pytest.mark.gpu

INFER_BIN = ROOT / "build" / "resnet_infer"


@pytest.fixture(scope="module")
def r152_workdir(tmp_path_factory):
    """CWD laid out the way main.cu expects: weights_bin/<key> and test_bins/<image>.bin."""
    from resnet_c_b200 import weights
    d = tmp_path_factory.mktemp("r152h")
    (d / "weights_bin").symlink_to(weights.cached_weights_dir("resnet152", 0))
    (d / "test_bins").mkdir()
    (d / "test_bins" / "ILSVRC2012_val_00004749.bin").symlink_to(GOLDEN / "ILSVRC2012_val_00004749.bin")
    return d


@pytest.mark.parametrize("dtype", ["bf16", "tf32"])
def test_whole_model_driver_from_host_tensors(r152_workdir, dtype):
    """ResNet::predictHost: the reference's load -> loadToCuda -> forward -> cpu() sequence (main.cu:233-251) in one call
    on CPU tensors, with the host cores rounding the batch to BF16 where the stem takes it (forced here; the TF32 model
    keeps plain copies): logits bit-identical to the device-tensor path, the reference's index for every image."""
    assert INFER_BIN.exists(), "build/resnet_infer missing: run python -m resnet_c_b200.build dropin"
    expect = load_golden("ref_class_resnet152")["top1"].tolist()
    env = dict(os.environ, RNB_CHECK_HOST_PATH="1", RNB_HOST_PACK="1")
    r = subprocess.run([str(INFER_BIN), "resnet152", dtype, "40"], cwd=r152_workdir, capture_output=True,
                       text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-500:] + r.stderr[-2000:]
    assert "host path: identical" in r.stdout
    assert [int(m) for m in re.findall(r"max index is (\d+)", r.stdout)] == expect * 40
