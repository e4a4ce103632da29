"""Shared test plumbing.

  -m "not gpu"  oracle vs golden vectors, host logic, C-ABI loads and exports every declared symbol
                (no compute calls) — runs in minutes on a CPU-only box.
  -m gpu        parity tests proper: the CUDA path (through the C ABI in include/rnb.h) against the
                oracle, the committed goldens and, where present, the reference's own CUDA build.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")  # exists in the build container only — never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def jpeg_tensor():
    """The reference's test image, preprocessed (committed fixture, see make_golden.py)."""
    from resnet_c_b200 import weights
    return weights.load_image_bin(GOLDEN / "ILSVRC2012_val_00004749.bin")


@pytest.fixture(scope="session")
def oracle_lib():
    """oracle/libref_ops.so, (re)built on demand — building the checker is not using it."""
    from resnet_c_b200 import build
    build.build_oracle()
    from oracle import ref_model
    return ref_model


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz")


def rel_err(got, ref):
    """max|d| / max|y| per row, worst row — the definition SURVEY.md section 8(c) fixes."""
    got = np.asarray(got, dtype=np.float64).reshape(len(ref), -1)
    ref = np.asarray(ref, dtype=np.float64).reshape(len(ref), -1)
    return float((np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)).max())
