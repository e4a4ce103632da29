"""The C-ABI boundary (include/rnb.h <-> resnet_c_b200/librnb.so), checked without a GPU."""
import ctypes as C
import subprocess

import pytest
import torch

from resnet_c_b200 import _lib


def test_header_declares_expected_surface():
    syms = _lib.header_symbols()
    for required in ("rnb_init", "rnb_last_error", "rnb_model_create", "rnb_model_forward",
                     "rnb_model_forward_host", "rnb_model_destroy", "rnb_conv_bn_act_forward",
                     "rnb_stem_forward", "rnb_tail_forward", "rnb_conv2d_forward",
                     "rnb_batchnorm2d_forward", "rnb_relu_forward", "rnb_add_forward",
                     "rnb_maxpool2d_forward", "rnb_avgpool2d_forward", "rnb_linear_forward",
                     "rnb_argmax_forward"):
        assert required in syms


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    for name in _lib.header_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/rnb.h but not exported by librnb.so"


def test_python_prototypes_cover_the_header():
    assert sorted(_lib.PROTOTYPES) == _lib.header_symbols()


def test_signatures_have_no_cxx_or_torch_types():
    import re
    text = re.sub(r"/\*.*?\*/", "", _lib.HEADER_PATH.read_text(), flags=re.S)  # declarations only
    for banned in ("std::", "torch", "at::", "Tensor", "template", "class ", "&"):
        assert banned not in text, f"include/rnb.h must stay a plain C header (found {banned!r})"


def test_library_links_no_vendor_dnn_or_blas():
    """The hot path is hand-written: no cuDNN / cuBLAS / torch in the shared object's dependencies."""
    out = subprocess.run(["ldd", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout.lower()
    for banned in ("cudnn", "cublas", "libtorch", "c10"):
        assert banned not in out


def test_version_string():
    assert b"sm_100a" in _lib.lib().rnb_version()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu():
    lib = _lib.lib()
    rc = lib.rnb_init(0)
    assert rc != 0
    msg = lib.rnb_last_error().decode()
    assert "no CPU fallback" in msg or "CUDA" in msg
    # compute entry points refuse to run rather than silently doing something else
    buf = (C.c_float * 4)()
    rc = lib.rnb_relu_forward(C.cast(buf, C.c_void_p), C.cast(buf, C.c_void_p), 4, None)
    assert rc != 0
    handle = C.c_void_p()
    rc = lib.rnb_model_create(b"resnet18", 0, b"/nonexistent", 1, 0, C.byref(handle))
    assert rc != 0 and not handle.value


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_python_face_refuses_cpu_tensors():
    from resnet_c_b200 import engine
    with pytest.raises(_lib.RnbError):
        engine.relu_forward(torch.zeros(4))
