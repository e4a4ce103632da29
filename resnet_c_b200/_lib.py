"""ctypes binding of librnb.so (the C ABI declared in include/rnb.h).

The library is the product; this module only declares prototypes. Importing the package works
without a GPU (so the CPU test-suite can check that the .so loads and exports every symbol), but any
compute call fails loudly when there is no B200 — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
# RNB_LIB: another build of the same C ABI (A/B runs of two library versions on one box, tools/ab_lib.sh)
LIB_PATH = Path(os.environ["RNB_LIB"]).resolve() if os.environ.get("RNB_LIB") else Path(__file__).resolve().parent / "librnb.so"
HEADER_PATH = ROOT / "include" / "rnb.h"

RNB_OK = 0
DTYPE_BF16 = 0
DTYPE_TF32 = 1
DTYPE_FP8 = 2
DTYPES = {"bf16": DTYPE_BF16, "tf32": DTYPE_TF32, "fp8": DTYPE_FP8}


class RnbError(RuntimeError):
    pass


_lib = None

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p

class ConvParams(C.Structure):
    """rnb_conv_params_t (include/rnb.h)."""
    _fields_ = [("w", C.c_void_p), ("bn_weight", C.c_void_p), ("bn_bias", C.c_void_p), ("bn_mean", C.c_void_p),
                ("bn_var", C.c_void_p), ("Cin", C.c_int), ("Cout", C.c_int), ("k", C.c_int), ("stride", C.c_int),
                ("pad", C.c_int)]


# name -> (restype, argtypes). Pointers to device memory are passed as integers (c_void_p).
PROTOTYPES = {
    "rnb_init": (C.c_int, [C.c_int]),
    "rnb_last_error": (C.c_char_p, []),
    "rnb_version": (C.c_char_p, []),
    "rnb_model_create": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "rnb_model_destroy": (C.c_int, [_vp]),
    "rnb_model_device": (C.c_int, [_vp]),
    "rnb_model_warmup": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rnb_model_calibrate": (C.c_int, [_vp, _vp, C.c_int]),
    "rnb_conv_fp8_forward": (C.c_int, [_vp] * 8 + [C.c_int] * 9 + [C.c_float] * 3 + [_vp]),
    "rnb_model_save_packed": (C.c_int, [_vp, C.c_char_p]),
    "rnb_model_create_packed": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_vp)]),
    "rnb_model_forward": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rnb_model_forward_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "rnb_model_submit_host": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "rnb_model_wait_host": (C.c_int, [_vp, C.c_int]),
    "rnb_model_set_normalization": (C.c_int, [_vp, _f32p, _f32p]),
    "rnb_model_forward_u8": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rnb_model_submit_host_u8": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "rnb_model_set_host_pack": (C.c_int, [_vp, C.c_int]),
    "rnb_model_set_host_pack_fraction": (C.c_int, [_vp, C.c_double]),
    "rnb_model_host_pack": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "rnb_host_pack_threads": (C.c_int, []),
    "rnb_host_pack_split": (C.c_double, [C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int)]),
    "rnb_f32_to_bf16_host": (C.c_int, [_vp, _vp, C.c_size_t]),
    "rnb_model_forward_bf16": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp]),
    "rnb_model_num_classes": (C.c_int, [_vp]),
    "rnb_model_num_convs": (C.c_int, [_vp]),
    "rnb_model_launches_per_forward": (C.c_int, [_vp, C.c_int]),
    "rnb_model_flops_per_image": (C.c_double, [_vp]),
    "rnb_model_profile": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(C.c_int), _f32p,
                                    C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int,
                                    C.POINTER(C.c_int), _vp]),
    "rnb_model_repeat_launch": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp]),
    "rnb_model_get_activation": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(C.c_int64), _vp]),
    "rnb_block_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(ConvParams), C.c_int, C.POINTER(_vp)]),
    "rnb_block_destroy": (C.c_int, [_vp]),
    "rnb_block_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp]),
    "rnb_block_num_launches": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "rnb_resize_crop_u8": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.c_int, C.c_int, _vp]),
    "rnb_group_create": (C.c_int, [C.c_char_p, C.c_int, C.c_char_p, C.POINTER(C.c_int), C.c_int, C.c_int,
                                   C.POINTER(_vp)]),
    "rnb_group_destroy": (C.c_int, [_vp]),
    "rnb_group_size": (C.c_int, [_vp]),
    "rnb_group_model": (_vp, [_vp, C.c_int]),
    "rnb_group_direct_stores": (C.c_int, [_vp, C.c_int]),
    "rnb_shard_bounds": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rnb_group_shard": (C.c_int, [_vp, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "rnb_group_warmup": (C.c_int, [_vp, C.c_int]),
    "rnb_group_forward": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, _vp, _vp, _vp]),
    "rnb_group_forward_u8": (C.c_int, [_vp, C.POINTER(_vp), C.c_int, _vp, _vp, _vp]),
    "rnb_group_synchronize": (C.c_int, [_vp]),
    "rnb_group_submit_host": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "rnb_group_submit_host_u8": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp, _vp]),
    "rnb_group_wait_host": (C.c_int, [_vp, C.c_int]),
    "rnb_group_forward_host": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp]),
    "rnb_conv_bn_act_forward": (C.c_int, [_vp] * 8 + [C.c_int] * 10 + [_vp]),
    "rnb_stem_forward": (C.c_int, [_vp] * 7 + [C.c_int] * 4 + [_vp]),
    "rnb_tail_forward": (C.c_int, [_vp] * 5 + [C.c_int] * 4 + [_vp]),
    "rnb_conv2d_forward": (C.c_int, [_vp] * 3 + [C.c_int] * 8 + [_vp]),
    "rnb_batchnorm2d_forward": (C.c_int, [_vp] * 6 + [C.c_int] * 3 + [_vp]),
    "rnb_relu_forward": (C.c_int, [_vp, _vp, C.c_int64, _vp]),
    "rnb_add_forward": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp]),
    "rnb_maxpool2d_forward": (C.c_int, [_vp, _vp] + [C.c_int] * 7 + [_vp]),
    "rnb_avgpool2d_forward": (C.c_int, [_vp, _vp] + [C.c_int] * 7 + [_vp]),
    "rnb_linear_forward": (C.c_int, [_vp] * 4 + [C.c_int] * 3 + [_vp]),
    "rnb_argmax_forward": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp]),
    "rnb_softmax_topk_forward": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp]),
    "rnb_save_f32": (C.c_int, [_vp, C.c_int64, C.c_char_p]),
}


def header_symbols() -> list[str]:
    """Every function name declared in include/rnb.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rnb_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    """Load librnb.so (once). Raises if the extension has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RnbError(
                f"{LIB_PATH} is missing: build it with `python -m resnet_c_b200.build librnb` "
                "(there is no Python/CPU fallback for the hot path)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != RNB_OK:
        msg = lib().rnb_last_error().decode(errors="replace")
        raise RnbError(f"librnb error {rc}: {msg}")


_initialised_device = None


def init(device: int = 0) -> None:
    global _initialised_device
    if _initialised_device != device:
        check(lib().rnb_init(device))
        _initialised_device = device
