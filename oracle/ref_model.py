"""ORACLE (test infrastructure — only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this; the product path never does).

CPU restatement of the reference's *CUDA* inference path: the op kernels are oracle/ref_ops.c (one C
function per kernel of cuda/ops.cu), and this module restates the host graph that
cuda/inference/main.cu wires by hand:
  createLayer       :53-89    weight-name scheme "layer{L}.{i}.conv{1,2,3}" / ".bn{1,2,3}" / ".downsample.{0,1}"
  layerForward      :127-166  [downsample conv->bn]  conv1->bn->relu  conv2->bn->relu  conv3->bn->add->relu
  resnet152Forward  :168-226  conv1->bn1->relu->maxpool->layer1..4->avgpool(7)->flatten->fc
  main              :243-251  per-row arg-max, first maximum wins
BasicBlock depths (18/34) follow torchvision, since the reference has no such block.

PIN: tests/test_oracle_pin.py checks this against oracle/torch_model.py (itself pinned to the
reference's PyTorch class and to torchvision) and against the committed goldens; on a GPU box the
-m gpu tests additionally compare it with the reference's own CUDA build (oracle/_ref).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libref_ops.so"

ARCH_SPECS = {
    "resnet18": (False, (2, 2, 2, 2)),
    "resnet34": (False, (3, 4, 6, 3)),
    "resnet50": (True, (3, 4, 6, 3)),
    "resnet101": (True, (3, 4, 23, 3)),
    "resnet152": (True, (3, 8, 36, 3)),  # main.cu:116-119
}

_lib = None
_u64 = C.c_uint64
_fp = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle libref_ops.so`")
        h = C.CDLL(str(LIB_PATH))
        h.ref_conv_output_size.restype = _u64
        h.ref_conv_output_size.argtypes = [_u64] * 4
        h.ref_conv2d_forward.argtypes = [_fp, _fp, _fp] + [_u64] * 10
        h.ref_maxpool2d_forward.argtypes = [_fp, _fp] + [_u64] * 9
        h.ref_avgpool2d_forward.argtypes = [_fp, _fp] + [_u64] * 9
        h.ref_linear_forward.argtypes = [_fp, _fp, _fp, _fp] + [_u64] * 3
        h.ref_relu_forward.argtypes = [_fp, _fp, _u64]
        h.ref_batchnorm2d_forward.argtypes = [_fp] * 6 + [_u64] * 3
        h.ref_add_forward.argtypes = [_fp, _fp, _fp, _u64]
        h.ref_argmax_rows.argtypes = [_fp, C.POINTER(C.c_int32), _u64, _u64]
        for name in ("ref_conv2d_forward", "ref_maxpool2d_forward", "ref_avgpool2d_forward",
                     "ref_linear_forward", "ref_relu_forward", "ref_batchnorm2d_forward",
                     "ref_add_forward", "ref_argmax_rows"):
            getattr(h, name).restype = None
        _lib = h
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def _f32(a) -> np.ndarray:
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def conv_output_size(x, k, stride, pad):
    return int(lib().ref_conv_output_size(x, k, stride, pad))


# ---- one Python function per host launcher of cuda/nn.cu
def conv2d(x, w, stride=1, pad=0):
    """Conv2d::forward (nn.cu:3-16) over conv2dForwardKernel."""
    x, w = _f32(x), _f32(w)
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    oh, ow = conv_output_size(H, k, stride, pad), conv_output_size(W, k, stride, pad)
    out = np.empty((B, Cout, oh, ow), np.float32)
    lib().ref_conv2d_forward(_p(x), _p(out), _p(w), k, stride, pad, oh, ow, B, Cin, Cout, H, W)
    return out


def batchnorm2d(x, weight, bias, mean, var):
    """BatchNorm2d::forward (nn.cu:18-29)."""
    x = _f32(x)
    B, Cc, H, W = x.shape
    out = np.empty_like(x)
    lib().ref_batchnorm2d_forward(_p(x), _p(out), _p(_f32(weight)), _p(_f32(bias)), _p(_f32(mean)),
                                  _p(_f32(var)), B, Cc, H * W)
    return out


def relu(x):
    x = _f32(x)
    out = np.empty_like(x)
    lib().ref_relu_forward(_p(x), _p(out), x.size)
    return out


def add(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty_like(a)
    lib().ref_add_forward(_p(a), _p(b), _p(out), a.size)
    return out


def _pool(fn, x, k, stride, pad):
    x = _f32(x)
    B, Cc, H, W = x.shape
    oh, ow = conv_output_size(H, k, stride, pad), conv_output_size(W, k, stride, pad)
    out = np.empty((B, Cc, oh, ow), np.float32)
    fn(_p(x), _p(out), k, stride, pad, oh, ow, B, Cc, H, W)
    return out


def maxpool2d(x, k, stride=1, pad=0):
    return _pool(lib().ref_maxpool2d_forward, x, k, stride, pad)


def avgpool2d(x, k, stride=1, pad=0):
    return _pool(lib().ref_avgpool2d_forward, x, k, stride, pad)


def linear(x, w, bias=None):
    x, w = _f32(x), _f32(w)
    B, fin = x.shape
    fout = w.shape[0]
    out = np.empty((B, fout), np.float32)
    lib().ref_linear_forward(_p(x), _p(out), _p(w), None if bias is None else _p(_f32(bias)), B, fin, fout)
    return out


def argmax_rows(x):
    x = _f32(x)
    out = np.empty(x.shape[0], np.int32)
    lib().ref_argmax_rows(_p(x), out.ctypes.data_as(C.POINTER(C.c_int32)), x.shape[0], x.shape[1])
    return out


# ---- the graph (main.cu)
def _bn(sd, name, x):
    return batchnorm2d(x, sd[name + ".weight"], sd[name + ".bias"], sd[name + ".running_mean"],
                       sd[name + ".running_var"])


def block_forward(sd, prefix, x, bottleneck, stride, has_ds):
    """One ResnetBlock of layerForward (main.cu:130-164) / one torchvision BasicBlock."""
    shortcut = x
    if has_ds:
        shortcut = _bn(sd, prefix + "downsample.1", conv2d(x, sd[prefix + "downsample.0.weight"], stride, 0))
    if bottleneck:
        y = relu(_bn(sd, prefix + "bn1", conv2d(x, sd[prefix + "conv1.weight"], 1, 0)))
        y = relu(_bn(sd, prefix + "bn2", conv2d(y, sd[prefix + "conv2.weight"], stride, 1)))
        y = _bn(sd, prefix + "bn3", conv2d(y, sd[prefix + "conv3.weight"], 1, 0))
    else:
        y = relu(_bn(sd, prefix + "bn1", conv2d(x, sd[prefix + "conv1.weight"], stride, 1)))
        y = _bn(sd, prefix + "bn2", conv2d(y, sd[prefix + "conv2.weight"], 1, 1))
    return relu(add(y, shortcut))


def resnet_forward(arch, sd, x, taps=None):
    """resnet152Forward (main.cu:168-226) for any depth; returns (logits [B,classes], top1 [B])."""
    bottleneck, counts = ARCH_SPECS[arch]
    sd = {k: _f32(v) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    y = relu(_bn(sd, "bn1", conv2d(x, sd["conv1.weight"], 2, 3)))
    if taps is not None:
        taps["stem"] = y
    y = maxpool2d(y, 3, 2, 1)
    if taps is not None:
        taps["maxpool"] = y
    in_ch = 64
    for li, n in enumerate(counts, start=1):
        mid = 64 << (li - 1)
        out_ch = mid * 4 if bottleneck else mid
        for bi in range(n):
            stride = (1 if li == 1 else 2) if bi == 0 else 1
            has_ds = bi == 0 and (stride != 1 or in_ch != out_ch)  # main.cu:71
            y = block_forward(sd, f"layer{li}.{bi}.", y, bottleneck, stride, has_ds)
            if taps is not None:
                taps[f"layer{li}.{bi}"] = y
            in_ch = out_ch
    y = avgpool2d(y, y.shape[2])  # Pool2d(2048, 7), main.cu:120
    pooled = y.reshape(y.shape[0], -1)
    if taps is not None:
        taps["avgpool"] = pooled
    logits = linear(pooled, sd["fc.weight"], sd["fc.bias"])
    return logits, argmax_rows(logits)
