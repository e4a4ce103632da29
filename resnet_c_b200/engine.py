"""Python face of librnb.so for tests and benchmarks.

torch is used for device memory and streams only; every number comes out of the CUDA kernels behind
include/rnb.h. Names follow the reference's host API (cuda/nn.cuh): Conv2d / BatchNorm2d / Pool2d /
Linear forwards, reluForward, addForward, plus the whole-model object.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import DTYPES, RnbError, check


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32_cuda(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RnbError(f"{what} must be a CUDA tensor: librnb has no CPU path")
    if t.dtype != torch.float32:
        raise RnbError(f"{what} must be float32")
    return t.contiguous()


def conv_out(x: int, k: int, stride: int, pad: int) -> int:
    """convOutputSize, cuda/ops.cuh:9-13."""
    return (2 * pad + x - k) // stride + 1


# --------------------------------------------------------------------------- per-op fp32 NCHW
def conv2d_forward(x, weight, stride=1, padding=0):
    """Conv2d::forward, cuda/nn.cu:3-16."""
    x, weight = _f32_cuda(x, "x"), _f32_cuda(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, Cin2, k, k2 = weight.shape
    assert Cin == Cin2 and k == k2
    out = torch.empty(B, Cout, conv_out(H, k, stride, padding), conv_out(W, k, stride, padding),
                      device=x.device, dtype=torch.float32)
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_conv2d_forward(_ptr(x), _ptr(out), _ptr(weight), B, Cin, H, W, Cout, k,
                                        stride, padding, _stream()))
    return out


def batchnorm2d_forward(x, weight, bias, mean, var, out=None):
    """BatchNorm2d::forward, cuda/nn.cu:18-29 (in place when out is x)."""
    x = _f32_cuda(x, "x")
    B, Cc, H, W = x.shape
    out = torch.empty_like(x) if out is None else out
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_batchnorm2d_forward(_ptr(x), _ptr(out), _ptr(_f32_cuda(weight, "weight")),
                                             _ptr(_f32_cuda(bias, "bias")), _ptr(_f32_cuda(mean, "mean")),
                                             _ptr(_f32_cuda(var, "var")), B, Cc, H * W, _stream()))
    return out


def relu_forward(x, out=None):
    """reluForward, cuda/nn.cu:66-75."""
    x = _f32_cuda(x, "x")
    out = torch.empty_like(x) if out is None else out
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_relu_forward(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def add_forward(a, b, out=None):
    """addForward, cuda/nn.cu:77-87."""
    a, b = _f32_cuda(a, "a"), _f32_cuda(b, "b")
    assert a.shape == b.shape
    out = torch.empty_like(a) if out is None else out
    _lib.init(a.device.index or 0)
    check(_lib.lib().rnb_add_forward(_ptr(a), _ptr(b), _ptr(out), a.numel(), _stream()))
    return out


def _pool(fn_name, x, k, stride, padding):
    x = _f32_cuda(x, "x")
    B, Cc, H, W = x.shape
    out = torch.empty(B, Cc, conv_out(H, k, stride, padding), conv_out(W, k, stride, padding),
                      device=x.device, dtype=torch.float32)
    _lib.init(x.device.index or 0)
    check(getattr(_lib.lib(), fn_name)(_ptr(x), _ptr(out), B, Cc, H, W, k, stride, padding, _stream()))
    return out


def maxpool2d_forward(x, k, stride=1, padding=0):
    """Pool2d::maxforward, cuda/nn.cu:43-53."""
    return _pool("rnb_maxpool2d_forward", x, k, stride, padding)


def avgpool2d_forward(x, k, stride=1, padding=0):
    """Pool2d::avgforward, cuda/nn.cu:31-41."""
    return _pool("rnb_avgpool2d_forward", x, k, stride, padding)


def linear_forward(x, weight, bias=None):
    """Linear::forward, cuda/nn.cu:55-64."""
    x, weight = _f32_cuda(x, "x"), _f32_cuda(weight, "weight")
    B, fin = x.shape
    fout = weight.shape[0]
    out = torch.empty(B, fout, device=x.device, dtype=torch.float32)
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_linear_forward(_ptr(x), _ptr(out), _ptr(weight),
                                        _ptr(None if bias is None else _f32_cuda(bias, "bias")), B, fin,
                                        fout, _stream()))
    return out


def argmax_forward(x):
    """Row arg-max with the tie rule of cuda/inference/main.cu:243-251."""
    x = _f32_cuda(x, "x")
    B, n = x.shape
    out = torch.empty(B, device=x.device, dtype=torch.int32)
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_argmax_forward(_ptr(x), _ptr(out), B, n, _stream()))
    return out


def softmax_topk_forward(logits, k=5, full=False):
    """Row softmax + top-k of the logits (the caller's next step after main.cu:240-251).
    Returns (top_probs [B,k], top_idx [B,k] int32[, probs [B,n])."""
    logits = _f32_cuda(logits, "logits")
    B, n = logits.shape
    top_p = torch.empty(B, k, device=logits.device, dtype=torch.float32)
    top_i = torch.empty(B, k, device=logits.device, dtype=torch.int32)
    probs = torch.empty(B, n, device=logits.device, dtype=torch.float32) if full else None
    _lib.init(logits.device.index or 0)
    check(_lib.lib().rnb_softmax_topk_forward(_ptr(logits), _ptr(probs), _ptr(top_p), _ptr(top_i), B, n, k,
                                              _stream()))
    return (top_p, top_i, probs) if full else (top_p, top_i)


def save_f32(t, path):
    """Raw FP32 dump of a CUDA tensor (Tensor::save format, tensor.cuh:154-163)."""
    t = _f32_cuda(t, "tensor")
    _lib.init(t.device.index or 0)
    check(_lib.lib().rnb_save_f32(_ptr(t), t.numel(), str(path).encode()))


def resize_crop_u8(images: torch.Tensor, resize: int = 256, crop: int = 224) -> torch.Tensor:
    """[n,H,W,3] uint8 CUDA tensor (decoded images of one size) -> [n,crop,crop,3] uint8: resize + centre crop of the
    torchvision preset (convert_imgs_to_bin.py:12,18), Pillow arithmetic, bit-exact."""
    if not images.is_cuda or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
        raise RnbError("resize_crop_u8 takes a [n,H,W,3] uint8 CUDA tensor")
    images = images.contiguous()
    n, H, W, _ = images.shape
    out = torch.empty(n, crop, crop, 3, device=images.device, dtype=torch.uint8)
    _lib.init(images.device.index or 0)
    check(_lib.lib().rnb_resize_crop_u8(_ptr(images), n, H, W, _ptr(out), resize, crop, _stream()))
    return out


# --------------------------------------------------------------------------- fused tensor-core ops
def conv_bn_act_forward(x, weight, bn=None, residual=None, relu=True, stride=1, padding=0,
                        dtype="bf16"):
    """conv -> bn -> (+residual) -> relu in one tcgen05 launch (layerForward, main.cu:138-163).
    `bn` is (weight, bias, running_mean, running_var) or None."""
    x, weight = _f32_cuda(x, "x"), _f32_cuda(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, _, k, _ = weight.shape
    out = torch.empty(B, Cout, conv_out(H, k, stride, padding), conv_out(W, k, stride, padding),
                      device=x.device, dtype=torch.float32)
    bnp = [None] * 4 if bn is None else [_f32_cuda(t, "bn") for t in bn]
    res = None if residual is None else _f32_cuda(residual, "residual")
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_conv_bn_act_forward(_ptr(x), _ptr(weight), _ptr(bnp[0]), _ptr(bnp[1]),
                                             _ptr(bnp[2]), _ptr(bnp[3]), _ptr(res), _ptr(out), B, Cin, H,
                                             W, Cout, k, stride, padding, int(bool(relu)),
                                             DTYPES[dtype], _stream()))
    return out


def conv_fp8_forward(x, weight, bn=None, residual=None, relu=True, stride=1, padding=0, in_scale=1.0, res_scale=1.0,
                     out_scale=1.0):
    """One FP8 (E4M3) conv with explicit per-tensor scales (rnb_conv_fp8_forward); returns the de-quantised output."""
    x, weight = _f32_cuda(x, "x"), _f32_cuda(weight, "weight")
    B, Cin, H, W = x.shape
    Cout, _, k, _ = weight.shape
    out = torch.empty(B, Cout, conv_out(H, k, stride, padding), conv_out(W, k, stride, padding),
                      device=x.device, dtype=torch.float32)
    bnp = [None] * 4 if bn is None else [_f32_cuda(t, "bn") for t in bn]
    res = None if residual is None else _f32_cuda(residual, "residual")
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_conv_fp8_forward(_ptr(x), _ptr(weight), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]), _ptr(bnp[3]),
                                          _ptr(res), _ptr(out), B, Cin, H, W, Cout, k, stride, padding, int(bool(relu)),
                                          float(in_scale), float(res_scale), float(out_scale), _stream()))
    return out


def stem_forward(x, weight, bn=None, dtype="bf16"):
    """conv 7x7/2 + bn + relu + maxpool 3x3/2 (main.cu:176-192)."""
    x, weight = _f32_cuda(x, "x"), _f32_cuda(weight, "weight")
    B, _, H, W = x.shape
    oh, ow = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
    out = torch.empty(B, 64, conv_out(oh, 3, 2, 1), conv_out(ow, 3, 2, 1), device=x.device,
                      dtype=torch.float32)
    bnp = [None] * 4 if bn is None else [_f32_cuda(t, "bn") for t in bn]
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_stem_forward(_ptr(x), _ptr(weight), _ptr(bnp[0]), _ptr(bnp[1]), _ptr(bnp[2]),
                                      _ptr(bnp[3]), _ptr(out), B, H, W, DTYPES[dtype], _stream()))
    return out


def tail_forward(x, fc_weight, fc_bias):
    """avgpool + fc + arg-max (main.cu:213-224, 243-251). x is [B,C,h,w] fp32."""
    x = _f32_cuda(x, "x")
    B, Cc, H, W = x.shape
    classes = fc_weight.shape[0]
    logits = torch.empty(B, classes, device=x.device, dtype=torch.float32)
    top1 = torch.empty(B, device=x.device, dtype=torch.int32)
    _lib.init(x.device.index or 0)
    check(_lib.lib().rnb_tail_forward(_ptr(x), _ptr(_f32_cuda(fc_weight, "fc_weight")),
                                      _ptr(_f32_cuda(fc_bias, "fc_bias")), _ptr(logits), _ptr(top1), B,
                                      Cc, H * W, classes, _stream()))
    return logits, top1


# --------------------------------------------------------------------------- whole model
class ResNet:
    """Planned engine over a save_weights.py-format directory (ResnetModel, main.cu:91-125)."""

    def __init__(self, arch: str, weights_dir, dtype: str = "bf16", max_batch: int = 256,
                 chunk: int = 0, device: int = 0):
        _lib.init(device)
        self.arch, self.dtype, self.max_batch, self.device = arch, dtype, max_batch, device
        handle = C.c_void_p()
        check(_lib.lib().rnb_model_create(arch.encode(), DTYPES[dtype], str(weights_dir).encode(),
                                          max_batch, chunk, C.byref(handle)))
        self._h = handle
        self.num_classes = _lib.lib().rnb_model_num_classes(self._h)
        self.flops_per_image = _lib.lib().rnb_model_flops_per_image(self._h)

    @classmethod
    def from_packed(cls, path, max_batch: int = 256, chunk: int = 0, device: int = 0):
        """Model from a pre-packed weight blob written by save_packed() (one read + one H2D copy)."""
        _lib.init(device)
        self = cls.__new__(cls)
        self.arch, self.dtype, self.max_batch, self.device = "packed", "packed", max_batch, device
        handle = C.c_void_p()
        check(_lib.lib().rnb_model_create_packed(str(path).encode(), max_batch, chunk, C.byref(handle)))
        self._h = handle
        self.num_classes = _lib.lib().rnb_model_num_classes(self._h)
        self.flops_per_image = _lib.lib().rnb_model_flops_per_image(self._h)
        return self

    def save_packed(self, path) -> None:
        check(_lib.lib().rnb_model_save_packed(self._h, str(path).encode()))

    def calibrate(self, x: torch.Tensor) -> None:
        """FP8 models: fix the activation scales from this batch (no-op for other dtypes / once calibrated)."""
        x = _f32_cuda(x, "x")
        check(_lib.lib().rnb_model_calibrate(self._h, _ptr(x), x.shape[0]))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().rnb_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches_per_forward(self, batch: int) -> int:
        return _lib.lib().rnb_model_launches_per_forward(self._h, batch)

    def forward(self, x: torch.Tensor, logits=None, top1=None):
        """x: [B,3,224,224] fp32 CUDA tensor -> (logits [B,classes] fp32, top1 [B] int32)."""
        x = _f32_cuda(x, "x")
        B = x.shape[0]
        if logits is None:
            logits = torch.empty(B, self.num_classes, device=x.device, dtype=torch.float32)
        if top1 is None:
            top1 = torch.empty(B, device=x.device, dtype=torch.int32)
        check(_lib.lib().rnb_model_forward(self._h, _ptr(x), B, _ptr(logits), _ptr(top1), _stream()))
        return logits, top1

    def forward_host(self, x: torch.Tensor, logits=None, top1=None):
        """x: [B,3,224,224] fp32 HOST tensor (pinned for full PCIe speed) -> host logits / top1.
        Host->device copy, compute and device->host copy all happen inside the call."""
        if x.is_cuda or x.dtype != torch.float32:
            raise RnbError("forward_host takes a float32 host tensor")
        x = x.contiguous()
        B = x.shape[0]
        if logits is None:
            logits = torch.empty(B, self.num_classes, dtype=torch.float32).pin_memory()
        if top1 is None:
            top1 = torch.empty(B, dtype=torch.int32).pin_memory()
        check(_lib.lib().rnb_model_forward_host(self._h, _ptr(x), B, _ptr(logits), _ptr(top1)))
        return logits, top1

    def submit_host(self, slot: int, x: torch.Tensor, logits: torch.Tensor, top1: torch.Tensor) -> None:
        """Pipelined host path: queue H2D + forward + D2H of one (pinned) host batch into slot 0/1 and
        return at once; results are valid after wait_host(slot)."""
        if x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            raise RnbError("submit_host takes a contiguous float32 host tensor")
        check(_lib.lib().rnb_model_submit_host(self._h, slot, _ptr(x), x.shape[0], _ptr(logits), _ptr(top1)))

    def forward_bf16(self, x: torch.Tensor, logits=None, top1=None):
        """x: [B,3,224,224] bfloat16 CUDA tensor = the FP32 image rounded to nearest even (what the packed host paths
        upload). Bit-identical to forward() on the FP32 tensor; BF16 / FP8 models only."""
        if not x.is_cuda or x.dtype != torch.bfloat16 or x.dim() != 4:
            raise RnbError("forward_bf16 takes a [B,3,H,W] bfloat16 CUDA tensor")
        x = x.contiguous()
        B = x.shape[0]
        if logits is None:
            logits = torch.empty(B, self.num_classes, device=x.device, dtype=torch.float32)
        if top1 is None:
            top1 = torch.empty(B, device=x.device, dtype=torch.int32)
        check(_lib.lib().rnb_model_forward_bf16(self._h, _ptr(x), B, _ptr(logits), _ptr(top1), _stream()))
        return logits, top1

    def set_host_pack(self, mode: int) -> None:
        """Host paths: -1 decide by timing at the next host call, 0 plain FP32 copies, 1 round every image to BF16 on the
        host cores."""
        check(_lib.lib().rnb_model_set_host_pack(self._h, int(mode)))

    def set_host_pack_fraction(self, fraction: float) -> None:
        """Host paths: this fraction of every batch (whole 16-image pieces) is rounded to BF16 on the host cores, the
        rest crosses PCIe as FP32."""
        check(_lib.lib().rnb_model_set_host_pack_fraction(self._h, float(fraction)))

    def host_pack(self) -> dict:
        """Current choice of the host paths and what the decision measured (GB/s; zeros when forced)."""
        g = (C.c_double * 4)()
        choice = _lib.lib().rnb_model_host_pack(self._h, g)
        return {"choice": choice, "fraction": g[3], "threads": _lib.lib().rnb_host_pack_threads(),
                "convert_gbps": g[0], "h2d_f32_gbps": g[1], "h2d_bf16_gbps": g[2]}

    def forward_u8(self, x: torch.Tensor, logits=None, top1=None):
        """x: [B,224,224,3] uint8 CUDA tensor (decoded, resized, cropped image, HWC). The /255 + mean/std
        normalisation of convert_imgs_to_bin.py:18 runs inside the stem's layout pre-pass."""
        if not x.is_cuda or x.dtype != torch.uint8 or x.dim() != 4 or x.shape[-1] != 3:
            raise RnbError("forward_u8 takes a [B,H,W,3] uint8 CUDA tensor")
        x = x.contiguous()
        B = x.shape[0]
        if logits is None:
            logits = torch.empty(B, self.num_classes, device=x.device, dtype=torch.float32)
        if top1 is None:
            top1 = torch.empty(B, device=x.device, dtype=torch.int32)
        check(_lib.lib().rnb_model_forward_u8(self._h, _ptr(x), B, _ptr(logits), _ptr(top1), _stream()))
        return logits, top1

    def submit_host_u8(self, slot: int, x: torch.Tensor, logits: torch.Tensor, top1: torch.Tensor) -> None:
        """submit_host for a [B,224,224,3] uint8 HOST tensor (a quarter of the PCIe bytes)."""
        if x.is_cuda or x.dtype != torch.uint8 or not x.is_contiguous():
            raise RnbError("submit_host_u8 takes a contiguous uint8 host tensor")
        check(_lib.lib().rnb_model_submit_host_u8(self._h, slot, _ptr(x), x.shape[0], _ptr(logits), _ptr(top1)))

    def set_normalization(self, mean, std) -> None:
        m = (C.c_float * 3)(*[float(v) for v in mean])
        s = (C.c_float * 3)(*[float(v) for v in std])
        check(_lib.lib().rnb_model_set_normalization(self._h, m, s))

    def wait_host(self, slot: int) -> None:
        check(_lib.lib().rnb_model_wait_host(self._h, slot))

    KINDS = ("stem_conv", "maxpool", "conv_igemm", "avgpool", "fc", "argmax")

    def profile(self, x: torch.Tensor, iters: int = 3):
        """Per-launch device times of one chunk: list of dicts(kind, ms, flops, bytes)."""
        x = _f32_cuda(x, "x")
        cap = 1024
        kind = (C.c_int * cap)()
        ms = (C.c_float * cap)()
        flops = (C.c_double * cap)()
        nbytes = (C.c_double * cap)()
        n = C.c_int()
        check(_lib.lib().rnb_model_profile(self._h, _ptr(x), x.shape[0], iters, kind, ms, flops, nbytes,
                                           cap, C.byref(n), _stream()))
        return [dict(kind=self.KINDS[kind[i]], ms=ms[i], flops=flops[i], bytes=nbytes[i])
                for i in range(n.value)]

    def repeat_launch(self, batch: int, index: int, repeat: int) -> None:
        """Enqueue conv launch `index` of the plan for `batch` images `repeat` times (profiling aid, no sync)."""
        check(_lib.lib().rnb_model_repeat_launch(self._h, batch, index, repeat, _stream()))

    def activation(self, name: str) -> torch.Tensor:
        """Intermediate activation of the last forward (first chunk) as fp32 NCHW (flat)."""
        n = C.c_int64()
        check(_lib.lib().rnb_model_get_activation(self._h, name.encode(), None, C.byref(n), _stream()))
        out = torch.empty(n.value, device=f"cuda:{self.device}", dtype=torch.float32)
        check(_lib.lib().rnb_model_get_activation(self._h, name.encode(), _ptr(out), C.byref(n), _stream()))
        return out


# --------------------------------------------------------------------------- replicas on several GPUs, one process
class ResNetGroup:
    """rnb_group_*: one replica per device in THIS process; the batch shards by image and every replica's last
    kernels store their rows of logits / top-1 straight into the gathering buffers on devices[0] (include/rnb.h)."""

    def __init__(self, arch: str, weights_dir, devices, dtype: str = "bf16", max_batch_per_device: int = 256):
        self.devices = [int(d) for d in devices]
        self.arch, self.dtype = arch, dtype
        handle = C.c_void_p()
        devs = (C.c_int * len(self.devices))(*self.devices)
        check(_lib.lib().rnb_group_create(arch.encode(), DTYPES[dtype], str(weights_dir).encode(), devs,
                                          len(self.devices), max_batch_per_device, C.byref(handle)))
        self._h = handle
        self.num_classes = _lib.lib().rnb_model_num_classes(_lib.lib().rnb_group_model(self._h, 0))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().rnb_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return _lib.lib().rnb_group_size(self._h)

    def direct_stores(self, r: int) -> bool:
        return bool(_lib.lib().rnb_group_direct_stores(self._h, r))

    def shard(self, batch: int, r: int):
        first, count = C.c_int(), C.c_int()
        check(_lib.lib().rnb_group_shard(self._h, batch, r, C.byref(first), C.byref(count)))
        return first.value, count.value

    def warmup(self, batch: int) -> None:
        check(_lib.lib().rnb_group_warmup(self._h, batch))

    def forward(self, shards, logits=None, top1=None):
        """shards[r]: replica r's slice on devices[r] (float32 [n_r,3,224,224], or uint8 [n_r,224,224,3]).
        Returns (logits [B,classes], top1 [B]) on devices[0], valid after synchronize()."""
        batch = sum(int(s.shape[0]) for s in shards)
        u8 = shards[0].dtype == torch.uint8
        root = torch.device("cuda", self.devices[0])
        if logits is None:
            logits = torch.empty(batch, self.num_classes, device=root, dtype=torch.float32)
        if top1 is None:
            top1 = torch.empty(batch, device=root, dtype=torch.int32)
        ptrs = (C.c_void_p * len(shards))(*[s.data_ptr() if s.shape[0] else None for s in shards])
        fn = _lib.lib().rnb_group_forward_u8 if u8 else _lib.lib().rnb_group_forward
        check(fn(self._h, ptrs, batch, _ptr(logits), _ptr(top1), None))
        return logits, top1

    def synchronize(self) -> None:
        check(_lib.lib().rnb_group_synchronize(self._h))

    def forward_host(self, x: torch.Tensor, logits=None, top1=None):
        """x: host [B,3,224,224] float32 -> host (logits, top1); blocking."""
        x = x.contiguous()
        B = x.shape[0]
        if logits is None:
            logits = torch.empty(B, self.num_classes, dtype=torch.float32).pin_memory()
        if top1 is None:
            top1 = torch.empty(B, dtype=torch.int32).pin_memory()
        check(_lib.lib().rnb_group_forward_host(self._h, _ptr(x), B, _ptr(logits), _ptr(top1)))
        return logits, top1

    def submit_host(self, slot: int, x: torch.Tensor, logits: torch.Tensor, top1: torch.Tensor) -> None:
        fn = _lib.lib().rnb_group_submit_host_u8 if x.dtype == torch.uint8 else _lib.lib().rnb_group_submit_host
        check(fn(self._h, slot, _ptr(x), x.shape[0], _ptr(logits), _ptr(top1)))

    def wait_host(self, slot: int) -> None:
        check(_lib.lib().rnb_group_wait_host(self._h, slot))


# --------------------------------------------------------------------------- planned block (module classes)
class Block:
    """rnb_block_*: one conv (+BN), BasicBlock or Bottleneck with BN folded and weights packed once (include/rnb.h).
    `convs` is a list of dicts(w=[Cout,Cin,k,k] cuda fp32, bn=(weight,bias,mean,var)|None, stride, pad) in the
    order conv1, conv2[, conv3][, downsample]."""

    KINDS = {"conv": 0, "basic": 1, "bottleneck": 2}

    def __init__(self, kind: str, convs, dtype: str = "bf16"):
        dev = convs[0]["w"].device.index or 0
        _lib.init(dev)
        self.kind = kind
        arr = (_lib.ConvParams * len(convs))()
        self._keep = []
        for i, c in enumerate(convs):
            w = _f32_cuda(c["w"], "w")
            bn = c.get("bn")
            bnp = [None] * 4 if bn is None else [_f32_cuda(t, "bn") for t in bn]
            self._keep += [w] + bnp
            arr[i].w = w.data_ptr()
            for name, t in zip(("bn_weight", "bn_bias", "bn_mean", "bn_var"), bnp):
                setattr(arr[i], name, None if t is None else t.data_ptr())
            arr[i].Cout, arr[i].Cin, arr[i].k = w.shape[0], w.shape[1], w.shape[2]
            arr[i].stride, arr[i].pad = int(c.get("stride", 1)), int(c.get("pad", 0))
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(_lib.lib().rnb_block_create(self.KINDS[kind], DTYPES[dtype], arr, len(convs), C.byref(h)))
        self._h = h
        self._keep = None  # the weights were copied
        c_last = convs[2] if kind == "bottleneck" else (convs[1] if kind == "basic" else convs[0])
        self.cout = c_last["w"].shape[0]
        c1, cs = convs[0], (convs[1] if kind == "bottleneck" else convs[0])
        self._geom = (c1["w"].shape[2], int(c1.get("stride", 1)), int(c1.get("pad", 0))) if kind == "conv" else (
            3, int(cs.get("stride", 1)), 1)

    def forward(self, x, residual=None, relu=True):
        x = _f32_cuda(x, "x")
        B, _, H, W = x.shape
        k, s, p = self._geom
        out = torch.empty(B, self.cout, conv_out(H, k, s, p), conv_out(W, k, s, p), device=x.device, dtype=torch.float32)
        res = None if residual is None else _f32_cuda(residual, "residual")
        check(_lib.lib().rnb_block_forward(self._h, _ptr(x), B, H, W, _ptr(res), int(bool(relu)), _ptr(out), _stream()))
        return out

    def num_launches(self, B, H, W) -> int:
        return _lib.lib().rnb_block_num_launches(self._h, B, H, W)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().rnb_block_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
