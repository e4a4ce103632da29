"""TEST ORACLE (not part of the product path; only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import this). numpy restatement of what follows the hot path:

    /root/reference/cuda/inference/main.cu:240-251   fc_out.cpu() and the arg-max loop (strict '<': on ties the
                                                     LOWEST index wins) — generalised here to top-k
    /root/reference/cuda/tensor.cuh:154-163          Tensor::save: raw native-endian float32, no header
    /root/reference/pytorch_inference.py:8-11        check_out(): the two dumps compared with allclose

The reference computes no softmax; `softmax_topk` is the textbook definition evaluated in float64 (parity unpinned
by reference goldens — it is pinned against torch.softmax / torch.topk in tests/test_oracle_ops.py)."""
import numpy as np


def softmax_topk(logits, k):
    """logits [B, n] -> (top_probs [B, k] float64, top_idx [B, k] int64, probs [B, n] float64); order: value
    descending, index ascending on equal values."""
    x = np.asarray(logits, dtype=np.float64)
    e = np.exp(x - x.max(axis=1, keepdims=True))
    p = e / e.sum(axis=1, keepdims=True)
    # stable sort on the negated LOGITS keeps the lowest index first among equals
    order = np.argsort(-np.asarray(logits, dtype=np.float64), axis=1, kind="stable")[:, :k]
    return np.take_along_axis(p, order, axis=1), order, p


def load_f32(path):
    return np.fromfile(path, dtype=np.float32)
