"""Multi-GPU host logic: the batch shards by image (independent units), each rank runs a full
replica on its slice with no data-path communication, and ONE all-gather of logits + top-1 closes
the step (BASELINE.json north_star). The reference has no multi-GPU path at all
(`B = 1`, single device: cuda/inference/main.cu:230), so there is nothing to mirror beyond the
per-image semantics.

torch.distributed is plumbing here: NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests.
"""
from __future__ import annotations

import torch


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` images owned by `rank`; earlier ranks take the remainder."""
    if world <= 0 or not 0 <= rank < world or total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def any_rank(flag: bool, device=None, group=None) -> bool:
    """True on EVERY rank if `flag` is true on any of them (one all-reduce; no process group = this rank alone).
    For decisions a rank takes by itself — e.g. the host-packing choice each replica times on its own — when the
    code that follows holds collectives: every rank has to take the same branch."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return bool(flag)
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(t.item())


def gather_results(logits: torch.Tensor, top1: torch.Tensor, total: int, group=None):
    """All-gather per-rank (logits [n_r, classes], top1 [n_r]) into ([total, classes], [total]) in image
    order. Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    lo, hi = sizes[rank]
    if logits.shape[0] != hi - lo or top1.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {logits.shape[0]} rows, expected {hi - lo}")
    width = max(h - l for l, h in sizes)
    classes = logits.shape[1]
    pad_l = torch.zeros(width, classes, dtype=logits.dtype, device=logits.device)
    pad_t = torch.zeros(width, dtype=top1.dtype, device=top1.device)
    pad_l[: hi - lo] = logits
    pad_t[: hi - lo] = top1
    all_l = torch.empty(world * width, classes, dtype=logits.dtype, device=logits.device)
    all_t = torch.empty(world * width, dtype=top1.dtype, device=top1.device)
    dist.all_gather_into_tensor(all_l, pad_l, group=group)
    dist.all_gather_into_tensor(all_t, pad_t, group=group)
    keep = torch.cat([torch.arange(r * width, r * width + (h - l), device=logits.device)
                      for r, (l, h) in enumerate(sizes)])
    return all_l.index_select(0, keep), all_t.index_select(0, keep)


def sharded_forward(model, x_global_host: torch.Tensor, group=None):
    """Each rank forwards its slice of `x_global_host` ([total,3,224,224], host) on its own GPU and all
    ranks return the full ([total, classes], [total]) result."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = x_global_host.shape[0]
    lo, hi = shard_bounds(total, world, rank)
    dev = torch.device("cuda", model.device)
    if hi > lo:
        logits, top1 = model.forward(x_global_host[lo:hi].to(dev))
    else:
        logits = torch.empty(0, model.num_classes, device=dev)
        top1 = torch.empty(0, dtype=torch.int32, device=dev)
    return gather_results(logits, top1, total, group)
