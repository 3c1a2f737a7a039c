"""Sharding of independent hyper-parameter candidates across ranks (one process per GPU).

Only the naturally independent work is partitioned: the candidate rows of the LML scan
(optz/GpHparaX0.py:39-45) and multi-start rows (optz/OptzLkd.py:249-270).  X and y are replicated
(a few KB); each rank evaluates a contiguous slice and the scalar results are exchanged with ONE
all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests).  No matrix ever crosses GPUs.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(B: int, rank: int, size: int):
    """Contiguous slice [lo, hi) of B rows owned by `rank`; slices differ by at most one row."""
    base, rem = divmod(B, size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, B: int, group=None) -> torch.Tensor:
    """All-gather the [b_local, W] result rows of every rank into the full [B, W] table (same on all ranks)."""
    rank, size = world(group)
    if size == 1:
        return local
    W = local.shape[1]
    per = (B + size - 1) // size
    pad = torch.zeros((per, W), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    full = torch.empty((size * per, W), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(full, pad, group=group)
    else:
        dist.all_gather(list(full.view(size, per, W).unbind(0)), pad, group=group)
    rows = []
    for r in range(size):
        lo, hi = shard_bounds(B, r, size)
        rows.append(full[r * per: r * per + (hi - lo)])
    return torch.cat(rows, dim=0)


def sharded_eval(eval_rows, cand: torch.Tensor, group=None) -> torch.Tensor:
    """Evaluate `eval_rows(cand[lo:hi]) -> [hi-lo, W]` on this rank's slice and gather the full table."""
    rank, size = world(group)
    B = cand.shape[0]
    lo, hi = shard_bounds(B, rank, size)
    local = eval_rows(cand[lo:hi]) if hi > lo else None
    if local is None:
        raise RuntimeError("more ranks than candidate rows; shrink the process group")
    return gather_rows(local, B, group)
