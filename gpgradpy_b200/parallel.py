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


_gather_buf: dict = {}


def _buffers(per: int, W: int, size: int, dtype, device):
    """Padded send row block and receive table, kept per (shape, dtype, device): the scan calls this once per batch,
    so the two allocations and the zero fill would otherwise sit on its latency path."""
    key = (per, W, size, dtype, str(device))
    b = _gather_buf.get(key)
    if b is None:
        if len(_gather_buf) > 16:
            _gather_buf.clear()
        b = (torch.zeros((per, W), dtype=dtype, device=device), torch.empty((size * per, W), dtype=dtype, device=device))
        _gather_buf[key] = b
    return b


def gather_rows(local, B: int, group=None, width: int | None = None) -> torch.Tensor:
    """All-gather the [b_local, W] result rows of every rank into the full [B, W] table (same on all ranks).
    A rank whose shard is empty (more ranks than rows) passes `local=None` (or a [0, W] tensor) together with `width`
    and a reference tensor is not needed: every rank always enters the collective."""
    rank, size = world(group)
    if size == 1:
        return local
    if local is None or local.shape[0] == 0:
        assert width is not None or local is not None, "an empty shard must state the row width"
        W = int(width if width is not None else local.shape[1])
        dev = local.device if local is not None else (torch.device("cuda", torch.cuda.current_device())
                                                      if dist.get_backend(group) == "nccl" else torch.device("cpu"))
        dtype = local.dtype if local is not None else torch.float64
        nloc = 0
    else:
        W, dev, dtype, nloc = local.shape[1], local.device, local.dtype, local.shape[0]
    per = (B + size - 1) // size
    pad, full = _buffers(per, W, size, dtype, dev)
    if nloc:
        pad[:nloc] = local
    if dev.type == "cuda":
        dist.all_gather_into_tensor(full, pad, group=group)
    else:
        dist.all_gather(list(full.view(size, per, W).unbind(0)), pad, group=group)
    if B % size == 0:
        return full.clone()            # equal shards: the padded table IS the table
    rows = []
    for r in range(size):
        lo, hi = shard_bounds(B, r, size)
        rows.append(full[r * per: r * per + (hi - lo)])
    return torch.cat(rows, dim=0)


def sharded_eval(eval_rows, cand: torch.Tensor, group=None, width: int | None = None) -> torch.Tensor:
    """Evaluate `eval_rows(cand[lo:hi]) -> [hi-lo, W]` on this rank's slice and gather the full table.  With more ranks
    than rows the surplus ranks evaluate nothing but still take part in the collective (`width` = W must then be
    given, since such a rank has no result to read it from)."""
    rank, size = world(group)
    B = cand.shape[0]
    lo, hi = shard_bounds(B, rank, size)
    local = eval_rows(cand[lo:hi]) if hi > lo else None
    if local is None and size > 1 and width is None:
        raise ValueError("sharded_eval: pass width= when the process group may have more ranks than candidate rows")
    return gather_rows(local, B, group, width=width)
