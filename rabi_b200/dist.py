"""Multi-GPU plumbing: one process per GPU (torchrun), ``torch.distributed`` over NCCL on the box' NVLink/NVSwitch.

The hot path shards by observation row -- rows are independent (rbibm/rbibm/utils/batched_processing.py:5-15 already
walks them in independent chunks; the only coupling is the explicit 1/chunk factor) -- so there is NO
data-path collective.  NCCL is used for exactly two things:
  * ``gather_rows``: all_gather of the per-observation losses (and optionally x~) before the quantile summary
    (rbibm/rbibm/metric/base.py:116-135 needs all values);
  * ``allreduce_grads``: one flat all-reduce of the gradient per training step for data-parallel FIM / TRADES training.
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n_rows: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> slice:
    """Contiguous block of rows owned by ``rank``; the first ``n_rows % world`` ranks get one extra row."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(n_rows, world_size)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def gather_rows(local: Tensor, n_rows: int, group=None) -> Tensor:
    """Concatenate the per-rank row blocks (``shard_rows`` layout) on every rank.  Works for uneven shards."""
    rank, w = world()
    if w == 1:
        return local
    sizes = [shard_rows(n_rows, r, w) for r in range(w)]
    maxlen = max(s.stop - s.start for s in sizes)
    pad = torch.zeros((maxlen,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(w)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: s.stop - s.start] for o, s in zip(out, sizes)], 0)


def allreduce_grads(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> None:
    """One flat all-reduce of all ``.grad`` tensors (missing grads count as zeros so every rank sends the same size)."""
    _, w = world()
    if w == 1:
        return
    params = [p for p in params if p.requires_grad]
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= w
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off: off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n


def sharded_robustness_eval(metric, xs: Tensor, group=None) -> Tensor:
    """Each rank attacks and scores its own block of observations; returns the full per-observation loss vector on
    every rank (ready for ``summary_reduce``)."""
    sl = shard_rows(xs.shape[0])
    local = metric._eval(xs[sl].to(metric.device))
    return gather_rows(local, xs.shape[0], group=group)
