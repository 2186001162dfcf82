"""Multi-GPU plumbing for the batched paths.

Columns / systems are independent (src/linearalgebra.jl:95-97), so a batch shards over
ranks by contiguous column blocks with NO data-path collective; the only exchange is the
spectrum of the factorization, broadcast once (NCCL over NVLink on GPUs; the same code runs
over gloo on CPU tensors in the tests).  One process per GPU, ``torch.distributed`` for the
plumbing.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def column_shard(ncols: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous column block [start, stop) of ``rank``; the remainder goes to the first ranks.
    Blocks start on even columns when possible so real columns keep pairing up (two real
    columns ride one complex FFT)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    pairs = (ncols + 1) // 2
    base, rem = divmod(pairs, world)
    p0 = rank * base + min(rank, rem)
    p1 = p0 + base + (1 if rank < rem else 0)
    return min(2 * p0, ncols), min(2 * p1, ncols)


def all_shards(ncols: int, world: int) -> List[Tuple[int, int]]:
    return [column_shard(ncols, world, r) for r in range(world)]


def broadcast_bytes(buf: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Broadcast a raw byte buffer (the spectrum) from ``src``; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(buf, src=src, group=group)
    return buf


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing reduction the bench uses: the slowest rank defines the step time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def factorize_shared(make_matrix, kind: int, dtype, shape, src: int = 0, group=None):
    """``factorize`` once on rank ``src`` and broadcast the spectrum; other ranks allocate an empty
    factorization of the same plan.  ``make_matrix()`` builds the device matrix (called on ``src``)."""
    from . import api
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if rank == src:
        F = api.factorize(make_matrix())
    else:
        F = api.factorize_empty_like(kind, dtype, shape)
    torch.cuda.synchronize()
    broadcast_bytes(api.spectrum_tensor(F), src, group)
    torch.cuda.synchronize()
    return F


def sharded_mul_(Y_local: torch.Tensor, F, X_local: torch.Tensor, alpha=True, beta=False) -> torch.Tensor:
    """Each rank multiplies its own column block; outputs stay sharded (no gather)."""
    from . import api
    return api.mul_(Y_local, F, X_local, alpha, beta)
