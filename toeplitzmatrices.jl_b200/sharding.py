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
    api.spectrum_ready(F)
    return F


class Communicator:
    """The C ABI's own NCCL communicator (``tb200_comm_*``): what a non-torch host (the Julia extension) uses to
    broadcast the spectrum.  ``Communicator.all_local(ndev)`` = one process driving ``ndev`` GPUs;
    ``Communicator.from_torch_group()`` = one process per GPU, the 128-byte NCCL id shipped through the existing
    ``torch.distributed`` group (any host-side transport does)."""

    def __init__(self, handle):
        self._c = handle

    @classmethod
    def all_local(cls, ndev: int, devices=None):
        import ctypes
        from . import _lib as L
        out = ctypes.c_void_p()
        devs = (ctypes.c_int * ndev)(*(devices if devices is not None else range(ndev)))
        L.check(L.lib().tb200_comm_init(ndev, devs, ctypes.byref(out)))
        return cls(out.value)

    @classmethod
    def from_torch_group(cls, group=None):
        import ctypes
        from . import _lib as L
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        ident = ctypes.create_string_buffer(128)
        if rank == 0:
            L.check(L.lib().tb200_comm_unique_id(ident))
        box = [bytes(ident.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        out = ctypes.c_void_p()
        L.check(L.lib().tb200_comm_init_rank(world, rank, box[0], ctypes.byref(out)))
        return cls(out.value)

    def size(self):
        import ctypes
        from . import _lib as L
        a, b = ctypes.c_int(), ctypes.c_int()
        L.check(L.lib().tb200_comm_size(self._c, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def bcast_spectrum(self, factorizations, root: int = 0, streams=None) -> None:
        """``factorizations[i]`` lives on local device i of the communicator (a single factorization in the
        one-process-per-GPU mode)."""
        import ctypes
        from . import _lib as L
        fs = list(factorizations) if isinstance(factorizations, (list, tuple)) else [factorizations]
        hs = (ctypes.c_void_p * len(fs))(*[f._h for f in fs])
        if streams is None:
            streams = []
            for f in fs:
                with torch.cuda.device(f.device_index):
                    streams.append(torch.cuda.current_stream().cuda_stream)
        st = (ctypes.c_void_p * len(fs))(*streams)
        L.check(L.lib().tb200_bcast_spectrum(self._c, hs, len(fs), root, st))

    def destroy(self) -> None:
        from . import _lib as L
        if self._c:
            L.lib().tb200_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def sharded_mul_(Y_local: torch.Tensor, F, X_local: torch.Tensor, alpha=True, beta=False) -> torch.Tensor:
    """Each rank multiplies its own column block; outputs stay sharded (no gather)."""
    from . import api
    return api.mul_(Y_local, F, X_local, alpha, beta)
