"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: column partition, the spectrum
broadcast plumbing and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ncols, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from toeplitz_b200 import sharding
    lo, hi = sharding.column_shard(ncols, world, rank)
    # "spectrum": rank 0 owns the real bytes, the others an empty buffer of the same plan
    spec = torch.arange(4096, dtype=torch.float64).view(torch.uint8).clone() if rank == 0 else torch.zeros(4096 * 8, dtype=torch.uint8)
    sharding.broadcast_bytes(spec, src=0)
    ok_spec = bool(torch.equal(spec.view(torch.float64), torch.arange(4096, dtype=torch.float64)))
    # every rank applies the same operator to its column block; blocks tile the batch exactly
    X = torch.from_numpy(np.random.default_rng(0).standard_normal((8, ncols)))
    w = spec.view(torch.float64)[:8]
    Y_local = X[:, lo:hi] * w[:, None]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, Y_local.numpy()))
    t = sharding.max_over_ranks(1.0 + rank)
    if rank == 0:
        out.put((gathered, ok_spec, t))
    else:
        out.put((None, ok_spec, t))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("ncols", [1024, 7, 2])
def test_two_rank_sharding(ncols):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ncols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "spectrum broadcast failed"
    assert all(abs(r[2] - 2.0) < 1e-12 for r in res), "max-over-ranks reduction"
    gathered = [r[0] for r in res if r[0] is not None][0]
    cover = np.zeros(ncols, dtype=int)
    X = np.random.default_rng(0).standard_normal((8, ncols))
    w = np.arange(8, dtype=float)
    for lo, hi, Y in gathered:
        cover[lo:hi] += 1
        assert np.allclose(Y, X[:, lo:hi] * w[:, None])
    assert (cover == 1).all(), "column blocks must tile the batch exactly once"


def test_column_shard_properties():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from toeplitz_b200 import sharding
    for ncols in (0, 1, 2, 5, 255, 256, 1024, 1000003):
        for world in (1, 2, 3, 4, 8):
            sh = sharding.all_shards(ncols, world)
            assert sh[0][0] == 0 and sh[-1][1] == ncols
            for (a0, a1), (b0, b1) in zip(sh, sh[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 2 == 0 for lo, _ in sh if lo < ncols)
            sizes = [hi - lo for lo, hi in sh]
            assert max(sizes) - min(sizes) <= 3      # one pair of columns + an odd tail
