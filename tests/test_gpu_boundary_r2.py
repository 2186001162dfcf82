"""Round-2 boundary hardening, on the GPU through the C ABI:

* strided 1-D inputs (`X[:, j]` of a row-major matrix, `x[::2]`) -- the reference accepts any StridedVector
  (src/linearalgebra.jl:5,43); ADVICE r1 (high).
* one handle used from two streams at once (per-stream workspaces; the reference's shared `tmp`,
  src/ToeplitzMatrices.jl:97-101, forbids this).
* destroy right after an asynchronous product on a non-default stream (stream-ordered release).
* `tb200_mul_host` straight after an asynchronous `factorize` (waits for the spectrum by event).
* the C ABI's own NCCL communicator (`tb200_comm_*`), single-device group, and one process driving two GPUs
  (`test_comm_two_devices_single_process`, skipped on a one-GPU box).
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def devmat(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


@pytest.mark.parametrize("n", [101, 3000, 70000])
def test_strided_vectors(tb, n):
    rng = np.random.default_rng(n)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    Ao, Ad = O.Toeplitz(vc, vr), tb.Toeplitz(torch.from_numpy(vc), torch.from_numpy(vr))
    X = rng.standard_normal((n, 4))
    Xrow = torch.from_numpy(X).cuda()                 # row-major: X[:, j] has stride 4
    x2 = torch.from_numpy(rng.standard_normal(2 * n)).cuda()
    for j in (0, 3):
        col = Xrow[:, j]
        assert col.stride(0) != 1
        y = tb.mul(Ad, col).cpu().numpy()
        assert rel(y, O.matvec(Ao, X[:, j])) <= 1e-12
        yo = torch.zeros(n, dtype=torch.float64, device="cuda")
        tb.mul_(yo, Ad, col)
        assert rel(yo.cpu().numpy(), O.matvec(Ao, X[:, j])) <= 1e-12
    xs = x2[::2]
    assert rel(tb.mul(Ad, xs).cpu().numpy(), O.matvec(Ao, x2.cpu().numpy()[::2])) <= 1e-12
    # strided OUTPUT vector
    ybig = torch.zeros(2 * n, dtype=torch.float64, device="cuda")
    tb.mul_(ybig[::2], Ad, xs)
    assert rel(ybig.cpu().numpy()[::2], O.matvec(Ao, x2.cpu().numpy()[::2])) <= 1e-12
    assert float(ybig[1::2].abs().max()) == 0.0
    # Circulant ldiv! / ldiv on strided right-hand sides
    v = 0.9 ** np.arange(n)
    Co, Cd = O.Circulant(v), tb.Circulant(torch.from_numpy(v))
    b = Xrow[:, 1].clone()                            # clone of a strided view is dense; keep a strided one too
    bs = Xrow[:, 1]
    ref = O.circ_ldiv(Co, X[:, 1].copy())
    assert rel(tb.ldiv(Cd, bs).cpu().numpy(), ref) <= 1e-12
    F = tb.factorize(Cd)
    tb.circ_ldiv_(F, bs)                              # in place on the strided view
    assert rel(Xrow[:, 1].cpu().numpy(), ref) <= 1e-12
    assert rel(Xrow[:, 0].cpu().numpy(), X[:, 0]) == 0.0
    del b


@pytest.mark.parametrize("n,l", [(3000, 6), (1 << 17, 8)])
def test_one_handle_two_streams(tb, n, l):
    """Two streams multiply different batches with the SAME factorization concurrently, many times."""
    rng = np.random.default_rng(7)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    Ao = O.Toeplitz(vc, vr)
    F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc), torch.from_numpy(vr)))
    Xa, Xb = rng.standard_normal((n, l)), rng.standard_normal((n, l))
    Xad, Xbd = devmat(Xa), devmat(Xb)
    Ya, Yb = tb.colmajor(n, l), tb.colmajor(n, l)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(20):
        with torch.cuda.stream(s1):
            tb.mul_(Ya, F, Xad)
        with torch.cuda.stream(s2):
            tb.mul_(Yb, F, Xbd)
    torch.cuda.synchronize()
    assert rel(Ya.cpu().numpy(), O.matvec(Ao, Xa)) <= 1e-12
    assert rel(Yb.cpu().numpy(), O.matvec(Ao, Xb)) <= 1e-12


def test_destroy_right_after_async_product(tb):
    """`mul!` on a matrix type factorizes, multiplies asynchronously and drops the factorization at once (the Julia
    finalizer does the same); the release must be ordered after the product even on a non-default stream, while other
    handles are created and used from the pool in between."""
    n, l = 1 << 16, 4
    rng = np.random.default_rng(3)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    Ao = O.Toeplitz(vc, vr)
    A = tb.Toeplitz(torch.from_numpy(vc), torch.from_numpy(vr))
    X = rng.standard_normal((n, l))
    Xd = devmat(X)
    ref = O.matvec(Ao, X)
    s1 = torch.cuda.Stream()
    outs = []
    torch.cuda.synchronize()
    for i in range(12):
        Y = tb.colmajor(n, l)
        with torch.cuda.stream(s1):
            tb.mul_(Y, A, Xd)                         # temporary factorization destroyed on return
        # churn the pool from the default stream: same-size spectra / workspaces are handed out again
        F2 = tb.factorize(tb.SymmetricToeplitz(torch.from_numpy(vc)))
        tb.mul(F2, Xd[:, :2])
        del F2
        outs.append(Y)
    torch.cuda.synchronize()
    for Y in outs:
        assert rel(Y.cpu().numpy(), ref) <= 1e-12


def test_mul_host_waits_for_async_factorize(tb):
    n, l = 1 << 18, 6
    rng = np.random.default_rng(5)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    X = np.asfortranarray(rng.standard_normal((n, l)))
    Xh = torch.from_numpy(np.ascontiguousarray(X.T)).t()
    Yh = torch.empty((l, n), dtype=torch.float64).t()
    s1 = torch.cuda.Stream()
    vcd, vrd = torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        # a long queue in front of the factorization so that it is certainly unfinished when mul_host_ starts
        junk = torch.randn(1 << 26, device="cuda")
        for _ in range(10):
            junk = junk * 1.0001
        F = tb.factorize(tb.Toeplitz(vcd, vrd))
    tb.mul_host_(Yh, F, Xh)
    assert rel(Yh.numpy(), O.matvec(O.Toeplitz(vc, vr), X)) <= 1e-12


def test_comm_single_device_group(tb):
    """tb200_comm_init / tb200_bcast_spectrum / tb200_comm_destroy on a one-device communicator: the broadcast is a
    self-copy, the handle's readiness event is recorded, products still match."""
    from toeplitz_b200 import sharding
    n = 4096
    rng = np.random.default_rng(11)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc), torch.from_numpy(vr)))
    comm = sharding.Communicator.all_local(1)
    assert comm.size() == (1, 1)
    comm.bcast_spectrum([F], root=0)
    x = rng.standard_normal(n)
    y = tb.mul(F, torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel(y, O.matvec(O.Toeplitz(vc, vr), x)) <= 1e-12
    with pytest.raises(tb.ArgumentError):
        comm.bcast_spectrum([F, F], root=0)
    comm.destroy()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_comm_two_devices_single_process(tb):
    """One process driving two GPUs (the Julia host's usual mode): factorize on device 0, empty handle on device 1,
    tb200_bcast_spectrum, then each device multiplies its own column shard."""
    from toeplitz_b200 import sharding
    n, l = 1 << 15, 8
    rng = np.random.default_rng(13)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    X = rng.standard_normal((n, l))
    ref = O.matvec(O.Toeplitz(vc, vr), X)
    with torch.cuda.device(0):
        F0 = tb.factorize(tb.Toeplitz(torch.from_numpy(vc).cuda(0), torch.from_numpy(vr).cuda(0)))
    with torch.cuda.device(1):
        F1 = tb.factorize_empty_like(0, torch.float64, (n, n))
    comm = sharding.Communicator.all_local(2)
    comm.bcast_spectrum([F0, F1], root=0)
    shards = sharding.all_shards(l, 2)
    outs = []
    for d, F in enumerate((F0, F1)):
        a, b = shards[d]
        with torch.cuda.device(d):
            Xd = torch.from_numpy(np.ascontiguousarray(X[:, a:b].T)).cuda(d).t()
            outs.append(tb.mul(F, Xd))
    Y = np.concatenate([o.cpu().numpy() for o in outs], axis=1)
    assert rel(Y, ref) <= 1e-12
    comm.destroy()


@pytest.mark.parametrize("n", [300, 5000, 70000])
def test_unpaired_real_columns(tb, n):
    """`pair_columns = 0`: every real column is transformed alone, as in the reference's column loop
    (src/linearalgebra.jl:95-97) -- a NaN column or a column 1e100 times larger leaves its neighbours untouched.
    With the default pairing the product is the same to 1e-12 on ordinary columns (checked here as well)."""
    rng = np.random.default_rng(n)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    Ao = O.Toeplitz(vc, vr)
    F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()))
    X = rng.standard_normal((n, 5))
    Yo = np.stack([O.matvec(Ao, X[:, j]) for j in range(5)], axis=1)
    Yp = tb.mul(F, devmat(X)).cpu().numpy()
    tb.set_option("pair_columns", 0)
    try:
        Yu = tb.mul(F, devmat(X)).cpu().numpy()
        assert rel(Yu, Yo) <= 1e-12 and rel(Yp, Yo) <= 1e-12
        Xbad = X.copy()
        Xbad[7, 0] = np.nan                      # partner of column 1 under pairing
        Xbad[:, 2] *= 1e100                      # partner of column 3
        Yb = tb.mul(F, devmat(Xbad)).cpu().numpy()
        assert np.isnan(Yb[:, 0]).all() or np.isnan(Yb[:, 0]).any()
        for j in (1, 3, 4):
            assert np.isfinite(Yb[:, j]).all()
            assert rel(Yb[:, j], Yo[:, j]) <= 1e-12
        assert rel(Yb[:, 2], 1e100 * Yo[:, 2]) <= 1e-12
    finally:
        tb.set_option("pair_columns", 1)


@pytest.mark.parametrize("n,cols", [(3000, 1), (1 << 16, 1), (1 << 16, 24), (1 << 20, 6)])
def test_products_are_graph_capturable(tb, n, cols):
    """After one warm-up product on the capturing stream (workspaces allocated, the factorization seen complete) a
    product enqueues kernels and event fork/joins only, so it can be captured into a CUDA graph and replayed on new
    contents of the same buffers; replay results are bit-identical to the eager call."""
    rng = np.random.default_rng(n + cols)
    vc, vr = 0.9 ** np.arange(n), 0.4 ** np.arange(n)
    F = tb.factorize(tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()))   # on the default stream
    shape = (n,) if cols == 1 else (cols, n)
    x = torch.from_numpy(rng.standard_normal(shape)).cuda()
    x = x if cols == 1 else x.t()
    y = torch.empty(n, dtype=torch.float64, device="cuda") if cols == 1 else tb.colmajor(n, cols, torch.float64)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        tb.mul_(y, F, x)                         # warm-up on the capturing stream
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            tb.mul_(y, F, x)
    x.copy_(torch.from_numpy(rng.standard_normal(tuple(x.shape))).cuda())
    y.zero_()
    g.replay()
    torch.cuda.synchronize()
    got = y.clone()
    y.zero_()
    tb.mul_(y, F, x)
    torch.cuda.synchronize()
    assert torch.equal(got, y)
    xo = x.cpu().numpy() if cols == 1 else x[:, 0].cpu().numpy()
    yo = got.cpu().numpy() if cols == 1 else got[:, 0].cpu().numpy()
    assert rel(yo, O.matvec(O.Toeplitz(vc, vr), xo)) <= 1e-12
