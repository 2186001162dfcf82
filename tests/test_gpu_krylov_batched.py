"""Batched cgs / cg (all columns of ldiv!(A, B::Matrix) iterating together, src/linearalgebra.jl:102-110) against the
oracle's per-column solves (src/iterativeLinearSolvers.jl:9-215): same iterates, iteration counts and flags per column,
including columns that converge at different iterations, zero columns (early return) and a fixed iteration budget."""
import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O
import parity_tools as PT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _devmat(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def _rel(a, b):
    den = np.linalg.norm(np.asarray(b).ravel())
    return np.linalg.norm((np.asarray(a) - np.asarray(b)).ravel()) / (den if den > 0 else 1.0)


def _p(base, n, dtype):
    return (base ** np.arange(n)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.complex64], ids=lambda d: np.dtype(d).name)
def test_cgs_batched_matches_per_column_oracle(tb, dtype):
    n, l = 3000, 7
    rng = np.random.default_rng(21)
    single = np.dtype(dtype) == np.dtype(np.complex64)
    v = _p(0.5, n, dtype)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(_dev(v))
    B = rng.standard_normal((n, l)).astype(dtype)
    B[:, 2] = 0                                  # early return: error 0 < tol at the start
    B[:, 4] = Ao.dense()[:, 0]                   # solved almost at once: converges well before the others
    tol = 1e-5 if single else 100 * np.finfo(np.float64).eps
    Mo, Md = O.factorize(O.strang(Ao)), tb.factorize(tb.strang(Ad))
    Dn = Ao.dense().astype(np.complex128)
    kA = PT.cond2(Dn)                            # KMS(0.5): ~9
    X = tb.colmajor(n, l, _dev(B[:1]).dtype)
    X.zero_()
    X, errs, its, flags = tb.cgs_batched(Ad, X, _devmat(B), Md, 50, tol)
    Xh = X.cpu().numpy()
    seen = set()
    for j in range(l):
        xo, eo, ito, fo = O.cgs(Ao, np.zeros(n, dtype), B[:, j].copy(), Mo, 50, tol)
        assert flags[j] == fo and abs(its[j] - ito) <= 1, (j, its[j], ito, flags[j], fo)
        # both sides stop at a relative residual <= tol: the solutions agree to cond * (res_o + res_d)
        PT.check("cgs batched col %d" % j, Xh[:, j], xo, dtype, kA, extra=2.0 * kA * (eo + errs[j]))
        assert errs[j] <= max(10 * eo, tol)
        seen.add(its[j])
    assert its[2] == 0 and len(seen) >= 3        # columns really stopped at different iterations
    # fixed budget without convergence: every column runs exactly 3 iterations and reports flag 1
    X2 = tb.colmajor(n, l, X.dtype)
    X2.zero_()
    Bn = rng.standard_normal((n, l)).astype(dtype)
    X2, errs2, its2, flags2 = tb.cgs_batched(Ad, X2, _devmat(Bn), Md, 3, 0.0)
    for j in range(l):
        xo, eo, ito, fo = O.cgs(Ao, np.zeros(n, dtype), Bn[:, j].copy(), Mo, 3, 0.0)
        assert (its2[j], flags2[j]) == (ito, fo) == (3, 1)
        # the budget outlasts convergence: iterates agree to cond * (true residual floor), measured on both sides
        xd = X2[:, j].cpu().numpy()
        nb = np.linalg.norm(Bn[:, j])
        floor = (np.linalg.norm(Bn[:, j] - Dn @ xo) + np.linalg.norm(Bn[:, j] - Dn @ xd)) / nb
        PT.check("cgs batched 3 iterations col %d" % j, xd, xo, dtype, kA, extra=2.0 * kA * floor)


def test_cg_batched_matches_per_column_oracle(tb):
    n, l = 2500, 6
    rng = np.random.default_rng(22)
    v = _p(0.5, n, np.float64)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(_dev(v))
    B = rng.standard_normal((n, l))
    B[:, 1] = 0                                  # the reference's value-less early return -> flag 2
    tol = 100 * np.finfo(np.float64).eps
    Mo, Md = O.factorize(O.strang(Ao)), tb.factorize(tb.strang(Ad))
    kA = PT.cond2(Ao.dense())
    X = tb.colmajor(n, l)
    X.zero_()
    X, errs, its, flags = tb.cg_batched(Ad, X, _devmat(B), Md, 60, tol)
    for j in range(l):
        res = O.cg(Ao, np.zeros(n), B[:, j].copy(), Mo, 60, tol)
        if res is None:
            assert flags[j] == 2 and its[j] == 0
            continue
        xo, eo, ito, fo = res
        assert flags[j] == fo and abs(its[j] - ito) <= 1
        PT.check("cg batched col %d" % j, X[:, j].cpu().numpy(), xo, np.float64, kA, extra=2.0 * kA * (eo + errs[j]))
    with pytest.raises(tb.ArgumentError):
        tb.cg_batched(Ad, X.to(torch.complex128), _devmat(B.astype(np.complex128)), Md, 5, tol)


@pytest.mark.parametrize("kind", ["toeplitz", "symmetric", "lower", "upper"])
def test_matrix_ldiv_goes_through_the_batched_solvers(tb, kind):
    """ldiv!(A, B::Matrix) (test/runtests.jl:58-59,103-105) with enough columns to matter, odd count, padded ld."""
    n, l = 1500, 9
    rng = np.random.default_rng(23)
    if kind == "toeplitz":
        vc, vr = _p(0.9, n, np.float64), _p(0.4, n, np.float64)
        Ao, Ad = O.Toeplitz(vc, vr), tb.Toeplitz(_dev(vc), _dev(vr))
    else:
        v = _p(0.9, n, np.float64)
        cls = {"symmetric": "SymmetricToeplitz", "lower": "LowerTriangularToeplitz", "upper": "UpperTriangularToeplitz"}[kind]
        Ao, Ad = getattr(O, cls)(v), getattr(tb, cls)(_dev(v))
    B = rng.standard_normal((n, l))
    ref = np.linalg.solve(Ao.dense(), B)
    big = torch.full((l, n + 5), float("nan"), dtype=torch.float64, device="cuda")
    big[:, :n] = _dev(B.T.copy())
    Bd = big.t()[:n, :]
    l0 = tb.launch_count()
    tb.ldiv_(Ad, Bd)
    assert _rel(Bd.cpu().numpy(), ref) <= np.sqrt(np.finfo(float).eps)
    assert tb.launch_count() - l0 < 60 * 40      # far fewer launches than nine separate solves
    assert bool(torch.isnan(big[:, n:]).all())   # padding untouched


def test_batched_chunking_and_errors(tb):
    n, l = 1024, 11
    rng = np.random.default_rng(24)
    v = _p(0.5, n, np.float64)
    Ad = tb.SymmetricToeplitz(_dev(v))
    Md = tb.factorize(tb.strang(Ad))
    B = rng.standard_normal((n, l))
    tol = 100 * np.finfo(np.float64).eps
    X1 = tb.colmajor(n, l); X1.zero_()
    X1, e1, i1, f1 = tb.cgs_batched(Ad, X1, _devmat(B), Md, 40, tol)
    tb.set_option("krylov_work_bytes", 8 * n * 8 * 4)        # four columns per chunk -> 3 chunks, the last one odd
    try:
        X2 = tb.colmajor(n, l); X2.zero_()
        X2, e2, i2, f2 = tb.cgs_batched(Ad, X2, _devmat(B), Md, 40, tol)
    finally:
        tb.set_option("krylov_work_bytes", 16 << 30)
    assert i1 == i2 and f1 == f2 and _rel(X2.cpu().numpy(), X1.cpu().numpy()) <= 1e-13
    with pytest.raises(tb.DimensionMismatch):
        tb.cgs_batched(Ad, tb.colmajor(n, l), _devmat(B[:, :3].copy()), Md, 5, tol)
    # an empty iteration budget: the reference falls through its flag logic with rho == 0 (cgs: -1) / error > tol (cg: 1)
    X0 = tb.colmajor(n, l); X0.zero_()
    _, e0, i0, f0 = tb.cgs_batched(Ad, X0, _devmat(B), Md, 0, tol)
    xo, eo, ito, fo = O.cgs(O.SymmetricToeplitz(v), np.zeros(n), B[:, 0].copy(), O.factorize(O.strang(O.SymmetricToeplitz(v))), 0, tol)
    assert set(f0) == {fo} == {-1} and set(i0) == {ito} == {0} and abs(e0[0] - eo) <= 1e-12
    _, _, i0, f0 = tb.cg_batched(Ad, X0, _devmat(B), Md, 0, tol)
    assert set(f0) == {1} and set(i0) == {0}


@pytest.mark.parametrize("n,l", [(6000, 7), (40000, 5)])
def test_batched_two_level_transforms_with_stragglers(tb, n, l):
    """Embeddings beyond one shared-memory transform (two-level path, several column groups per product): the column
    selection and the per-column alpha of the fused residual update must address the same user columns in every group."""
    rng = np.random.default_rng(25)
    vc, vr = _p(0.9, n, np.float64), _p(0.4, n, np.float64)
    Ao, Ad = O.Toeplitz(vc, vr), tb.Toeplitz(_dev(vc), _dev(vr))
    B = rng.standard_normal((n, l))
    B[:, 1] = 0
    B[:, 3] = Ao.dense()[:, 0] if n <= 8000 else O.matvec(Ao, np.eye(1, n, 0).ravel())
    tol = 100 * np.finfo(np.float64).eps
    Mo, Md = O.factorize(O.strang(Ao)), tb.factorize(tb.strang(Ad))
    # Toeplitz(0.9^k, 0.4^k): symbol f(z) = 1/(1-0.9z) + 0.4/(z-0.4); cond <= max|f| / min|f| on |z| = 1 (computed on a grid)
    z = np.exp(2j * np.pi * np.arange(4096) / 4096)
    f = 1.0 / (1.0 - 0.9 * z) + 0.4 / (z - 0.4)
    kA = float(np.abs(f).max() / np.abs(f).min())
    tb.set_option("group_bytes", 1 << 20)                     # one pair per group -> several groups per product
    try:
        X = tb.colmajor(n, l)
        X.zero_()
        X, errs, its, flags = tb.cgs_batched(Ad, X, _devmat(B), Md, 12, tol)
    finally:
        tb.set_option("group_bytes", 48 << 20)
    Xh = X.cpu().numpy()
    for j in range(l):
        xo, eo, ito, fo = O.cgs(Ao, np.zeros(n), B[:, j].copy(), Mo, 12, tol)
        assert abs(its[j] - ito) <= 1, (j, its[j], ito)
        if flags[j] == fo == 0:
            PT.check("cgs batched two-level n=%d col %d" % (n, j), Xh[:, j], xo, np.float64, kA, extra=2.0 * kA * (eo + errs[j]))
        # a column that has not converged inside the budget still reports a finite residual of the oracle's order
        assert np.isfinite(errs[j]) and errs[j] <= max(1e3 * eo, 1e3 * tol), (j, errs[j], eo)
    # single-vector solver == one-column batched solver
    x1, e1, i1, f1 = tb.cgs(Ad, torch.zeros(n, dtype=torch.float64, device="cuda"), _dev(B[:, 0].copy()), Md, 12, tol)
    assert i1 == its[0] and f1 == flags[0] and _rel(x1.cpu().numpy(), Xh[:, 0]) <= 1e-12


@pytest.mark.parametrize("solver", ["cgs", "cg"])
def test_nan_column_does_not_reach_its_pair(tb, solver):
    """Real columns share complex transforms two by two; a column whose residual is NaN is frozen at once (with what
    the reference's remaining iterations would report: iter = max_it, x all NaN, flag 1 for cgs, and 0 for cg whose
    final `error > tol` is false for NaN, src/iterativeLinearSolvers.jl:95,205-213), so its partner's iterates are
    those of the reference's independent column loop (src/linearalgebra.jl:102-110)."""
    n, l = 3000, 4
    rng = np.random.default_rng(5)
    v = _p(0.5, n, np.float64)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(_dev(v))
    Mo, Md = O.factorize(O.strang(Ao)), tb.factorize(tb.strang(Ad))
    kA = PT.cond2(Ao.dense())
    B = rng.standard_normal((n, l))
    B[17, 0] = np.nan                            # column 0 pairs with column 1
    tol = 100 * np.finfo(np.float64).eps
    X = tb.colmajor(n, l, torch.float64)
    X.zero_()
    run = tb.cgs_batched if solver == "cgs" else tb.cg_batched
    ref = O.cgs if solver == "cgs" else O.cg
    X, errs, its, flags = run(Ad, X, _devmat(B), Md, 40, tol)
    Xh = X.cpu().numpy()
    assert np.isnan(Xh[:, 0]).all() and np.isnan(errs[0]) and its[0] == 40
    assert flags[0] == (1 if solver == "cgs" else 0)
    for j in range(1, l):
        xo, eo, ito, fo = ref(Ao, np.zeros(n), B[:, j].copy(), Mo, 40, tol)
        assert np.isfinite(Xh[:, j]).all()
        assert flags[j] == fo and abs(its[j] - ito) <= 1
        PT.check("%s batched beside a NaN column, col %d" % (solver, j), Xh[:, j], xo, np.float64, kA, extra=2.0 * kA * (eo + errs[j]))


import os

_KSCALE = max(1, int(os.environ.get("TB200_FUZZ_SCALE", "1")))
_KSEED0 = int(os.environ.get("TB200_FUZZ_SEED0", "0"))


@pytest.mark.parametrize("seed", range(_KSEED0, _KSEED0 + 6 * _KSCALE))
def test_batched_krylov_random_batches(tb, seed):
    """Seeded sweep of the batched drivers against the oracle's column loop: random order (one shared-memory transform
    or the two-level path, forced into several column groups), column count, element type, preconditioner (Strang /
    T. Chan), chunking of the work vectors, and a random mix of column kinds -- zero columns (early return),
    a column solved at once, columns 1e6 times larger or smaller than their pair partner, ordinary ones -- so that columns
    leave the active set at different iterations.  Per column: same flag, iteration count within one, solution within
    cond * (both residuals)."""
    rng = np.random.default_rng(7000 + seed)
    n = int(np.exp(rng.uniform(np.log(300), np.log(30000))))
    l = int(rng.integers(1, 11))
    use_cg = rng.random() < 0.35
    dtype = np.float64 if use_cg else [np.float64, np.complex128, np.complex64][int(rng.integers(3))]
    single = np.dtype(dtype) == np.dtype(np.complex64)
    rho = float(rng.uniform(0.2, 0.7))
    v = _p(rho, n, dtype)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(_dev(v))
    kA = ((1.0 + rho) / (1.0 - rho)) ** 2          # KMS symbol (1 - rho^2) / |1 - rho z|^2: max / min on the unit circle
    pc = ["strang", "chan"][int(rng.integers(2))]
    Mo = O.factorize(getattr(O, pc)(Ao))
    Md = tb.factorize(getattr(tb, pc)(Ad))
    B = rng.standard_normal((n, l)).astype(dtype)
    if np.dtype(dtype).kind == "c":
        B = (B + 1j * rng.standard_normal((n, l))).astype(dtype)
    kinds = []
    e0 = np.zeros(n, dtype); e0[0] = 1
    for j in range(l):
        k = int(rng.integers(6))
        kinds.append(k)
        if k == 0:
            B[:, j] = 0
        elif k == 1:
            B[:, j] = O.matvec(Ao, e0)
        elif k == 2:
            B[:, j] *= 1e6
        elif k == 3:
            B[:, j] *= 1e-6
    tol = 2e-5 if single else 1e-13
    max_it = int(rng.integers(3, 40))
    tb.set_option("group_bytes", 1 << 20)
    tb.set_option("krylov_work_bytes", int(rng.choice([16 << 30, 8 * n * np.dtype(dtype).itemsize * 3, 8 * n * np.dtype(dtype).itemsize * 5])))
    try:
        X = tb.colmajor(n, l, _dev(B[:1]).dtype)
        X.zero_()
        run = tb.cg_batched if use_cg else tb.cgs_batched
        X, errs, its, flags = run(Ad, X, _devmat(B), Md, max_it, tol)
    finally:
        tb.set_option("group_bytes", 48 << 20)
        tb.set_option("krylov_work_bytes", 16 << 30)
    Xh = X.cpu().numpy()
    tag = "seed %d: %s n=%d l=%d %s pc=%s max_it=%d kinds=%s" % (seed, "cg" if use_cg else "cgs", n, l, np.dtype(dtype).name, pc,
                                                              max_it, kinds)
    for j in range(l):
        res = (O.cg if use_cg else O.cgs)(Ao, np.zeros(n, dtype), B[:, j].copy(), Mo, max_it, tol)
        if res is None:
            assert flags[j] == 2 and its[j] == 0, tag
            continue
        xo, eo, ito, fo = res
        near_tol = min(eo, errs[j]) <= 4 * tol and max(eo, errs[j]) >= tol / 4      # stopped on different sides of tol
        assert abs(its[j] - ito) <= 1, (tag, j, its[j], ito)
        assert flags[j] == fo or near_tol, (tag, j, flags[j], fo, errs[j], eo)
        if np.linalg.norm(xo) > 0:
            PT.check("krylov fuzz %s col %d" % (tag, j), Xh[:, j], xo, dtype, kA, extra=2.0 * kA * (eo + errs[j]))
        else:
            assert np.linalg.norm(Xh[:, j]) == 0, tag
