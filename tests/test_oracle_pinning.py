"""Pin the CPU oracle against every deterministic assertion the reference's own
test-suite makes for the hot path (/root/reference/test/runtests.jl), restated.

The reference has no golden vectors; its numeric bar is
``isapprox`` with ``rtol = sqrt(eps)`` against the dense matrix.  RNG-drawn
inputs of the reference (Julia's version-specific stream) are replaced by
numpy PCG64 draws with fixed seeds.
"""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import toeplitz_oracle as O
from oracle import build_oracle as OC

RTOL = np.sqrt(np.finfo(np.float64).eps)
ns, nl = 101, 2000
rng = np.random.default_rng(0)
xs = rng.standard_normal((ns, 5))
xl = rng.standard_normal((nl, 5))


def approx(a, b, rtol=RTOL):
    a = np.asarray(a)
    b = np.asarray(b)
    return np.linalg.norm((a - b).ravel()) <= rtol * max(np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel()))


def p(base, n, cplx=False):
    v = base ** np.arange(n)
    return v.astype(np.complex128) if cplx else v


CASES = [  # test/runtests.jl:23-48
    ("real general", lambda n: O.Toeplitz(p(0.9, n), p(0.4, n))),
    ("complex general", lambda n: O.Toeplitz(p(0.9, n, True), p(0.4, n, True))),
    ("real circulant", lambda n: O.Circulant(p(0.9, n))),
    ("complex circulant", lambda n: O.Circulant(p(0.9, n, True))),
    ("real upper", lambda n: O.UpperTriangularToeplitz(p(0.9, n))),
    ("complex upper", lambda n: O.UpperTriangularToeplitz(p(0.9, n, True))),
    ("real lower", lambda n: O.LowerTriangularToeplitz(p(0.9, n))),
    ("complex lower", lambda n: O.LowerTriangularToeplitz(p(0.9, n, True))),
]


@pytest.mark.parametrize("name,mk", CASES, ids=[c[0] for c in CASES])
def test_mul_adjoint_ldiv_vs_dense(name, mk):
    """test/runtests.jl:50-63."""
    for n, x in ((ns, xs), (nl, xl)):
        A = mk(n)
        M = A.dense()
        assert approx(O.matvec(A, x), M @ x)
        assert approx(O.matvec(A.adjoint(), x), M.conj().T @ x)
        assert np.array_equal(A.adjoint().dense(), M.conj().T)
        assert np.array_equal(A.transpose().dense(), M.T)
        B = np.asfortranarray(x.astype(A.dtype))
        O.ldiv_mat(A, B, O.ldiv)
        assert approx(B, np.linalg.solve(M, x))


def test_mixed_types_exact():
    """test/runtests.jl:74-81."""
    A = O.Toeplitz(np.array([1, 2]), np.array([1, 2]))
    assert A.dtype.kind == "i"
    assert np.array_equal(O.matvec(A, np.ones(2)), np.full(2, 3.0))
    C = O.Circulant(np.arange(1, 4, dtype=np.float32))
    assert np.array_equal(O.matvec(C, np.ones(3)), np.full(3, 6.0))


@pytest.mark.parametrize("cplx", [False, True])
def test_rectangular(cplx):
    """test/runtests.jl:83-95."""
    A1 = O.Toeplitz(p(0.9, nl, cplx), p(0.4, ns, cplx))
    A2 = O.Toeplitz(p(0.9, ns, cplx), p(0.4, nl, cplx))
    assert approx(O.matvec(A1, xs), A1.dense() @ xs)
    assert approx(O.matvec(A2, xl), A2.dense() @ xl)


def test_symmetric_toeplitz():
    """test/runtests.jl:97-110."""
    As = O.SymmetricToeplitz(p(0.9, ns))
    Ab = O.SymmetricToeplitz(np.abs(np.random.default_rng(1).standard_normal(ns)) + 0.0)
    Al = O.SymmetricToeplitz(p(0.9, nl))
    assert np.array_equal(As.dense(), As.dense().T)
    for A, x in ((As, xs), (Ab, xs), (Al, xl)):
        assert approx(O.matvec(A, x), A.dense() @ x)
    for A, x in ((As, xs), (Al, xl)):
        B = np.asfortranarray(x.copy())
        O.ldiv_mat(A, B, O.ldiv)
        assert approx(B, np.linalg.solve(A.dense(), x))
    # StatsBase.levinson(As, xs) ~ dense solve  (:107)  -- our own levinson on columns
    for j in range(xs.shape[1]):
        assert approx(O.levinson(As, xs[:, j].copy()), np.linalg.solve(As.dense(), xs[:, j]))
        assert approx(OC.c_levinson_sym(As.vc, xs[:, j]), np.linalg.solve(As.dense(), xs[:, j]))


def _levinson_kernels():
    n = 8
    x = np.linspace(-1, 1, n + 1)
    a_exp = np.exp(-np.abs(x[0] - x))
    a_rbf = np.exp(-np.abs(x[0] - x) ** 2) / 2
    a_rand = np.random.default_rng(2).random(n + 1) / n
    return n, (a_rand, a_exp, a_rbf)


def test_durbin_trench_levinson_n8():
    """test/runtests.jl:112-151."""
    n, kernels = _levinson_kernels()
    g = np.random.default_rng(3)
    for a in kernels:
        a = a.copy()
        a[0] = 1
        r = a[1:]
        T = O.SymmetricToeplitz(np.concatenate([[1.0], r[:-1]]))
        TM = T.dense()
        scale = 2 + g.standard_normal() / 10
        TS = O.SymmetricToeplitz(scale * T.vc)
        TSM = TS.dense()
        # 1. durbin
        y = O.durbin(r)
        assert approx(-np.linalg.solve(TM, r), y)
        assert approx(OC.c_durbin(r), y, 1e-15)
        # 2. trench
        invTM = np.linalg.inv(TM)
        assert approx(O.trench(r[:-1]), invTM)
        assert approx(O.trench(T), invTM)
        assert approx(O.trench(TS), np.linalg.inv(TSM))
        # 3. levinson
        b = g.standard_normal(n)
        yl = O.levinson(r[:-1], b)
        Tb = np.linalg.solve(TM, b)
        assert approx(Tb, yl)
        assert approx(Tb, O.levinson(T, b))
        assert approx(np.linalg.solve(TSM, b), O.levinson(TS, b))
        assert approx(OC.c_levinson(r[:-1], b), yl, 1e-15)
        assert approx(OC.c_levinson_sym(TS.vc, b), O.levinson(TS, b), 1e-15)


def test_levinson_vs_scipy_solve_toeplitz():
    g = np.random.default_rng(4)
    n = 64
    r = g.random(n - 1) / n
    b = g.standard_normal(n)
    ref = sla.solve_toeplitz(np.concatenate([[1.0], r]), b)
    assert approx(O.levinson(r, b), ref, 1e-12)
    R = np.asfortranarray(np.stack([r, r * 0.5], axis=1))
    B = np.asfortranarray(np.stack([b, b[::-1]], axis=1))
    X = OC.c_levinson_batched(R, B)
    assert approx(X[:, 0], ref, 1e-12)
    assert approx(X[:, 1], sla.solve_toeplitz(np.concatenate([[1.0], r * 0.5]), b[::-1]), 1e-12)
    Y = OC.c_durbin_batched(np.asfortranarray(np.stack([np.append(r, 0.01)], axis=1)))
    assert approx(Y[:, 0], O.durbin(np.append(r, 0.01)), 1e-15)


def test_hankel_literal_and_products():
    """test/runtests.jl:154-215."""
    H = O.Hankel(np.array([1.0, 2, 3, 4, 5]), vr=np.array([5.0, 6, 7, 8, 9]))
    assert np.array_equal(H.dense(), np.array([[1, 2, 3, 4, 5], [2, 3, 4, 5, 6], [3, 4, 5, 6, 7],
                                               [4, 5, 6, 7, 8], [5, 6, 7, 8, 9]], dtype=float))
    x = np.ones(5)
    assert approx(O.mul_vec(x.copy(), H, x), H.dense() @ x)
    for cplx in (False, True):
        c = (lambda v: v.astype(np.complex128)) if cplx else (lambda v: v)
        Hs = O.Hankel(c(0.9 ** np.arange(ns - 1, -1, -1)), vr=c(0.4 ** np.arange(ns)))
        Hl = O.Hankel(c(0.9 ** np.arange(nl - 1, -1, -1)), vr=c(0.4 ** np.arange(nl)))
        assert approx(O.matvec(Hs, xs[:, 0]), Hs.dense() @ xs[:, 0])
        assert approx(O.matvec(Hs, xs), Hs.dense() @ xs)
        assert approx(O.matvec(Hl, xl), Hl.dense() @ xl)
        # rectangular
        Hr1 = O.Hankel(c(0.9 ** np.arange(ns - 1, -1, -1)), vr=c(0.4 ** np.arange(nl)))
        Hr2 = O.Hankel(c(0.9 ** np.arange(nl - 1, -1, -1)), vr=c(0.4 ** np.arange(ns)))
        assert approx(O.matvec(Hr1, xl[:, 0]), Hr1.dense() @ xl[:, 0])
        assert approx(O.matvec(Hr1, xl), Hr1.dense() @ xl)
        assert approx(O.matvec(Hr2, xs), Hr2.dense() @ xs)


def test_circulant_mathematics():
    """test/runtests.jl:501-596 (the spectral subset)."""
    g = np.random.default_rng(5)
    C1 = O.Circulant(g.random(5))
    C2 = O.Circulant(g.random(5))
    C3 = O.Circulant(g.random(5).astype(np.complex128))
    C4 = O.Circulant(np.ones(5))
    C5 = O.Circulant(np.ones(5, dtype=np.complex128))
    M1, M2, M3 = C1.dense(), C2.dense(), C3.dense()
    adj = lambda F: F.adjoint()
    ident = lambda F: F
    for t1, mt1 in ((ident, ident), (adj, lambda M: M.conj().T)):
        for t2, mt2 in ((ident, ident), (adj, lambda M: M.conj().T)):
            C = O.circ_mul(t1(O.factorize(C1)), t2(O.factorize(C2)))
            assert isinstance(C, O.Circulant)
            assert approx(C.dense(), mt1(M1) @ mt2(M2))
    for C, M in ((C1, M1), (C3, M3)):
        Ci = O.circ_inv(C)
        assert approx(Ci.dense(), np.linalg.inv(M))
        assert approx(np.fft.fft(Ci.vc), O.factorize(Ci).vcvr_dft)          # :543,547
        Cs = O.circ_sqrt(C)
        assert approx(O.circ_mul(O.factorize(Cs), O.factorize(Cs)).dense(), M)
        assert approx(O.circ_det(C), np.linalg.det(M))
    for C in (C1, C3, C4, C5):
        Cp = O.circ_pinv(C)
        assert approx(Cp.dense(), np.linalg.pinv(C.dense()))
        assert approx(np.fft.fft(Cp.vc), O.factorize(Cp).vcvr_dft)
    v1 = O.circ_eigvals(C1)
    v2 = np.linalg.eigvals(M1)
    for v in v1:
        assert np.min(np.abs(v - v2)) < RTOL
    # issue #47
    e = g.random(5)
    I1 = O.circ_mul(O.factorize(O.circ_inv(C1)), O.factorize(C1))
    I2 = O.circ_mul(O.circ_inv_factorization(O.factorize(C1)), O.factorize(C1))
    assert approx(O.matvec(I1, e), e) and approx(O.matvec(I2, e), e)


def test_issue73_identity_circulant():
    """test/runtests.jl:794-803."""
    b = np.random.default_rng(6).random(6)
    P = O.Circulant(np.array([1.0, 0, 0, 0, 0, 0]))
    F = O.factorize(P)
    assert F.shape == P.shape
    x = b.copy()
    O.circ_ldiv_factorization(F, x)
    assert approx(x, b)


def test_dimension_errors():
    A = O.Toeplitz(p(0.9, 600), p(0.4, 600))
    with pytest.raises(O.DimensionMismatch):
        O.mul_vec(np.zeros(599), A, np.zeros(600))
    with pytest.raises(O.DimensionMismatch):
        O.mul_vec(np.zeros(600), A, np.zeros(601))
    with pytest.raises(O.DimensionMismatch):
        O.circ_ldiv_factorization(O.factorize(O.Circulant(np.ones(4))), np.zeros(5))
    with pytest.raises(RuntimeError):
        O.ldiv_mat(O.Toeplitz(p(0.9, 4), p(0.4, 3)), np.zeros((4, 1)), O.ldiv)
    with pytest.raises(ValueError):
        O.Toeplitz(np.array([1.0, 2]), np.array([2.0, 1]))


def test_cgs_flags_and_pad_invariance():
    # cgs returns the 4-tuple; converged flag on a well-conditioned KMS system
    n = 600
    A = O.SymmetricToeplitz(0.5 ** np.arange(n))
    b = np.random.default_rng(7).standard_normal(n)
    x, err, it, flag = O.cgs(A, np.zeros(n), b.copy(), O.factorize(O.strang(A)), 1000, 100 * np.finfo(float).eps)
    assert flag == 0 and err <= 100 * np.finfo(float).eps and 1 <= it < 50
    assert approx(x, np.linalg.solve(A.dense(), b), 1e-10)
    # zero-padded power-of-two embedding gives the same product
    T = O.Toeplitz(p(0.9, 700), p(0.4, 700))
    xx = np.random.default_rng(8).standard_normal(700)
    Np = 2048
    c = np.zeros(Np, dtype=complex)
    c[:700] = T.vc
    c[Np - 699:] = T.vr[:0:-1]
    xp = np.zeros(Np, dtype=complex)
    xp[:700] = xx
    yp = np.fft.ifft(np.fft.fft(xp) * np.fft.fft(c)).real[:700]
    assert approx(yp, O.matvec(T, xx), 1e-13)


def test_triangular_inverse_recursive_doubling():
    """src/linearalgebra.jl:337-365 restated; test/runtests.jl:637-639,656-662 (lower branch) vs the dense inverse
    on a well-conditioned decaying column."""
    for n in (2, 64, 65, 128, 200):
        v = np.concatenate([[1.0], 0.5 * 0.7 ** np.arange(1, n)])
        Linv = O.tri_inv_lower(v)
        dense = np.linalg.inv(O.LowerTriangularToeplitz(v).dense())
        assert approx(O.LowerTriangularToeplitz(Linv).dense(), dense, 1e-10)
