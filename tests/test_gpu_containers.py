"""Container algebra on device vectors against dense numpy, the way the reference tests it
(test/runtests.jl:340-420: ``TA+TB == A+B``, ``tril(TA,k) == tril(A,k)``, ``isa(reverse(TA), Hankel)`` ...)."""
import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _mk(tb, kind, rng, n, dtype, m=None):
    c = (rng.standard_normal(n + (m or n)) + (1j * rng.standard_normal(n + (m or n)) if np.dtype(dtype).kind == "c" else 0)).astype(dtype)
    if kind == "toeplitz":
        vc, vr = c[: (m or n)].copy(), c[(m or n):].copy()
        vr[0] = vc[0]
        return O.Toeplitz(vc, vr), tb.Toeplitz(_dev(vc), _dev(vr))
    v = c[:n].copy()
    cls = {"symmetric": "SymmetricToeplitz", "circulant": "Circulant", "lower": "LowerTriangularToeplitz",
           "upper": "UpperTriangularToeplitz"}[kind]
    return getattr(O, cls)(v), getattr(tb, cls)(_dev(v))


KINDS = ["toeplitz", "symmetric", "circulant", "lower", "upper"]


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.float32], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("kind", KINDS)
def test_algebra_matches_dense(tb, kind, dtype):
    rng = np.random.default_rng(11)
    n = 9
    Ao, Ad = _mk(tb, kind, rng, n, dtype)
    Bo, Bd = _mk(tb, kind, rng, n, dtype)
    A, B = Ao.dense(), Bo.dense()
    D = lambda T: tb.dense(T).cpu().numpy()
    assert np.array_equal(D(Ad), A)
    assert np.array_equal(D(Ad + Bd), A + B) and np.array_equal(D(Ad - Bd), A - B)
    assert np.array_equal(D(-Ad), -A) and np.array_equal(D(2 * Ad), 2 * A) and np.array_equal(D(Ad * 3), A * 3)
    assert np.array_equal(D(tb.containers.conj(Ad)), A.conj())
    assert np.array_equal(D(tb.containers.real(Ad)), A.real) and np.array_equal(D(tb.containers.imag(Ad)), A.imag)
    assert np.array_equal(D(tb.containers.zero(Ad)), 0 * A)
    assert np.array_equal(D(tb.transpose(Ad)), A.T) and np.array_equal(D(tb.adjoint(Ad)), A.conj().T)
    # single-vector types keep their type (src/special.jl:24-33,77-79); mixed types fall back to Toeplitz
    if kind != "toeplitz":
        assert type(Ad + Bd) is type(Ad) and type(-Ad) is type(Ad) and type(2 * Ad) is type(Ad)
        To, Td = _mk(tb, "toeplitz", rng, n, dtype)
        assert type(Ad + Td) is tb.Toeplitz and np.array_equal(D(Ad + Td), A + To.dense())
    assert (Ad == tb.containers.copy(Ad)) and not (Ad == Bd)
    # tril / triu for every offset the reference sweeps (test/runtests.jl:374-388)
    for k in range(-n - 1, n + 2):
        L_, U_ = tb.tril(Ad, k), tb.triu(Ad, k)
        assert np.array_equal(D(L_), np.tril(A, k)), (kind, k)
        assert np.array_equal(D(U_), np.triu(A, k)), (kind, k)
        if kind in ("lower", "upper"):
            assert type(L_) is type(Ad) and type(U_) is type(Ad)
        else:
            assert type(L_) is tb.Toeplitz and type(U_) is tb.Toeplitz
    assert np.array_equal(D(Ad), A)                     # tril/triu did not touch the original
    # in-place scalar ops and copyto!
    C = tb.containers.copy(Ad)
    tb.lmul_(2, C)
    tb.rmul_(C, 0.5)
    assert np.array_equal(D(C), A)
    tb.copyto_(C, Bd)
    assert np.array_equal(D(C), B)
    assert tb.isdiag(Ad) == bool(np.array_equal(A, np.diag(np.diag(A))))
    for (i, j) in ((1, 1), (2, 5), (7, 3), (9, 9)):
        assert tb.getindex(Ad, i, j).item() == A[i - 1, j - 1]
    with pytest.raises(IndexError):
        tb.getindex(Ad, 0, 1)


def test_products_after_container_ops(tb):
    """The results feed straight back into the hot path: (2A - tril(A))*x, all on the device."""
    rng = np.random.default_rng(12)
    n = 3000
    Ao, Ad = _mk(tb, "toeplitz", rng, n, np.float64)
    x = rng.standard_normal(n)
    A = Ao.dense()
    want = (2 * A - np.tril(A, -2)) @ x
    Md = 2 * Ad - tb.tril(Ad, -2)
    got = tb.mul(Md, _dev(x)).cpu().numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-12


def test_rectangular_and_fill(tb):
    rng = np.random.default_rng(13)
    Ao, Ad = _mk(tb, "toeplitz", rng, 5, np.float64, m=8)
    A = Ao.dense()
    D = lambda T: tb.dense(T).cpu().numpy()
    for k in range(-9, 7):
        assert np.array_equal(D(tb.tril(Ad, k)), np.tril(A, k)) and np.array_equal(D(tb.triu(Ad, k)), np.triu(A, k))
    C = tb.containers.copy(Ad)
    tb.fill_(C, 7.0)
    assert np.array_equal(D(C), np.full((8, 5), 7.0))
    with pytest.raises(tb.DimensionMismatch):
        Ad + _mk(tb, "toeplitz", rng, 5, np.float64, m=7)[1]


def test_hankel_algebra_and_reverse(tb):
    """src/hankel.jl:75-130."""
    rng = np.random.default_rng(14)
    m, n = 6, 4
    v = rng.standard_normal(m + n - 1) + 1j * rng.standard_normal(m + n - 1)
    w = rng.standard_normal(m + n - 1) + 1j * rng.standard_normal(m + n - 1)
    Ho, Hd, Gd = O.Hankel(v, size=(m, n)), tb.Hankel(_dev(v), size=(m, n)), tb.Hankel(_dev(w), size=(m, n))
    H, G = Ho.dense(), O.Hankel(w, size=(m, n)).dense()
    D = lambda T: tb.dense(T).cpu().numpy()
    assert np.array_equal(D(Hd), H)
    assert np.array_equal(D(Hd + Gd), H + G) and np.array_equal(D(Hd - Gd), H - G) and np.array_equal(D(-Hd), -H)
    assert np.array_equal(D(2 * Hd), 2 * H) and np.array_equal(D(tb.containers.conj(Hd)), H.conj())
    assert np.array_equal(D(tb.transpose(Hd)), H.T) and np.array_equal(D(tb.adjoint(Hd)), H.conj().T)
    assert (Hd == tb.containers.copy(Hd)) and not (Hd == Gd)
    with pytest.raises(tb.DimensionMismatch):
        Hd + tb.Hankel(_dev(v), size=(n, m))
    # reverse: Hankel -> Toeplitz, both dims (src/hankel.jl:112-121)
    T1, T2 = tb.reverse(Hd, dims=1), tb.reverse(Hd, dims=2)
    assert type(T1) is tb.Toeplitz and type(T2) is tb.Toeplitz
    assert np.array_equal(D(T1), H[::-1, :]) and np.array_equal(D(T2), H[:, ::-1])
    with pytest.raises(tb.ArgumentError):
        tb.reverse(Hd, dims=3)
    # reverse: Toeplitz -> Hankel (src/hankel.jl:122-130); dims=1 is checked by value, dims=2 by the vector the
    # reference builds (its corner entry is duplicated; the reference only type-checks it, test/runtests.jl:229)
    vc, vr = rng.standard_normal(m), rng.standard_normal(n)
    vr[0] = vc[0]
    Td = tb.Toeplitz(_dev(vc), _dev(vr))
    R1, R2 = tb.reverse(Td), tb.reverse(Td, dims=2)
    assert type(R1) is tb.Hankel and type(R2) is tb.Hankel
    assert np.array_equal(D(R1), O.Toeplitz(vc, vr).dense()[::-1, :])
    assert np.array_equal(R2.v.cpu().numpy(), np.concatenate([vr[::-1], vc])) and R2.shape == (m, n)
    # the hot path accepts the result: reverse(H, dims=2) * reverse(x) == H * x (src/hankel.jl:133-134)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y1 = tb.mul(T2, _dev(x[::-1].copy())).cpu().numpy()
    assert np.linalg.norm(y1 - H @ x) / np.linalg.norm(H @ x) <= 1e-12
