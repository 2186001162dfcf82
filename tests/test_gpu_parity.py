"""GPU parity tests: the CUDA path, called through the host mirror -> C ABI, against the
CPU oracle on the same seeded inputs.  Tolerances are the north_star's: relative 2-norm
error <= 1e-12 (Float64/ComplexF64) and <= 1e-5 (Float32/ComplexF32)."""
import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O
from oracle import build_oracle as OC
import parity_tools as PT

pytestmark = pytest.mark.gpu

TOL64, TOL32 = 1e-12, 1e-5


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t


def devmat(a):
    """numpy (n, l) -> column-major CUDA tensor (n, l)."""
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def host(t):
    return t.detach().cpu().numpy()


def p(base, n, dtype=np.float64):
    return (base ** np.arange(n)).astype(dtype)


def tol_of(dtype):
    return TOL32 if np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64)) else TOL64


def make_pair(tb, kind, n, dtype, m=None):
    """(oracle matrix, device matrix) for the reference's coefficient recipes (test/runtests.jl:23-48)."""
    m = n if m is None else m
    if kind == "toeplitz":
        vc, vr = p(0.9, m, dtype), p(0.4, n, dtype)
        return O.Toeplitz(vc, vr), tb.Toeplitz(dev(vc), dev(vr))
    v = p(0.9, n, dtype)
    cls = {"symmetric": "SymmetricToeplitz", "circulant": "Circulant", "lower": "LowerTriangularToeplitz",
           "upper": "UpperTriangularToeplitz"}[kind]
    return getattr(O, cls)(v), getattr(tb, cls)(dev(v))


KINDS = ["toeplitz", "symmetric", "circulant", "lower", "upper"]
DTYPES = [np.float64, np.complex128, np.float32, np.complex64]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [101, 2000])
def test_mul_matrix_and_adjoint(tb, kind, dtype, n):
    """test/runtests.jl:50-55: A*X and A'*X, X an n x 5 matrix (FFT path even at n=101)."""
    rng = np.random.default_rng(n)
    Ao, Ad = make_pair(tb, kind, n, dtype)
    rdt = np.zeros(1, dtype).real.dtype
    X = rng.standard_normal((n, 5)).astype(rdt)
    Y = host(tb.mul(Ad, devmat(X)))
    assert rel(Y, O.matvec(Ao, X)) <= tol_of(dtype)
    Ya = host(tb.mul(tb.adjoint(Ad), devmat(X)))
    assert rel(Ya, O.matvec(Ao.adjoint(), X)) <= tol_of(dtype)
    # complex right-hand sides, vector form (n=101 takes the direct N<512 path)
    xc = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.result_type(dtype, np.complex64))
    yc = host(tb.mul(Ad, dev(xc)))
    assert rel(yc, O.matvec(Ao, xc)) <= tol_of(dtype)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_rectangular(tb, dtype):
    """test/runtests.jl:83-95."""
    rng = np.random.default_rng(1)
    rdt = np.zeros(1, dtype).real.dtype
    for m, n in ((2000, 101), (101, 2000)):
        Ao, Ad = make_pair(tb, "toeplitz", n, dtype, m=m)
        X = rng.standard_normal((n, 5)).astype(rdt)
        assert rel(host(tb.mul(Ad, devmat(X))), O.matvec(Ao, X)) <= tol_of(dtype)
        x = X[:, 0].copy()
        assert rel(host(tb.mul(Ad, dev(x))), O.matvec(Ao, x)) <= tol_of(dtype)


def test_mixed_types_exact(tb):
    """test/runtests.jl:74-81 (exact small cases)."""
    A = tb.Toeplitz(torch.tensor([1, 2]), torch.tensor([1, 2]))
    assert A.dtype == torch.float64
    y = tb.mul(A, torch.ones(2, dtype=torch.float64, device="cuda"))
    assert np.array_equal(host(y), np.full(2, 3.0))
    C = tb.Circulant(torch.arange(1, 4, dtype=torch.float32))
    y = tb.mul(C, torch.ones(3, dtype=torch.float64, device="cuda"))
    assert y.dtype == torch.float64 and np.array_equal(host(y), np.full(3, 6.0))


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.complex64], ids=lambda d: np.dtype(d).name)
def test_hankel(tb, dtype):
    """test/runtests.jl:154-215."""
    rng = np.random.default_rng(2)
    H = tb.Hankel(torch.tensor([1.0, 2, 3, 4, 5]), vr=torch.tensor([5.0, 6, 7, 8, 9]))
    x = torch.ones(5, dtype=torch.float64, device="cuda")
    assert rel(host(tb.mul(H, x)), np.array([15.0, 20, 25, 30, 35])) <= 1e-14
    rdt = np.zeros(1, dtype).real.dtype
    for m, n in ((101, 101), (2000, 2000), (101, 2000), (2000, 101)):
        vc = (0.9 ** np.arange(m - 1, -1, -1)).astype(dtype)
        vr = (0.4 ** np.arange(n)).astype(dtype)
        Ho = O.Hankel(vc, vr=vr)
        Hd = tb.Hankel(dev(vc), vr=dev(vr))
        X = rng.standard_normal((n, 5)).astype(rdt)
        assert rel(host(tb.mul(Hd, devmat(X))), O.matvec(Ho, X)) <= tol_of(dtype)
        assert rel(host(tb.mul(Hd, dev(X[:, 0].copy()))), O.matvec(Ho, X[:, 0].copy())) <= tol_of(dtype)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128], ids=lambda d: np.dtype(d).name)
def test_alpha_beta_and_nan_safety(tb, dtype):
    """mul!(y, A, x, alpha, beta): beta == 0 must not read y (src/linearalgebra.jl:68-71)."""
    rng = np.random.default_rng(3)
    n = 700
    Ao, Ad = make_pair(tb, "toeplitz", n, dtype)
    X = rng.standard_normal((n, 3)).astype(dtype)
    Y0 = rng.standard_normal((n, 3)).astype(dtype)
    a, b = (1.5, -0.25) if np.dtype(dtype).kind == "f" else (1.5 - 0.5j, -0.25 + 2j)
    Yd = devmat(Y0.copy())
    tb.mul_(Yd, Ad, devmat(X), a, b)
    Yo = np.asfortranarray(Y0.copy())
    O.mul_mat(Yo, Ao, X, dtype(a) if np.dtype(dtype).kind == "f" else a, dtype(b) if np.dtype(dtype).kind == "f" else b)
    assert rel(host(Yd), Yo) <= TOL64
    Yn = devmat(np.full((n, 3), np.nan, dtype=dtype))
    tb.mul_(Yn, Ad, devmat(X), a, False)
    assert np.isfinite(host(Yn)).all()
    assert rel(host(Yn), a * O.matvec(Ao, X)) <= TOL64
    # a factorization can be reused (README.md:25-28)
    F = tb.factorize(Ad)
    assert rel(host(tb.mul(F, devmat(X))), O.matvec(Ao, X)) <= TOL64


@pytest.mark.parametrize("n,cols,dtype", [(5000, 4, np.float64), (9000, 3, np.complex128), (40000, 5, np.float64),
                                          (70000, 2, np.float32), (300000, 3, np.complex64)],
                         ids=lambda v: str(v))
def test_two_level_sizes(tb, n, cols, dtype):
    """Embeddings too long for one shared-memory FFT take the three-kernel two-level path."""
    rng = np.random.default_rng(n)
    rdt = np.zeros(1, dtype).real.dtype
    vc = rng.standard_normal(n).astype(dtype)
    vr = rng.standard_normal(n).astype(dtype)
    vr[0] = vc[0]
    Ao, Ad = O.Toeplitz(vc, vr), tb.Toeplitz(dev(vc), dev(vr))
    X = rng.standard_normal((n, cols)).astype(rdt if np.dtype(dtype).kind == "f" else dtype)
    assert rel(host(tb.mul(Ad, devmat(X))), O.matvec(Ao, X)) <= tol_of(dtype)
    Hv = rng.standard_normal(2 * n - 1).astype(dtype)
    Ho, Hd = O.Hankel(Hv, (n, n)), tb.Hankel(dev(Hv), (n, n))
    assert rel(host(tb.mul(Hd, dev(X[:, 0].copy()))), O.matvec(Ho, X[:, 0].copy())) <= tol_of(dtype)


def test_c1_config_single_vector(tb):
    """BASELINE config 1: Toeplitz{Float64} n = 2^16, one vector."""
    n = 1 << 16
    rng = np.random.default_rng(0)
    for vc, vr in ((p(0.9, n), p(0.4, n)), (lambda a, b: (a, np.concatenate([a[:1], b[1:]])))(
            np.random.default_rng(1).standard_normal(n), np.random.default_rng(2).standard_normal(n))):
        x = rng.standard_normal(n)
        y = host(tb.mul(tb.Toeplitz(dev(vc), dev(vr)), dev(x)))
        assert rel(y, O.matvec(O.Toeplitz(vc, vr), x)) <= TOL64


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.complex64], ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("n", [5, 6, 64, 101, 2000, 2048, 32768])
def test_circulant_family(tb, dtype, n):
    """src/linearalgebra.jl:183-315: power-of-two sizes (direct plan, incl. a two-level size) and the
    reference's own test sizes 5, 6, 101, 2000 (chirp-z plan)."""
    rng = np.random.default_rng(n + 7)
    tag = "circulant n=%d %s " % (n, np.dtype(dtype).name)

    def chk(name, got, ref, cond=1.0):
        PT.check(tag + name, got, ref, dtype, cond)
    v = p(0.9, n, dtype)
    v2 = (rng.random(n) + 0.5).astype(dtype)
    Co, Cd = O.Circulant(v), tb.Circulant(dev(v))
    C2o, C2d = O.Circulant(v2), tb.Circulant(dev(v2))
    Fo, Fd = O.factorize(Co), tb.factorize(Cd)
    assert Fd.size() == (n, n) and Fd.size(1) == n
    # eigvals: DFT order, fft(C.vc) == spectrum (test/runtests.jl:543)
    lam = np.fft.fft(v.astype(np.complex128))
    lam2 = np.fft.fft(v2.astype(np.complex128))
    kC, kC2 = PT.cond_spectrum(lam), PT.cond_spectrum(lam2)      # 2-norm condition numbers of C and C2 (normal matrices)
    chk("eigvals", host(tb.eigvals(Cd)), lam)
    # ldiv! vector / matrix / 3-arg
    B = (rng.standard_normal((n, 4)) + (1j * rng.standard_normal((n, 4)) if np.dtype(dtype).kind == "c" else 0)).astype(dtype)
    Bo = np.asfortranarray(B.copy())
    O.ldiv_mat(Co, Bo, O.circ_ldiv)
    Bd = devmat(B.copy())
    tb.ldiv_(Cd, Bd)
    chk("ldiv! matrix", host(Bd), Bo, kC)
    xo = torch.zeros(n, dtype=Bd.dtype, device="cuda")
    tb.circ_ldiv_(Fd, dev(B[:, 0].copy()), out=xo)
    chk("ldiv! 3-arg", host(xo), Bo[:, 0], kC)
    chk("ldiv vector", host(tb.ldiv(Cd, dev(B[:, 1].copy()))), Bo[:, 1], kC)
    # inv / pinv / sqrt / det / products
    chk("inv", host(tb.inv(Cd).vc), O.circ_inv(Co).vc, kC)
    chk("pinv", host(tb.pinv(Cd).vc), O.circ_pinv(Co).vc, kC)
    chk("sqrt", host(tb.sqrt(C2d).vc), O.circ_sqrt(C2o).vc, np.sqrt(kC2))
    for ta in (False, True):
        for tbb in (False, True):
            Ao_ = Fo.adjoint() if ta else Fo
            Bo_ = O.factorize(C2o).adjoint() if tbb else O.factorize(C2o)
            Ad_ = tb.adjoint(Fd) if ta else Fd
            Bd_ = tb.adjoint(tb.factorize(C2d)) if tbb else C2d
            chk("A*B adj=%d%d" % (ta, tbb), host(tb.circ_mul(Ad_, Bd_).vc), O.circ_mul(Ao_, Bo_).vc)
    # issue #47: inv(F) * C ~ I
    e = rng.random(n).astype(np.zeros(1, dtype).real.dtype)
    Iv = tb.circ_mul(tb.inv(Fd), Cd)
    chk("inv(F)*C ~ I", host(tb.mul(Iv, dev(e))), e, kC)
    if n <= 64:
        # a determinant is a product of n eigenvalues: relative error <= n * (error of one eigenvalue)
        chk("det", host(tb.det(C2d)).item(), O.circ_det(C2o), float(n))


def test_circulant_identity_issue73(tb):
    """test/runtests.jl:794-803."""
    n = 6
    b = np.random.default_rng(6).random(n)
    P = tb.Circulant(dev(np.eye(n)[0]))
    F = tb.factorize(P)
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    tb.circ_ldiv_(F, dev(b), out=x)
    assert rel(host(x), b) <= 1e-14


def test_pinv_singular(tb):
    """pinv of the all-ones circulant (test/runtests.jl:558-565)."""
    for n in (5, 64):
        for dtype in (np.float64, np.complex128):
            v = np.ones(n, dtype)
            assert rel(host(tb.pinv(tb.Circulant(dev(v)), 1e-10).vc), O.circ_pinv(O.Circulant(v), 1e-10).vc) <= 1e-12


@pytest.mark.parametrize("n", [1, 2, 8, 16, 17, 32, 33, 48, 49, 64, 65, 96, 97, 128, 129, 192, 193, 200, 256, 257, 384, 512, 513,
                               700, 768, 769, 1024, 1025, 1536, 1537, 2048, 2100])
def test_levinson_durbin_batched(tb, n):
    """src/directLinearSolvers.jl:15-28,100-134 batched over columns, vs the C oracle."""
    rng = np.random.default_rng(n)
    nsys = 37
    R = np.asfortranarray(rng.random((n, nsys)) / max(n, 2))          # diag-dominant recipe (test/runtests.jl:117)
    Bm = np.asfortranarray(rng.standard_normal((n, nsys)))
    if n >= 2:
        Y = host(tb.durbin(devmat(R)))
        assert rel(Y, OC.c_durbin_batched(R)) <= TOL64
    Rl = np.asfortranarray(R[: n - 1, :]) if n > 1 else np.zeros((0, nsys), order="F")
    X = tb.colmajor(n, nsys)
    tb.levinson_(X, devmat(Rl) if n > 1 else tb.colmajor(0, nsys), devmat(Bm))
    ref = OC.c_levinson_batched(Rl, Bm) if n > 1 else Bm.copy()
    assert rel(host(X), ref) <= TOL64


def test_levinson_reference_kernels(tb):
    """The n = 8 kernels of test/runtests.jl:112-151, incl. the scaled-diagonal wrapper."""
    n = 8
    xg = np.linspace(-1, 1, n + 1)
    g = np.random.default_rng(3)
    for a in (g.random(n + 1) / n, np.exp(-np.abs(xg[0] - xg)), np.exp(-np.abs(xg[0] - xg) ** 2) / 2):
        a = a.copy()
        a[0] = 1
        r = a[1:]
        T = O.SymmetricToeplitz(np.concatenate([[1.0], r[:-1]])).dense()
        kT = PT.cond2(T)                                  # exp(-|x|) and exp(-x^2)/2 kernels: cond up to ~1e3
        PT.check("durbin n=8 kernel", host(tb.durbin(dev(r))), O.durbin(r), np.float64, kT)
        b = g.standard_normal(n)
        PT.check("levinson n=8 kernel", host(tb.levinson(dev(r[:-1].copy()), dev(b))), O.levinson(r[:-1], b), np.float64, kT)
        TS = 2.1 * np.concatenate([[1.0], r[:-1]])
        xs = host(tb.levinson(tb.SymmetricToeplitz(dev(TS)), dev(b)))
        PT.check("levinson(T, b) vs dense solve", xs, np.linalg.solve(O.SymmetricToeplitz(TS).dense(), b), np.float64, kT)
        Bm = g.standard_normal((n, 3))
        Xs = host(tb.levinson(tb.SymmetricToeplitz(dev(TS)), devmat(Bm)))
        PT.check("levinson(T, B) vs dense solve", Xs, np.linalg.solve(O.SymmetricToeplitz(TS).dense(), Bm), np.float64, kT)


def test_levinson_float32(tb):
    rng = np.random.default_rng(11)
    n, nsys = 256, 9
    R = np.asfortranarray((rng.random((n - 1, nsys)) / n).astype(np.float32))
    Bm = np.asfortranarray(rng.standard_normal((n, nsys)).astype(np.float32))
    X = tb.colmajor(n, nsys, torch.float32)
    tb.levinson_(X, devmat(R), devmat(Bm))
    assert rel(host(X), OC.c_levinson_batched(R.astype(np.float64), Bm.astype(np.float64))) <= TOL32


@pytest.mark.parametrize("dtype", [np.float64, np.complex128, np.complex64], ids=lambda d: np.dtype(d).name)
def test_cgs_matches_oracle(tb, dtype):
    """cgs + Strang on a KMS SymmetricToeplitz: same iterates (count, flag, residual) as the oracle."""
    n = 3000
    rng = np.random.default_rng(7)
    v = p(0.5, n, dtype)
    b = rng.standard_normal(n).astype(dtype)
    Ao, Ad = O.SymmetricToeplitz(v), tb.SymmetricToeplitz(dev(v))
    tol = 100 * np.finfo(np.float64).eps if np.dtype(dtype) != np.complex64 else 1e-5
    xo, eo, ito, fo = O.cgs(Ao, np.zeros(n, dtype), b.copy(), O.factorize(O.strang(Ao)), 50, tol)
    xd, ed, itd, fd = tb.cgs(Ad, torch.zeros(n, dtype=dev(b).dtype, device="cuda"), dev(b),
                             tb.factorize(tb.strang(Ad)), 50, tol)
    assert fd == fo and abs(itd - ito) <= 1
    kA = PT.cond2(Ao.dense())                              # KMS(0.5): ~9
    # both runs stop at a relative residual <= tol: the solutions agree to cond * (res_o + res_d)
    PT.check("cgs to tolerance", host(xd), xo, dtype, kA, extra=2.0 * kA * (eo + ed))
    assert ed <= max(10 * eo, tol)
    # fixed iteration count (no convergence -> flag 1): the iteration keeps running after the residual has reached
    # its rounding floor, so the iterates agree to cond * (floor of the residual), measured on both sides
    xo2, eo2, ito2, fo2 = O.cgs(Ao, np.zeros(n, dtype), b.copy(), O.factorize(O.strang(Ao)), 3, 0.0)
    xd2, ed2, itd2, fd2 = tb.cgs(Ad, torch.zeros(n, dtype=dev(b).dtype, device="cuda"), dev(b),
                                 tb.factorize(tb.strang(Ad)), 3, 0.0)
    assert (itd2, fd2) == (ito2, fo2) == (3, 1)
    D = Ao.dense().astype(np.complex128)
    true_o = np.linalg.norm(b - D @ xo2) / np.linalg.norm(b)
    true_d = np.linalg.norm(b - D @ host(xd2)) / np.linalg.norm(b)
    PT.check("cgs 3 iterations", host(xd2), xo2, dtype, kA, extra=2.0 * kA * (true_o + true_d))


def test_strang_chan(tb):
    n = 1000
    for kind in ("toeplitz", "symmetric", "lower", "upper"):
        for dtype in (np.float64, np.complex128):
            Ao, Ad = make_pair(tb, kind, n, dtype)
            assert np.array_equal(host(tb.strang(Ad).vc), O.strang(Ao).vc)
            assert rel(host(tb.chan(Ad).vc), O.chan(Ao).vc) <= 1e-15


@pytest.mark.parametrize("n", [101, 1024, 2000])
@pytest.mark.parametrize("kind", ["toeplitz", "symmetric", "lower", "upper", "circulant"])
def test_ldiv_drivers(tb, kind, n):
    """ldiv!(A, b): cgs+Strang / cg+Strang / cgs+Chan / spectral (test/runtests.jl:58-59,103-105)."""
    rng = np.random.default_rng(9)
    Ao, Ad = make_pair(tb, kind, n, np.float64)
    B = rng.standard_normal((n, 2))
    ref = np.linalg.solve(Ao.dense(), B)
    Bd = devmat(B.copy())
    tb.ldiv_(Ad, Bd)
    assert rel(host(Bd), ref) <= np.sqrt(np.finfo(float).eps)
    with pytest.raises(tb.ArgumentError):
        tb.ldiv(Ad, dev(B[:, 0].astype(np.complex128)))


def test_cg_early_return_quirk(tb):
    """cg returns `nothing` when the start residual is already below tol (src/iterativeLinearSolvers.jl:57-59)."""
    n = 600
    A = tb.SymmetricToeplitz(dev(p(0.5, n)))
    b = torch.zeros(n, dtype=torch.float64, device="cuda")
    assert tb.cg(A, torch.zeros_like(b), b, tb.strang(A), 10, 1e-10) is None
    with pytest.raises(tb.ToeplitzB200Error):
        tb.ldiv_(A, b)


def test_errors(tb):
    n = 600
    A = tb.Toeplitz(dev(p(0.9, n)), dev(p(0.4, n)))
    z = lambda k: torch.zeros(k, dtype=torch.float64, device="cuda")
    with pytest.raises(tb.DimensionMismatch):
        tb.mul_(z(n - 1), A, z(n))
    with pytest.raises(tb.DimensionMismatch):
        tb.mul_(z(n), A, z(n + 1))
    with pytest.raises(tb.DimensionMismatch):
        tb.mul_(tb.colmajor(n, 2), A, tb.colmajor(n, 3))
    with pytest.raises(tb.DimensionMismatch):
        tb.circ_ldiv_(tb.factorize(tb.Circulant(z(64) + 1)), z(65))
    with pytest.raises(tb.ToeplitzB200Error):
        tb.ldiv_(tb.Toeplitz(dev(p(0.9, 4)), dev(p(0.4, 3))), tb.colmajor(4, 1))
    with pytest.raises(tb.ToeplitzB200Error):
        tb.Toeplitz(dev(np.array([1.0, 2])), dev(np.array([2.0, 1])))
    with pytest.raises(tb.DimensionMismatch):
        tb.circ_mul(tb.Circulant(z(64) + 1), tb.Circulant(z(128) + 1))
    with pytest.raises(tb.ArgumentError):
        tb.mul_(z(n), A, z(n).to(torch.complex128))        # complex product into a real y


def test_mul_host_pipeline(tb):
    """tb200_mul_host: host buffers in, host buffers out, chunked + overlapped."""
    n, l = 20000, 13
    rng = np.random.default_rng(5)
    vc, vr = p(0.9, n), p(0.4, n)
    X = rng.standard_normal((n, l))
    Xh = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory().t()
    Yh = torch.empty((l, n), dtype=torch.float64).pin_memory().t()
    F = tb.factorize_host(tb.Toeplitz, torch.from_numpy(vc), torch.from_numpy(vr))
    tb.set_option("host_chunk_bytes", 1 << 20)        # force several chunks
    tb.mul_host_(Yh, F, Xh)
    tb.set_option("host_chunk_bytes", 256 << 20)
    assert rel(Yh.numpy(), O.matvec(O.Toeplitz(vc, vr), X)) <= TOL64


def test_bluestein_factorization_products(tb):
    """mul! through a Circulant *factorization* of non-power-of-two size (chirp-z apply), real and complex,
    vector and matrix, alpha/beta."""
    rng = np.random.default_rng(21)
    for n in (6, 101, 2000):
        for dtype in (np.float64, np.complex128, np.complex64):
            v = (rng.random(n) + 0.5).astype(dtype)
            Co, Cd = O.Circulant(v), tb.Circulant(dev(v))
            F = tb.factorize(Cd)
            X = rng.standard_normal((n, 3)).astype(dtype)
            Y0 = rng.standard_normal((n, 3)).astype(dtype)
            Yd = devmat(Y0.copy())
            tb.mul_(Yd, F, devmat(X), 2.0, -0.5)
            wide = np.complex128 if np.dtype(dtype).kind == "c" else np.float64
            Dw, Xw = Co.dense().astype(wide), X.astype(wide)             # the definition, evaluated in double
            PT.check("bluestein n=%d mul! alpha/beta" % n, host(Yd), 2.0 * (Dw @ Xw) - 0.5 * Y0.astype(wide), dtype)
            PT.check("bluestein n=%d mul vector" % n, host(tb.mul(F, dev(X[:, 0].copy()))), Dw @ Xw[:, 0], dtype)


def test_edge_shapes(tb):
    """Empty batches, 1x1 and tiny matrices, odd column counts with beta, strided (ld > n) matrices."""
    rng = np.random.default_rng(22)
    n = 700
    Ao, Ad = make_pair(tb, "toeplitz", n, np.float64)
    F = tb.factorize(Ad)
    # zero columns: nothing to do, no error
    Y0 = tb.colmajor(n, 0)
    assert tb.mul_(Y0, F, tb.colmajor(n, 0)).shape == (n, 0)
    # odd column count, beta != 0 (pair packing with a dangling column)
    X = rng.standard_normal((n, 5))
    Yi = rng.standard_normal((n, 5))
    Yd = devmat(Yi.copy())
    tb.mul_(Yd, F, devmat(X), 1.0, 1.0)
    assert rel(host(Yd), Ao.dense() @ X + Yi) <= TOL64
    # leading dimension larger than the row count (views into a taller matrix)
    big_x = torch.zeros((5, n + 37), dtype=torch.float64, device="cuda")
    big_y = torch.full((5, n + 11), np.nan, dtype=torch.float64, device="cuda")
    big_x[:, :n] = torch.from_numpy(np.ascontiguousarray(X.T)).cuda()
    tb.mul_(big_y.t()[:n, :], F, big_x.t()[:n, :])
    assert rel(host(big_y[:, :n].t()), Ao.dense() @ X) <= TOL64
    assert torch.isnan(big_y[:, n:]).all()
    # 1x1 and 2x3 matrices (vector path: direct; matrix path: FFT with the minimum embedding)
    T1 = tb.Toeplitz(dev(np.array([2.5])), dev(np.array([2.5])))
    assert host(tb.mul(T1, dev(np.array([4.0]))))[0] == 10.0
    T23o = O.Toeplitz(np.array([1.0, 2.0]), np.array([1.0, 3.0, 4.0]))
    T23 = tb.Toeplitz(dev(T23o.vc), dev(T23o.vr))
    X3 = rng.standard_normal((3, 4))
    assert rel(host(tb.mul(T23, devmat(X3))), T23o.dense() @ X3) <= TOL64
    # Hankel with surplus anti-diagonals: the reference's mul! fails its size check (src/hankel.jl:117)
    with pytest.raises(tb.DimensionMismatch):
        tb.mul(tb.Hankel(dev(np.arange(12.0)), (5, 5)), dev(np.ones(5)))
    # levinson on zero systems
    assert tb.levinson_(tb.colmajor(8, 0), tb.colmajor(7, 0), tb.colmajor(8, 0)).shape == (8, 0)


@pytest.mark.parametrize("n", [1, 2, 8, 64, 65, 128, 700, 3000])
def test_triangular_inverse(tb, n):
    """inv(::Lower/UpperTriangularToeplitz) by recursive doubling (src/linearalgebra.jl:337-396), SURVEY 8f-2."""
    for dtype in (np.float64, np.complex128):
        v = np.concatenate([[1.0], 0.5 * 0.7 ** np.arange(1, n)]).astype(dtype)
        if np.dtype(dtype).kind == "c":
            v = v * np.exp(0.3j * np.arange(n))
        ref = O.tri_inv_lower(v)
        kL = PT.cond2(O.LowerTriangularToeplitz(v).dense()) if n <= 700 else 9.0      # |1 + 0.5 z/(1-0.7z)| on |z|=1: cond < 9
        PT.check("tri inv lower n=%d" % n, host(tb.inv(tb.LowerTriangularToeplitz(dev(v))).v), ref, dtype, kL)
        Ui = tb.inv(tb.UpperTriangularToeplitz(dev(v)))
        assert isinstance(Ui, tb.UpperTriangularToeplitz)
        if n <= 700:
            dense = np.linalg.inv(O.UpperTriangularToeplitz(v).dense())
            PT.check("tri inv upper n=%d" % n, O.UpperTriangularToeplitz(host(Ui.v)).dense(), dense, dtype, kL)


@pytest.mark.parametrize("n", [1, 2, 9, 100, 513, 1500])
def test_trench(tb, n):
    """trench(T) (src/directLinearSolvers.jl:36-85; test/runtests.jl:129-135), SURVEY 8f-3."""
    rng = np.random.default_rng(n)
    a = np.concatenate([[1.0], rng.random(n - 1) / max(n, 2)])
    T = O.SymmetricToeplitz(a)
    B = host(tb.trench(tb.SymmetricToeplitz(dev(a))))
    kT = PT.cond2(T.dense())
    PT.check("trench n=%d vs dense inv" % n, B, np.linalg.inv(T.dense()), np.float64, kT)
    if 2 <= n <= 100:
        PT.check("trench n=%d vs oracle" % n, B, O.trench(T), np.float64, kT)
    TS = O.SymmetricToeplitz(2.3 * a)
    PT.check("trench scaled n=%d" % n, host(tb.trench(tb.SymmetricToeplitz(dev(2.3 * a)))), np.linalg.inv(TS.dense()), np.float64, kT)
    if n >= 2:
        PT.check("trench(r) n=%d" % n, host(tb.trench(dev(a[1:].copy()))), np.linalg.inv(T.dense()), np.float64, kT)


def test_mul_host_pageable_and_complex(tb):
    """tb200_mul_host with pageable (unpinned) host memory, complex data, alpha/beta and a Hankel handle."""
    n, l = 9000, 7
    rng = np.random.default_rng(31)
    v = (rng.standard_normal(2 * n - 1) + 1j * rng.standard_normal(2 * n - 1)).astype(np.complex128)
    X = (rng.standard_normal((n, l)) + 1j * rng.standard_normal((n, l))).astype(np.complex128)
    Y0 = (rng.standard_normal((n, l)) + 1j * rng.standard_normal((n, l))).astype(np.complex128)
    Xh = torch.from_numpy(np.ascontiguousarray(X.T)).t()            # pageable, column-major
    Yh = torch.from_numpy(np.ascontiguousarray(Y0.T)).t()
    F = tb.factorize_host(tb.Hankel, torch.from_numpy(v), size=(n, n))
    tb.mul_host_(Yh, F, Xh, 0.5 - 1j, 2.0)
    ref = (0.5 - 1j) * O.matvec(O.Hankel(v, (n, n)), X) + 2.0 * Y0
    assert rel(Yh.numpy(), ref) <= TOL64


@pytest.mark.parametrize("n", [2, 5, 128, 129, 300, 512])
def test_levinson_multirhs_shared_recursion(tb, n):
    """levinson(T::SymmetricToeplitz, B::Matrix): one matrix, many right-hand sides (the column loop of
    ext/ToeplitzMatricesStatsBaseExt.jl:15-24) with the y-recursion tabulated once; vs the C oracle per column."""
    rng = np.random.default_rng(n)
    a = 2.5 * np.concatenate([[1.0], rng.random(n - 1) / n])          # non-unit diagonal
    Bm = np.asfortranarray(rng.standard_normal((n, 33)))
    X = host(tb.levinson(tb.SymmetricToeplitz(dev(a)), devmat(Bm)))
    ref = np.stack([OC.c_levinson_sym(a, Bm[:, j]) for j in range(Bm.shape[1])], axis=1)
    assert rel(X, ref) <= TOL64
    D = O.SymmetricToeplitz(a).dense()
    PT.check("levinson multirhs n=%d vs dense solve" % n, X, np.linalg.solve(D, Bm), np.float64, PT.cond2(D))


@pytest.mark.parametrize("n", [1000, 70000])
def test_spectrum_transplant_single_gpu(tb, n):
    """The multi-GPU plumbing on one device: an empty handle that receives another handle's spectrum bytes
    (what the NCCL broadcast does, sharding.factorize_shared) multiplies exactly like the original."""
    rng = np.random.default_rng(n)
    vc, vr = rng.standard_normal(n), rng.standard_normal(n)
    vr[0] = vc[0]
    F0 = tb.factorize(tb.Toeplitz(dev(vc), dev(vr)))
    F1 = tb.factorize_empty_like(0, torch.float64, (n, n))
    s0, s1 = tb.spectrum_tensor(F0), tb.spectrum_tensor(F1)
    assert s0.dtype == torch.uint8 and s0.shape == s1.shape and s0.numel() > 0
    s1.copy_(s0)
    X = devmat(rng.standard_normal((n, 3)))
    Y0, Y1 = tb.colmajor(n, 3), tb.colmajor(n, 3)
    tb.mul_(Y0, F0, X)
    tb.mul_(Y1, F1, X)
    assert torch.equal(Y0, Y1)
    assert rel(host(Y1), O.matvec(O.Toeplitz(vc, vr), host(X))) <= TOL64
