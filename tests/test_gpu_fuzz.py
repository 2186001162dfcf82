"""Seeded random-shape parity sweep of `mul!` through the C ABI against the CPU oracle.

The reference's tests use a handful of sizes (test/runtests.jl:23-95, 154-215); the CUDA path has more
regimes than the FFTW path has (direct product for N < 512, one shared-memory transform, two-level
transform, chirp-z for non-power-of-two circulants, two real columns packed per transform, odd column
tails, padded leading dimensions, alpha/beta epilogue), so this sweep draws sizes log-uniformly across
all of them.  Tolerances: relative 2-norm error 1e-12 (Float64 family) / 1e-5 (Float32 family)."""
import os

import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O
import parity_tools as PT

pytestmark = pytest.mark.gpu

# soak runs: TB200_FUZZ_SCALE=8 multiplies the number of seeds, TB200_FUZZ_SEED0 shifts them (defaults: the committed sweep)
_SCALE = max(1, int(os.environ.get("TB200_FUZZ_SCALE", "1")))
_SEED0 = int(os.environ.get("TB200_FUZZ_SEED0", "0"))


def _seeds(k):
    return range(_SEED0, _SEED0 + k * _SCALE)


KINDS = ["toeplitz", "symmetric", "circulant", "lower", "upper", "hankel"]
DTYPES = [np.float64, np.complex128, np.float32, np.complex64]


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def _rand(rng, shape, dtype):
    dt = np.dtype(dtype)
    a = rng.standard_normal(shape)
    if dt.kind == "c":
        a = a + 1j * rng.standard_normal(shape)
    return a.astype(dt)


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _devmat_padded(a, pad):
    """numpy (n, l) -> column-major CUDA view (n, l) whose leading dimension is n + pad."""
    n, l = a.shape
    big = torch.full((l, n + pad), float("nan"), dtype=torch.from_numpy(a[:0]).dtype, device="cuda")
    big[:, :n] = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    return big.t()[:n, :]


def _rel(a, b):
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)


def _amp(Ao, m, n, X, Y):
    """An output of few rows can come out small by chance (m = 1 is a single dot product of n terms): the rounding
    error of the FFT product -- reference and CUDA path alike -- is relative to the circular convolution the rows are
    cut from, about ||coefficients|| * ||x|| * sqrt(m / N) for these random inputs, not to ||y|| itself.  Returns the
    ratio (>= 1) that widens the relative tolerance in those cases; ~1 for ordinary shapes."""
    if hasattr(Ao, "vc"):
        cn = np.sqrt(np.linalg.norm(Ao.vc) ** 2 + np.linalg.norm(Ao.vr) ** 2)
    else:
        cn = np.linalg.norm(Ao.v)
    expected = float(cn) * float(np.linalg.norm(X)) * np.sqrt(m / float(m + n - 1))
    return max(1.0, expected / max(float(np.linalg.norm(Y)), 1e-300))


def _size(rng, hi):
    return int(np.exp(rng.uniform(0.0, np.log(hi))))


def _build(tb, rng, kind, dtype, hi):
    n = max(1, _size(rng, hi))
    if kind == "toeplitz":
        m = max(1, _size(rng, hi)) if rng.random() < 0.5 else n
        vc, vr = _rand(rng, m, dtype), _rand(rng, n, dtype)
        vr[0] = vc[0]
        return O.Toeplitz(vc, vr), tb.Toeplitz(_dev(vc), _dev(vr)), m, n
    if kind == "hankel":
        m = max(1, _size(rng, hi)) if rng.random() < 0.5 else n
        v = _rand(rng, m + n - 1, dtype)
        return O.Hankel(v, size=(m, n)), tb.Hankel(_dev(v), size=(m, n)), m, n
    v = _rand(rng, n, dtype)
    cls = {"symmetric": "SymmetricToeplitz", "circulant": "Circulant", "lower": "LowerTriangularToeplitz",
           "upper": "UpperTriangularToeplitz"}[kind]
    return getattr(O, cls)(v), getattr(tb, cls)(_dev(v)), n, n


@pytest.mark.parametrize("seed", _seeds(12))
def test_mul_random_shapes(tb, seed):
    rng = np.random.default_rng(1000 + seed)
    for case in range(14):
        kind = KINDS[int(rng.integers(len(KINDS)))]
        dtype = DTYPES[int(rng.integers(len(DTYPES)))]
        single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
        hi = 70000 if case % 7 == 0 else 6000          # a few cases reach the two-level transform
        Ao, Ad, m, n = _build(tb, rng, kind, dtype, hi)
        rdt = np.zeros(1, dtype).real.dtype
        cdt = np.result_type(dtype, np.complex64)
        xdt = cdt if rng.random() < 0.4 else rdt
        ncols = int(rng.integers(1, 8))
        tol = 1e-5 if single else 1e-12
        tag = "seed %d case %d: %s %s m=%d n=%d cols=%d x=%s" % (seed, case, kind, np.dtype(dtype).name, m, n, ncols,
                                                                np.dtype(xdt).name)
        X = _rand(rng, (n, ncols), xdt)
        want = O.matvec(Ao, X)
        # plain product, padded leading dimension on the input
        got = tb.mul(Ad, _devmat_padded(X, int(rng.integers(0, 9)))).cpu().numpy()
        amp = _amp(Ao, m, n, X, want)
        assert _rel(got, want) <= tol * amp, tag + " (amp %.2f)" % amp
        # vector form (takes the direct product when m + n - 1 < 512, src/linearalgebra.jl:19-34)
        gv = tb.mul(Ad, _dev(X[:, 0].copy())).cpu().numpy()
        amp0 = _amp(Ao, m, n, X[:, 0], want[:, 0])
        assert _rel(gv, want[:, 0]) <= tol * amp0, tag + " (vector, amp %.2f)" % amp0
        # alpha / beta epilogue into a padded y through a factorization (src/linearalgebra.jl:42-80)
        ydt = want.dtype
        alpha = float(rng.standard_normal())
        beta = float(rng.standard_normal())
        Y0 = _rand(rng, (m, ncols), ydt)
        Yd = _devmat_padded(Y0, int(rng.integers(0, 9)))
        F = tb.factorize(Ad)
        tb.mul_(Yd, F, _devmat_padded(X, 0), alpha, beta)
        ref = alpha * want + beta * Y0
        # alpha*A*x + beta*y can cancel: the error is relative to the two terms, not to their sum
        cancel = (abs(alpha) * np.linalg.norm(want) + abs(beta) * np.linalg.norm(Y0)) / max(np.linalg.norm(ref), 1e-300)
        assert _rel(Yd.cpu().numpy(), ref) <= tol * max(1.0, cancel) * amp, tag + " (alpha, beta; cancellation %.2f)" % cancel
        # the padding rows of y were not touched
        base = Yd._base if Yd._base is not None else Yd
        assert bool(torch.isnan(torch.view_as_real(base) if base.is_complex() else base).any()) == (Yd.stride(1) > m), tag


@pytest.mark.parametrize("seed", _seeds(4))
def test_circulant_solves_random_sizes(tb, seed):
    """ldiv!/inv/eigvals of Circulant (src/linearalgebra.jl:225-253,286-287) at random sizes: power-of-two plans
    (one-level and two-level) and chirp-z plans."""
    rng = np.random.default_rng(2000 + seed)
    for case in range(10):
        dtype = [np.float64, np.complex128, np.complex64][int(rng.integers(3))]
        n = max(1, _size(rng, 40000))
        if case % 3 == 0:
            n = 1 << int(rng.integers(0, 16))
        v = _rand(rng, n, dtype)
        v[0] += 4.0 * np.sqrt(n)               # keep the spectrum away from zero (|eig| >= ~sqrt(n))
        Co, Cd = O.Circulant(v), tb.Circulant(_dev(v))
        ncols = int(rng.integers(1, 6))
        B = _rand(rng, (n, ncols), dtype)
        Bo = np.asfortranarray(B.copy())
        O.ldiv_mat(Co, Bo, O.circ_ldiv)
        Bd = _devmat_padded(B, int(rng.integers(0, 5)))
        tb.ldiv_(Cd, Bd)
        tag = "fuzz circulant seed %d case %d: %s n=%d cols=%d" % (seed, case, np.dtype(dtype).name, n, ncols)
        lam = np.fft.fft(v.astype(np.complex128))
        kC = PT.cond_spectrum(lam)             # ldiv / inv divide by the spectrum: forward error ~ cond * eps on both sides
        PT.check(tag + " ldiv", Bd.cpu().numpy(), Bo, dtype, kC)
        PT.check(tag + " eigvals", tb.eigvals(Cd).cpu().numpy(), lam, dtype)
        PT.check(tag + " inv", tb.inv(Cd).vc.cpu().numpy(), O.circ_inv(Co).vc, dtype, kC)


@pytest.mark.parametrize("seed", _seeds(4))
def test_levinson_durbin_random_sizes(tb, seed):
    """Batched levinson!/durbin! (src/directLinearSolvers.jl:15-28,100-134) at random orders and batch sizes,
    one warp or lane group per system (n <= 512), two / four warps (n <= 2048), one CTA per system in shared memory or
    in an L2-resident workspace (beyond), against the C oracle."""
    from oracle import build_oracle as OC
    rng = np.random.default_rng(3000 + seed)
    for case in range(8):
        big = case == 7                             # one order per seed beyond the register kernels (one CTA per system)
        n = int(rng.integers(2049, 9000)) if big else max(2, _size(rng, 2048))
        nsys = int(rng.integers(1, 12 if big else 70))
        R = np.asfortranarray(rng.random((n, nsys)) / n)
        Bm = np.asfortranarray(rng.standard_normal((n, nsys)))
        tag = "seed %d case %d: n=%d nsys=%d" % (seed, case, n, nsys)
        Y = tb.durbin(_devmat_padded(R, int(rng.integers(0, 4)))).cpu().numpy()
        assert _rel(Y, OC.c_durbin_batched(R)) <= 1e-12, tag + " (durbin)"
        Rl = np.asfortranarray(R[: n - 1, :])
        X = tb.colmajor(n, nsys)
        tb.levinson_(X, _devmat_padded(Rl, int(rng.integers(0, 4))), _devmat_padded(Bm, int(rng.integers(0, 4))))
        assert _rel(X.cpu().numpy(), OC.c_levinson_batched(Rl, Bm)) <= 1e-12, tag + " (levinson)"
        if n <= 512:                              # one matrix, many right-hand sides
            vc = np.concatenate([[1.7], 1.7 * R[: n - 1, 0]])
            Xm = tb.levinson(tb.SymmetricToeplitz(_dev(vc)), _devmat_padded(Bm, 0)).cpu().numpy()
            Rrep = np.asfortranarray(np.repeat(Rl[:, :1], nsys, axis=1))
            assert _rel(Xm, OC.c_levinson_batched(Rrep, Bm) / 1.7) <= 1e-12, tag + " (multi-rhs)"


@pytest.mark.parametrize("seed", _seeds(3))
def test_mul_host_random_shapes(tb, seed):
    """`tb200_mul_host` (host matrices in, host matrices out through the three-slot copy / compute / copy pipeline) at
    random shapes: pinned or pageable memory, padded leading dimensions, odd column counts, chunk sizes that cut the
    columns into one to many chunks, alpha / beta, all four element types."""
    rng = np.random.default_rng(4000 + seed)
    for case in range(6):
        kind = ["toeplitz", "symmetric", "hankel", "lower"][int(rng.integers(4))]
        dtype = DTYPES[int(rng.integers(len(DTYPES)))]
        single = np.dtype(dtype) in (np.dtype(np.float32), np.dtype(np.complex64))
        Ao, Ad, m, n = _build(tb, rng, kind, dtype, 30000 if case % 3 == 0 else 4000)
        ncols = int(rng.integers(1, 24))
        padx, pady = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        X = _rand(rng, (n, ncols), dtype)
        Y0 = _rand(rng, (m, ncols), dtype)
        tdt = torch.from_numpy(X[:0]).dtype
        Xb = torch.zeros((ncols, n + padx), dtype=tdt)
        Yb = torch.full((ncols, m + pady), float("nan"), dtype=tdt)
        if rng.random() < 0.5:
            Xb, Yb = Xb.pin_memory(), Yb.pin_memory()
        Xb[:, :n] = torch.from_numpy(np.ascontiguousarray(X.T))
        Yb[:, :m] = torch.from_numpy(np.ascontiguousarray(Y0.T))
        Xh, Yh = Xb.t()[:n, :], Yb.t()[:m, :]
        with_beta = rng.random() < 0.5
        alpha = float(rng.standard_normal()) if with_beta else 1.0
        beta = float(rng.standard_normal()) if with_beta else 0.0
        col_bytes = max(n, m) * np.dtype(dtype).itemsize
        chunk = int(rng.choice([1 << 20, 3 * col_bytes, 8 * col_bytes, 256 << 20]))
        tag = "host seed %d case %d: %s %s m=%d n=%d cols=%d pad=(%d,%d) chunk=%d beta=%s" % (
            seed, case, kind, np.dtype(dtype).name, m, n, ncols, padx, pady, chunk, with_beta)
        F = tb.factorize(Ad)
        tb.set_option("host_chunk_bytes", chunk)
        try:
            tb.mul_host_(Yh, F, Xh, alpha, beta)
        finally:
            tb.set_option("host_chunk_bytes", 256 << 20)
        want = O.matvec(Ao, X)
        ref = alpha * want + beta * Y0
        tol = 1e-5 if single else 1e-12
        cancel = (abs(alpha) * np.linalg.norm(want) + abs(beta) * np.linalg.norm(Y0)) / max(np.linalg.norm(ref), 1e-300)
        amp = _amp(Ao, m, n, X, want)
        assert _rel(Yh.numpy(), ref) <= tol * max(1.0, cancel) * amp, tag
        if pady:
            assert bool(torch.isnan(Yb[:, m:]).all()), tag + " (padding)"


@pytest.mark.parametrize("seed", _seeds(3))
def test_circulant_algebra_and_triangular_inverse_random_sizes(tb, seed):
    """pinv / sqrt / C1*C2 of Circulants (src/linearalgebra.jl:194-222,276-284,311-315) and inv of triangular Toeplitz
    matrices by recursive doubling (:337-396) at random sizes (power-of-two and chirp-z plans; every doubling depth)."""
    rng = np.random.default_rng(5000 + seed)
    for case in range(6):
        dtype = [np.float64, np.complex128, np.complex64][int(rng.integers(3))]
        n = max(2, _size(rng, 20000))
        if case % 3 == 0:
            n = 1 << int(rng.integers(1, 15))
        v = _rand(rng, n, dtype)
        v[0] += 4.0 * np.sqrt(n)
        w = _rand(rng, n, dtype)
        Co, Cd = O.Circulant(v), tb.Circulant(_dev(v))
        Wo, Wd = O.Circulant(w), tb.Circulant(_dev(w))
        tag = "fuzz circulant algebra seed %d case %d: %s n=%d" % (seed, case, np.dtype(dtype).name, n)
        kC = PT.cond_spectrum(np.fft.fft(v.astype(np.complex128)))
        PT.check(tag + " pinv", tb.pinv(Cd).vc.cpu().numpy(), O.circ_pinv(Co).vc, dtype, kC)
        PT.check(tag + " C*W", tb.circ_mul(Cd, Wd).vc.cpu().numpy(), O.circ_mul(O.factorize(Co), O.factorize(Wo)).vc, dtype)
        if np.dtype(dtype).kind == "c":              # sqrt of a real Circulant with negative eigenvalues is an InexactError (:311-315)
            # square roots of eigenvalues next to the branch cut differ by sign between two roundings of the same
            # spectrum: compare the squares (sqrt(C)^2 == C), as the reference's own test does (test/runtests.jl:573-576)
            S = tb.sqrt(Cd)
            PT.check(tag + " sqrt(C)^2", tb.circ_mul(S, S).vc.cpu().numpy(), v, dtype, np.sqrt(kC))
    for case in range(4):
        dtype = [np.float64, np.complex128][int(rng.integers(2))]
        n = max(1, _size(rng, 6000))
        r = float(rng.uniform(0.3, 0.8))
        v = np.concatenate([[1.0], 0.5 * r ** np.arange(1, n)]).astype(dtype)
        if np.dtype(dtype).kind == "c":
            v = v * np.exp(0.3j * np.arange(n))
        # symbol 1 + 0.5 z / (1 - r z) on |z| = 1 bounds the condition number of every leading block
        z = np.exp(2j * np.pi * np.arange(2048) / 2048)
        f = 1.0 + 0.5 * z / (1.0 - r * z)
        kL = float(np.abs(f).max() / np.abs(f).min())
        tag = "fuzz tri inv seed %d case %d: %s n=%d r=%.2f" % (seed, case, np.dtype(dtype).name, n, r)
        got = tb.inv(tb.LowerTriangularToeplitz(_dev(v))).v.cpu().numpy()
        PT.check(tag, got, O.tri_inv_lower(v), dtype, kL)
        # the definition: L(v) * L(inv) = I, checked on the first column through the oracle's own product
        e0 = np.zeros(n, dtype); e0[0] = 1
        PT.check(tag + " L*inv(L) e0", O.matvec(O.LowerTriangularToeplitz(v), got), e0, dtype, kL)
