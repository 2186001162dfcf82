"""BASELINE.json configurations at (or near) full size: size-independent properties over the whole
batch plus oracle parity on sampled columns/systems.  Tolerances: 1e-12 (Float64), 1e-5 (Float32)."""
import numpy as np
import pytest
import torch

from oracle import toeplitz_oracle as O
from oracle import build_oracle as OC
import parity_tools as PT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())


def trel(a, b):
    return float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))


def test_c2_batched_toeplitz_f64(tb):
    """Config 2: Toeplitz{Float64} n = 2^20 x 1024 columns (256 here to bound test memory/time):
    sampled columns vs the oracle incl. shard edges, linearity over the whole batch, odd column count."""
    n, l = 1 << 20, 256
    k = np.arange(n, dtype=np.float64)
    vc, vr = 0.9 ** k, 0.4 ** k
    A = tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda())
    F = tb.factorize(A)
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn((l, n), dtype=torch.float64, device="cuda", generator=g).t()
    Y = tb.colmajor(n, l)
    tb.mul_(Y, F, X)
    Fo = O.factorize(O.Toeplitz(vc, vr))
    for j in (0, 1, 127, 128, 254, 255):
        yo = O.mul_factorization(np.zeros(n), Fo, X[:, j].cpu().numpy())
        assert rel(Y[:, j].cpu().numpy(), yo) <= 1e-12
    # linearity: A*(X[:, :128] + 2 X[:, 128:]) == Y[:, :128] + 2 Y[:, 128:]
    Z = tb.colmajor(n, 128)
    tb.mul_(Z, F, (X[:, :128] + 2 * X[:, 128:]).t().contiguous().t())
    assert trel(Z, Y[:, :128] + 2 * Y[:, 128:]) <= 1e-12
    # odd column count + alpha/beta on a strided (ld > n) view
    Y3 = Y[:, :3].clone().t().contiguous().t()
    tb.mul_(Y3, F, X[:, 5:8], 2.0, -1.0)
    assert trel(Y3, 2 * Y[:, 5:8] - Y[:, :3]) <= 1e-12


def test_c3_circulant_c64(tb):
    """Config 3: Circulant{ComplexF64} n = 2^22: factorize, eigvals, ldiv! over RHS columns, inv."""
    n, l = 1 << 22, 16
    v = (0.9 ** np.arange(n)).astype(np.complex128)
    C = tb.Circulant(torch.from_numpy(v).cuda())
    F = tb.factorize(C)
    lam = tb.eigvals(F)
    lam_o = np.fft.fft(v)
    assert rel(lam.cpu().numpy(), lam_o) <= 1e-12
    g = torch.Generator(device="cuda").manual_seed(2)
    B = torch.randn((l, n), dtype=torch.complex128, device="cuda", generator=g).t()
    X = B.clone()
    tb.ldiv_(F, X)
    for j in (0, l - 1):
        xo = np.fft.ifft(np.fft.fft(B[:, j].cpu().numpy()) / lam_o)
        assert rel(X[:, j].cpu().numpy(), xo) <= 1e-12
    # round trip: C * (C \ B) == B over the whole batch
    R = tb.mul(F, X)
    assert trel(R, B) <= 1e-12
    Ci = tb.inv(C)
    PT.check("C3 inv(C).vc n=2^22", Ci.vc.cpu().numpy(), np.fft.ifft(1.0 / lam_o), np.complex128, PT.cond_spectrum(lam_o))


def test_c4_levinson_durbin_many_systems(tb):
    """Config 4: SymmetricToeplitz{Float64} n = 512, 2*10^5 independent systems (10^6 in bench):
    sampled systems vs the C oracle, permutation invariance of the batch."""
    n, nsys = 512, 200000
    g = torch.Generator(device="cuda").manual_seed(3)
    R = (torch.rand((nsys, n), dtype=torch.float64, device="cuda", generator=g) / n).t()
    B = torch.randn((nsys, n), dtype=torch.float64, device="cuda", generator=g).t()
    X = tb.colmajor(n, nsys)
    tb.levinson_(X, R[: n - 1, :], B)
    Yd = tb.durbin(R)
    idx = [0, 1, 31, 32, 77777, nsys - 1]
    Rs = np.asfortranarray(R[:, idx].cpu().numpy())
    Bs = np.asfortranarray(B[:, idx].cpu().numpy())
    assert rel(X[:, idx].cpu().numpy(), OC.c_levinson_batched(np.asfortranarray(Rs[: n - 1]), Bs)) <= 1e-12
    assert rel(Yd[:, idx].cpu().numpy(), OC.c_durbin_batched(Rs)) <= 1e-12
    perm = torch.randperm(nsys, device="cuda", generator=g)[:4096]
    Xp = tb.colmajor(n, 4096)
    tb.levinson_(Xp, R[: n - 1, perm].t().contiguous().t(), B[:, perm].t().contiguous().t())
    assert torch.equal(Xp, X[:, perm])
    # KMS set (well conditioned): r_k = rho^k
    rho = 0.1 + 0.8 * torch.rand(1000, dtype=torch.float64, device="cuda", generator=g)
    Rk = (rho[None, :] ** torch.arange(1, n, dtype=torch.float64, device="cuda")[:, None]).t().contiguous().t()
    Bk = B[:, :1000].t().contiguous().t()
    Xk = tb.colmajor(n, 1000)
    tb.levinson_(Xk, Rk, Bk)
    # KMS(rho) has its spectrum inside ((1-rho)/(1+rho), (1+rho)/(1-rho)): cond < ((1+rho)/(1-rho))^2, worst rho here 0.9
    kK = float(((1 + rho.max()) / (1 - rho.max())) ** 2)
    PT.check("C4 KMS levinson, 1000 systems", Xk.cpu().numpy(),
             OC.c_levinson_batched(np.asfortranarray(Rk.cpu().numpy()), np.asfortranarray(Bk.cpu().numpy())), np.float64, kK, c=2.0)


def test_c5a_hankel_c32(tb):
    """Config 5a: Hankel{ComplexF32} n = 2^24 matvec (embedding 2^25, two-level Float32 plan)."""
    n = 1 << 24
    rng = np.random.default_rng(5)
    v = ((rng.standard_normal(2 * n - 1) + 1j * rng.standard_normal(2 * n - 1)) / np.sqrt(2)).astype(np.complex64)
    x = ((rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)).astype(np.complex64)
    H = tb.Hankel(torch.from_numpy(v).cuda(), (n, n))
    y = tb.mul(H, torch.from_numpy(x).cuda())
    # oracle in double precision on the same Float32 inputs (reference path would run in ComplexF32)
    yo = O.matvec(O.Hankel(v.astype(np.complex128), (n, n)), x.astype(np.complex128))
    assert rel(y.cpu().numpy(), yo) <= 1e-5


def test_c5b_cgs_strang_c32(tb):
    """Config 5b: SymmetricToeplitz cgs + Strang in ComplexF32.  Oracle parity at n = 2^20 for a fixed
    iteration count; at n = 2^24 the returned residual must match a recomputed ||b - A x||/||b||."""
    for n, check_oracle in ((1 << 20, True), (1 << 24, False)):
        v = (0.5 ** np.arange(n)).astype(np.complex64)
        rng = np.random.default_rng(7)
        b = rng.standard_normal(n).astype(np.complex64)
        A = tb.SymmetricToeplitz(torch.from_numpy(v).cuda())
        M = tb.factorize(tb.strang(A))
        bd = torch.from_numpy(b).cuda()
        x, err, it, flag = tb.cgs(A, torch.zeros_like(bd), bd, M, 4, 0.0)
        assert (it, flag) == (4, 1)
        r = bd - tb.mul(A, x)
        err2 = float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(bd))
        kA = 9.0                                   # KMS(0.5): cond < ((1+0.5)/(1-0.5))^2
        eps32 = PT.eps_of(np.float32)
        # the recurrence residual tracks the true one down to the Float32 floor of ||A|| ||x|| eps ~ cond * eps
        assert abs(err - err2) <= 8 * kA * eps32, (err, err2)
        assert err2 <= 8 * kA * eps32, err2
        if check_oracle:
            Ao = O.SymmetricToeplitz(v)
            xo, eo, ito, fo = O.cgs(Ao, np.zeros(n, np.complex64), b.copy(), O.factorize(O.strang(Ao)), 4, 0.0)
            assert (ito, fo) == (it, flag)
            eo2 = float(np.linalg.norm(b.astype(np.complex128) - O.matvec(O.SymmetricToeplitz(v.astype(np.complex128)), xo.astype(np.complex128)))
                        / np.linalg.norm(b))
            PT.check("C5b cgs 4 iterations n=2^20 vs oracle", x.cpu().numpy(), xo, np.complex64, kA, extra=2.0 * kA * (eo2 + err2))


def test_largest_float64_embedding(tb):
    """Float64 n = 2^24 (embedding 2^25 = 2^12 x 2^13, the largest Float64 two-level plan): one vector vs the oracle."""
    n = 1 << 24
    rng = np.random.default_rng(9)
    vc = rng.standard_normal(n)
    vr = rng.standard_normal(n)
    vr[0] = vc[0]
    x = rng.standard_normal(n)
    y = tb.mul(tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda()), torch.from_numpy(x).cuda())
    yo = O.matvec(O.Toeplitz(vc, vr), x, workers=-1)
    assert rel(y.cpu().numpy(), yo) <= 1e-12
    with pytest.raises(tb.ToeplitzB200Error):
        tb.factorize_empty_like(1, torch.float64, ((1 << 30) + 1, (1 << 30) + 1))    # embedding 2^32: beyond the three-level plan, refused


@pytest.mark.parametrize("dtype,log2n", [("float64", 25), ("complex64", 26)])
def test_three_level_plan(tb, dtype, log2n):
    """Embeddings beyond the two-level limit (2^26 in Float64, 2^27 in Float32): N' = 2^9 x 2^(l-9), the rows are a
    batched two-level transform on a child plan.  One vector against the oracle (reference FFT length m+n-1), a
    three-column product (pair packing + odd tail) against single products, alpha/beta, and the adjoint-free
    Circulant family at 2^26 / 2^27 points: eigvals == fft(vc), C * ldiv(C, b) == b."""
    n = 1 << log2n
    dt = np.dtype(dtype)
    rng = np.random.default_rng(31)
    cplx = dt.kind == "c"
    mk = lambda k: (rng.standard_normal(k) + (1j * rng.standard_normal(k) if cplx else 0)).astype(dt)
    vc, vr, x = mk(n), mk(n), mk(n)
    vr[0] = vc[0]
    A = tb.Toeplitz(torch.from_numpy(vc).cuda(), torch.from_numpy(vr).cuda())
    F = tb.factorize(A)
    xd = torch.from_numpy(x).cuda()
    y = tb.mul(F, xd)
    if dt == np.float64:
        # the oracle proper: reference algorithm at its own FFT length m+n-1 on the host
        yo = O.matvec(O.Toeplitz(vc, vr), x, workers=-1)
    else:
        # Float32 case at 2^27 points: an independent double-precision evaluation of the same circulant embedding on the
        # device (torch.fft = cuFFT, test-side only) keeps the suite's run time in bounds
        c = torch.zeros(2 * n, dtype=torch.complex128, device="cuda")
        c[:n] = torch.from_numpy(vc).cuda()
        c[n + 1:] = torch.from_numpy(vr[1:][::-1].copy()).cuda()
        xp = torch.zeros(2 * n, dtype=torch.complex128, device="cuda")
        xp[:n] = xd
        yo = torch.fft.ifft(torch.fft.fft(c) * torch.fft.fft(xp))[:n].cpu().numpy()
        del c, xp
    PT.check("three-level mul n=2^%d %s" % (log2n, dtype), y.cpu().numpy(), yo, dt)
    del yo
    X3 = tb.colmajor(n, 3, xd.dtype)
    X3[:, 0].copy_(xd); X3[:, 1].copy_(2 * xd); X3[:, 2].copy_(-xd)
    Y3 = tb.colmajor(n, 3, xd.dtype)
    Y3.fill_(1.0)
    tb.mul_(Y3, F, X3, 0.5, 2.0)
    for j, sc in enumerate((1.0, 2.0, -1.0)):
        ref = 0.5 * sc * y + 2.0
        assert float(torch.linalg.vector_norm(Y3[:, j] - ref) / torch.linalg.vector_norm(ref)) <= PT.ns_tol(dt), j
    del F, A, X3, Y3, y
    torch.cuda.empty_cache()
    # Circulant with exactly 2^(log2n+1) points
    nc = 2 * n
    cd = np.complex128 if dt.itemsize * (1 if cplx else 2) == 16 else np.complex64
    v = (0.9 ** np.arange(nc)).astype(cd)
    v[1] += 0.25
    C = tb.Circulant(torch.from_numpy(v).cuda())
    FC = tb.factorize(C)
    lam = tb.eigvals(FC).cpu().numpy()
    lam_o = np.fft.fft(v.astype(np.complex128))
    PT.check("three-level eigvals n=2^%d" % (log2n + 1), lam, lam_o, cd)
    b = torch.from_numpy((rng.standard_normal(nc) + 1j * rng.standard_normal(nc)).astype(cd)).cuda()
    xs = tb.ldiv(FC, b)
    r = tb.mul(FC, xs)
    assert float(torch.linalg.vector_norm(r - b) / torch.linalg.vector_norm(b)) <= PT.bound(cd, PT.cond_spectrum(lam_o))
    del FC, C, xs, r, b
    torch.cuda.empty_cache()
    if dt == np.float64:
        # cgs + Strang on a three-level operator (the Krylov driver takes its unfused route there): true residual
        vk = torch.from_numpy(0.5 ** np.arange(n)).cuda()
        S = tb.SymmetricToeplitz(vk)
        bb = torch.from_numpy(rng.standard_normal(n)).cuda()
        xk, err, it, flag = tb.cgs(S, torch.zeros_like(bb), bb, tb.factorize(tb.strang(S)), 20, 1e-13)
        assert flag == 0 and it <= 6, (flag, it, err)
        rr = bb - tb.mul(S, xk)
        assert float(torch.linalg.vector_norm(rr) / torch.linalg.vector_norm(bb)) <= 1e-12


@pytest.mark.parametrize("n", [2, 3, 5, 16, 17, 33, 64, 65, 100, 128, 129, 200, 256, 257, 300, 384, 385, 500, 511, 512])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_levinson_durbin_both_kernel_families(tb, n, dtype):
    """Generator-form kernels (csrc/levinson_gen.cu, the default for n <= 512) and the dot-product kernels (option
    lev_gen=0) against the C oracle of src/directLinearSolvers.jl:15-28,100-134 on the same systems, incl. the diag
    normalisation of levinson(T, b), at the north_star tolerances."""
    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def devmat(a):
        return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()

    dt = np.dtype(dtype)
    tol = 1e-12 if dt == np.float64 else 1e-5
    rng = np.random.default_rng(n)
    ns = 37
    R = np.asfortranarray(rng.random((max(n - 1, 0), ns)) / n).astype(dt)
    B = np.asfortranarray(rng.standard_normal((n, ns))).astype(dt)
    Rd = np.asfortranarray(rng.random((n, ns)) / n).astype(dt)
    diag = (1.0 + rng.random(ns)).astype(dt)
    xo = OC.c_levinson_batched(np.asfortranarray(R.astype(np.float64)), np.asfortranarray(B.astype(np.float64)))
    yo = OC.c_durbin_batched(np.asfortranarray(Rd.astype(np.float64)))
    for gen in (1, 0):
        try:
            tb.set_option("lev_gen", gen)
            X = tb.colmajor(n, ns, getattr(torch, dtype))
            tb.levinson_(X, devmat(R), devmat(B))
            Y = tb.colmajor(n, ns, getattr(torch, dtype))
            tb.durbin_(Y, devmat(Rd))
            Xd = tb.colmajor(n, ns, getattr(torch, dtype))
            tb.levinson_(Xd, devmat(R * diag), devmat(B), dev(diag))
        finally:
            tb.set_option("lev_gen", 1)
        assert rel(X.cpu().numpy(), xo) <= tol, ("levinson", gen)
        assert rel(Y.cpu().numpy(), yo) <= tol, ("durbin", gen)
        assert rel(Xd.cpu().numpy(), xo / diag.astype(np.float64)) <= tol, ("levinson diag", gen)


@pytest.mark.parametrize("n,dtype", [(2100, "float64"), (5000, "float64"), (6400, "float64"), (7000, "float64"), (20000, "float64"),
                                     (12000, "float32"), (30000, "float32")])
def test_levinson_durbin_any_order(tb, n, dtype):
    """Orders beyond the register kernels: one CTA per system in generator form, vectors in shared memory (n <= 6400 in
    Float64) or in an L2-resident global workspace (any n; the reference has no size limit, src/directLinearSolvers.jl:100)."""
    dt = np.dtype(dtype)
    rng = np.random.default_rng(n)
    ns = 3
    R = np.asfortranarray(rng.random((n - 1, ns)) / n).astype(dt)
    B = np.asfortranarray(rng.standard_normal((n, ns))).astype(dt)
    Rd = np.asfortranarray(rng.random((n, ns)) / n).astype(dt)
    diag = (1.0 + rng.random(ns)).astype(dt)
    dm = lambda a: torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
    X = tb.colmajor(n, ns, getattr(torch, dtype))
    tb.levinson_(X, dm(R * diag), dm(B), torch.from_numpy(diag).cuda())
    Y = tb.colmajor(n, ns, getattr(torch, dtype))
    tb.durbin_(Y, dm(Rd))
    xo = OC.c_levinson_batched(np.asfortranarray(R.astype(np.float64)), np.asfortranarray(B.astype(np.float64))) / diag.astype(np.float64)
    yo = OC.c_durbin_batched(np.asfortranarray(Rd.astype(np.float64)))
    # n recursion steps in the working precision against a Float64 oracle: rounding accumulates like a random walk,
    # ~sqrt(n) * eps (matters for Float32 only: 4 * sqrt(12000) * eps32 = 5e-5; the reference in Float32 drifts the same way)
    walk = 4.0 * np.sqrt(n) * PT.eps_of(dt)
    PT.check("levinson n=%d %s (block kernel)" % (n, dtype), X.cpu().numpy(), xo, dt, extra=walk)
    PT.check("durbin n=%d %s (block kernel)" % (n, dtype), Y.cpu().numpy(), yo, dt, extra=walk)
