"""Fused "pass C of one group + pass A of the next" launches (k_colsCA, option "ca_fuse": 1 = fused, 2 = fused with the
asynchronously staged x tile): bit-for-bit the products of the separate passes -- the arithmetic is the same, only the
launch structure changes -- for real column pairs, odd column counts, alpha/beta, complex operands, Hankel index
reversal, rectangular shapes, Float32 and the Circulant ldiv! path (no padding: full tiles), with workspaces small
enough that every product runs many groups on both overlap streams."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def _with(tb, fuse, group_mb, fn):
    tb.set_option("ca_fuse", fuse)
    tb.set_option("group_bytes", group_mb << 20)
    try:
        out = fn()
        torch.cuda.synchronize()
        return out
    finally:
        tb.set_option("ca_fuse", DEFAULT_FUSE)
        tb.set_option("group_bytes", 32 << 20)


DEFAULT_FUSE = 0          # restored after every case; the library default is set in capi.cu


def _mul(tb, F, X, alpha=True, beta=False, Y0=None, ydtype=None):
    def run():
        m = F.shape[0]
        Y = tb.colmajor(m, X.shape[1], ydtype or X.dtype)
        if Y0 is not None:
            Y.copy_(Y0)
        tb.mul_(Y, F, X, alpha, beta)
        return Y
    return run


@pytest.mark.parametrize("log2n,cols,group_mb", [(14, 16, 1), (14, 11, 1), (16, 9, 2), (20, 6, 32), (15, 33, 1)])
def test_fused_real_pairs(tb, log2n, cols, group_mb):
    n = 1 << log2n
    g = torch.Generator(device="cuda"); g.manual_seed(log2n * 100 + cols)
    k = torch.arange(n, dtype=torch.float64, device="cuda")
    F = tb.factorize(tb.Toeplitz(0.9 ** k, 0.4 ** k))
    X = torch.randn((cols, n), dtype=torch.float64, device="cuda", generator=g).t()
    Y0 = torch.randn((cols, n), dtype=torch.float64, device="cuda", generator=g).t()
    l0 = tb.launch_count()
    ref = _with(tb, 0, group_mb, _mul(tb, F, X))
    l_sep = tb.launch_count() - l0
    refb = _with(tb, 0, group_mb, _mul(tb, F, X, 1.5, -0.25, Y0))
    for fuse in (1, 2):
        l0 = tb.launch_count()
        out = _with(tb, fuse, group_mb, _mul(tb, F, X))
        l_fused = tb.launch_count() - l0
        assert torch.equal(out, ref), (fuse, float((out - ref).abs().max()))
        assert l_fused < l_sep, (fuse, l_fused, l_sep)           # the fused launches really ran
        outb = _with(tb, fuse, group_mb, _mul(tb, F, X, 1.5, -0.25, Y0))
        assert torch.equal(outb, refb), fuse


@pytest.mark.parametrize("dtype", [torch.complex128, torch.float32, torch.complex64])
def test_fused_other_dtypes(tb, dtype):
    n = 1 << 15
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    real = torch.float64 if dtype == torch.complex128 else torch.float32
    k = torch.arange(n, dtype=real, device="cuda")
    F = tb.factorize(tb.Toeplitz((0.9 ** k).to(dtype), (0.4 ** k).to(dtype)))
    X = torch.randn((13, n), dtype=dtype, device="cuda", generator=g).t()
    ref = _with(tb, 0, 1, _mul(tb, F, X))
    for fuse in (1, 2):
        assert torch.equal(_with(tb, fuse, 1, _mul(tb, F, X)), ref), fuse


def test_fused_hankel_rectangular_and_circulant(tb):
    g = torch.Generator(device="cuda"); g.manual_seed(12)
    n = 1 << 14
    v = torch.randn(2 * n - 1, dtype=torch.float64, device="cuda", generator=g)
    FH = tb.factorize(tb.Hankel(v, (n, n)))
    X = torch.randn((10, n), dtype=torch.float64, device="cuda", generator=g).t()
    ref = _with(tb, 0, 1, _mul(tb, FH, X))
    for fuse in (1, 2):
        assert torch.equal(_with(tb, fuse, 1, _mul(tb, FH, X)), ref), ("hankel", fuse)
    m = n + 1000                                     # rectangular: m > N'/2 is impossible here, but m != n changes the store clip
    vc = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    vr = torch.randn(n - 300, dtype=torch.float64, device="cuda", generator=g)
    vr[0] = vc[0]
    FR = tb.factorize(tb.Toeplitz(vc, vr))
    XR = torch.randn((9, n - 300), dtype=torch.float64, device="cuda", generator=g).t()
    ref = _with(tb, 0, 1, _mul(tb, FR, XR))
    for fuse in (1, 2):
        assert torch.equal(_with(tb, fuse, 1, _mul(tb, FR, XR)), ref), ("rect", fuse)
    nc = 1 << 15                                     # Circulant ldiv!: no padding on either side (full tiles)
    c = (0.9 ** torch.arange(nc, dtype=torch.float64, device="cuda")).to(torch.complex128)
    FC = tb.factorize(tb.Circulant(c))
    B = torch.randn((11, nc), dtype=torch.complex128, device="cuda", generator=g).t()

    def ldiv():
        Xo = tb.colmajor(nc, 11, torch.complex128)
        tb.circ_ldiv_(FC, B, out=Xo)
        return Xo
    ref = _with(tb, 0, 1, ldiv)
    for fuse in (1, 2):
        assert torch.equal(_with(tb, fuse, 1, ldiv), ref), ("circulant", fuse)
