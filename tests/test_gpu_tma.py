"""TMA-staged strided passes (csrc/tma_cols.cuh, option "tma"): bit-for-bit the same products as the LDG/STG kernels
on the plans of the BASELINE configs, including odd column counts (tensor-map out-of-bounds fill / clipping), Hankel
index reversal, rectangular shapes and the cases that must fall back (beta != 0, n not a multiple of N2)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def _run(tb, F, X, tma, ydtype=None, beta=None, Y0=None):
    tb.set_option("tma", tma)
    try:
        m = F.shape[0]
        Y = tb.colmajor(m, X.shape[1], ydtype or X.dtype) if X.dim() == 2 else torch.empty(m, dtype=ydtype or X.dtype, device="cuda")
        if Y0 is not None:
            Y.copy_(Y0)
        tb.mul_(Y, F, X, True, beta if beta is not None else False)
        torch.cuda.synchronize()
        return Y
    finally:
        tb.set_option("tma", 0)


@pytest.mark.parametrize("log2n,cols", [(17, 5), (20, 4), (19, 1)])
def test_tma_real_pairs_match(tb, log2n, cols):
    """Toeplitz{Float64}: plan 2^9 x 2^(l-9), real column pairs (config 2 at log2n = 20)."""
    n = 1 << log2n
    g = torch.Generator(device="cuda"); g.manual_seed(log2n)
    k = torch.arange(n, dtype=torch.float64, device="cuda")
    F = tb.factorize(tb.Toeplitz(0.9 ** k, 0.4 ** k))
    X = torch.randn((cols, n), dtype=torch.float64, device="cuda", generator=g).t()
    l0 = tb.launch_count()
    ref = _run(tb, F, X, 0)
    for mode in (1, 2, 3):
        out = _run(tb, F, X, mode)
        assert torch.equal(out, ref), (mode, float((out - ref).abs().max()))
    # beta != 0 falls back for the store side and still matches
    Y0 = torch.randn((cols, n), dtype=torch.float64, device="cuda", generator=g).t()
    assert torch.equal(_run(tb, F, X, 3, beta=0.5, Y0=Y0), _run(tb, F, X, 0, beta=0.5, Y0=Y0))
    assert tb.launch_count() > l0


def test_tma_rectangular_and_unaligned_fall_back(tb):
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    m, n = (1 << 17) + 4096 * 3, 1 << 17          # m a multiple of N2 = 2^10?  plan decides; either way results must match
    vc = torch.randn(m, dtype=torch.float64, device="cuda", generator=g)
    vr = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    vr[0] = vc[0]
    F = tb.factorize(tb.Toeplitz(vc, vr))
    X = torch.randn((3, n), dtype=torch.float64, device="cuda", generator=g).t()
    assert torch.equal(_run(tb, F, X, 3), _run(tb, F, X, 0))
    n2 = (1 << 17) - 7                               # not a multiple of N2: LDG path
    k = torch.arange(n2, dtype=torch.float64, device="cuda")
    F2 = tb.factorize(tb.Toeplitz(0.9 ** k, 0.4 ** k))
    X2 = torch.randn((2, n2), dtype=torch.float64, device="cuda", generator=g).t()
    assert torch.equal(_run(tb, F2, X2, 3), _run(tb, F2, X2, 0))


def test_tma_complex_circulant_config3_plan(tb):
    """Circulant{ComplexF64} n = 2^22 (plan 2^10 x 2^12, no padding: full tiles), 3 right-hand sides."""
    n = 1 << 22
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    v = (0.9 ** torch.arange(n, dtype=torch.float64, device="cuda")).to(torch.complex128)
    F = tb.factorize(tb.Circulant(v))
    B = torch.randn((3, n), dtype=torch.complex128, device="cuda", generator=g).t()
    ref = _run(tb, F, B, 0)
    for mode in (1, 2, 3):
        assert torch.equal(_run(tb, F, B, mode), ref), mode
    tb.set_option("tma", 3)
    try:
        X1 = tb.colmajor(n, 3, torch.complex128)
        tb.circ_ldiv_(F, B, out=X1)
    finally:
        tb.set_option("tma", 0)
    X0 = tb.colmajor(n, 3, torch.complex128)
    tb.circ_ldiv_(F, B, out=X0)
    assert torch.equal(X1, X0)


@pytest.mark.parametrize("log2n", [23, 24])
def test_tma_hankel_complex32_config5_plan(tb, log2n):
    """Hankel{ComplexF32}: plan 2^11 x 2^(l-11), reversed input indices inside the staged tile (config 5a at 24)."""
    n = 1 << log2n
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    v = torch.randn(2 * n - 1, dtype=torch.complex64, device="cuda", generator=g)
    x = torch.randn(n, dtype=torch.complex64, device="cuda", generator=g)
    F = tb.factorize(tb.Hankel(v, (n, n)))
    ref = _run(tb, F, x, 0)
    for mode in (1, 2, 3):
        assert torch.equal(_run(tb, F, x, mode), ref), mode
