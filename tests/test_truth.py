"""Definition-level parity: oracle AND CUDA path against tests/golden/truth_v1.npz.

truth_v1.npz (tests/golden/make_truth.py) holds the mathematical definitions -- dense products, dense LU solves,
the DFT sum, closed-form inverses verified by residual -- evaluated in 200-bit arithmetic (mpmath) on the reference's
deterministic recipes (test/runtests.jl:23-48,76-77,113-116,156-161,181-214,797).  It comes from no fast algorithm,
so the chain  definition -> oracle -> CUDA  is not self-referential:

* CPU tests (no marker): the oracle restatement of the reference path agrees with the definitions;
* GPU tests (`-m gpu`): the CUDA path, through the C ABI, agrees with the definitions at the north_star's tolerances
  (1e-12 Float64 / 1e-5 Float32) for products, spectra and Levinson-type direct solves; iterative and
  ill-conditioned solves are held to `c * cond(A) * eps` with cond(A) computed here and the measured error printed.
"""
import os

import numpy as np
import pytest

from oracle import toeplitz_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
T = np.load(os.path.join(HERE, "golden", "truth_v1.npz"))
EPS = np.finfo(np.float64).eps
TOL64, TOL32 = 1e-12, 1e-5


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def recipe(n):
    k = np.arange(n)
    return 0.9 ** k, 0.4 ** k, T["x_%d" % n], T["xi_%d" % n]


def oracle_matrices(n):
    vc, vr, _, _ = recipe(n)
    k = np.arange(n)
    return {
        "toeplitz": O.Toeplitz(vc, vr), "toeplitz_adj": O.Toeplitz(vc, vr).adjoint(), "symmetric": O.SymmetricToeplitz(vc),
        "circulant": O.Circulant(vc), "lower": O.LowerTriangularToeplitz(vc), "upper": O.UpperTriangularToeplitz(vc),
        "hankel": O.Hankel(0.9 ** k[::-1], vr=vr),
    }


# ======================================================================================================
# CPU: oracle vs definitions
# ======================================================================================================
@pytest.mark.parametrize("n", [101, 2000])
def test_oracle_products_match_definitions(n):
    _, _, x, xi = recipe(n)
    for name, A in oracle_matrices(n).items():
        e = rel(O.matvec(A, x), T["%s_%d" % (name, n)])
        assert e <= 20 * EPS, (name, n, e)
    vc, vr, _, _ = recipe(n)
    e = rel(O.matvec(O.Toeplitz(vc.astype(complex), vr.astype(complex)), x + 1j * xi), T["toeplitz_cplx_%d" % n])
    assert e <= 20 * EPS, e


def test_oracle_rectangular_match_definitions():
    k2, k1 = np.arange(2000), np.arange(101)
    assert rel(O.matvec(O.Toeplitz(0.9 ** k2, 0.4 ** k1), np.cos(0.37 * k1)), T["rect_2000x101"]) <= 20 * EPS
    assert rel(O.matvec(O.Toeplitz(0.9 ** k1, 0.4 ** k2), np.cos(0.37 * k2)), T["rect_101x2000"]) <= 20 * EPS


@pytest.mark.parametrize("n", [101, 2000])
def test_oracle_circulant_mathematics_match_definitions(n):
    vc, vr, x, _ = recipe(n)
    C = O.Circulant(vc)
    assert rel(O.circ_eigvals(C), T["circ_eigvals_%d" % n]) <= 50 * EPS          # fft(C.vc) == DFT definition
    assert rel(O.circ_ldiv(C, x.copy()), T["circ_ldiv_%d" % n]) <= 200 * EPS     # cond(C) ~ 19
    assert rel(O.circ_inv(C).vc, T["circ_inv_vc_%d" % n]) <= 200 * EPS
    assert rel(O.strang(O.Toeplitz(vc, vr)).vc, T["strang_%d" % n]) == 0.0
    assert rel(O.chan(O.Toeplitz(vc, vr)).vc, T["chan_%d" % n]) <= 4 * EPS
    for m in (5, 6):
        assert rel(O.circ_eigvals(O.Circulant(np.arange(1.0, m + 1))), T["circ_eigvals_small_%d" % m]) <= 20 * EPS


def test_oracle_levinson_family_match_definitions():
    b = T["kernel_b"]
    for name in ("exp", "rbf"):
        r = T["kernel_%s_r" % name]
        T8 = O.SymmetricToeplitz(np.concatenate([[1.0], r[:-1]])).dense()
        c = np.linalg.cond(T8)
        assert rel(O.durbin(r), T["durbin_" + name]) <= 8 * c * EPS, name
        assert rel(O.levinson(r[:-1], b), T["levinson_" + name]) <= 8 * c * EPS, name
        assert rel(O.trench(r[:-1]), T["trench_" + name]) <= 8 * c * EPS, name
    for n in (128, 256):
        r, bb = T["lev_r_%d" % n], T["lev_b_%d" % n]
        assert rel(O.levinson(r, bb), T["lev_x_%d" % n]) <= TOL64
        assert rel(O.durbin(T["dur_r_%d" % n]), T["dur_y_%d" % n]) <= TOL64


def test_oracle_iterative_solves_match_definitions():
    tol = 100 * EPS
    for n in (600, 4096):
        A = O.SymmetricToeplitz(0.5 ** np.arange(n))
        b = T["kms_b_%d" % n]
        x, err, it, flag = O.cgs(A, np.zeros(n), b.copy(), O.factorize(O.strang(A)), 1000, tol)
        assert flag == 0
        assert rel(x, T["kms_x_%d" % n]) <= 9 * tol * 10              # cond(KMS 0.5) = 9
    vc, vr, x101, _ = recipe(101)
    for name, A in (("toeplitz", O.Toeplitz(vc, vr)), ("symmetric", O.SymmetricToeplitz(vc)),
                    ("lower", O.LowerTriangularToeplitz(vc)), ("upper", O.UpperTriangularToeplitz(vc))):
        c = np.linalg.cond(A.dense())
        sol = O.ldiv(A, x101.copy())
        assert rel(sol, T["solve_%s_101" % name]) <= max(100 * c * EPS * 100, 1e-12), (name, c)


def test_exact_cases():
    assert np.array_equal(O.matvec(O.Toeplitz(np.array([1.0, 2]), np.array([1.0, 2])), np.ones(2)), T["exact_toeplitz_2"])
    assert rel(O.matvec(O.Hankel(np.array([1.0, 2, 3, 4, 5]), vr=np.array([5.0, 6, 7, 8, 9])), np.ones(5)), T["hankel_5x5"]) <= 4 * EPS


# ======================================================================================================
# GPU: CUDA path vs definitions
# ======================================================================================================
@pytest.fixture(scope="module")
def tb():
    import toeplitz_b200 as tb
    return tb


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def device_matrices(tb, n, dtype):
    vc, vr, _, _ = recipe(n)
    k = np.arange(n)
    vc, vr = vc.astype(dtype), vr.astype(dtype)
    return {
        "toeplitz": tb.Toeplitz(dev(vc), dev(vr)), "toeplitz_adj": tb.adjoint(tb.Toeplitz(dev(vc), dev(vr))),
        "symmetric": tb.SymmetricToeplitz(dev(vc)), "circulant": tb.Circulant(dev(vc)),
        "lower": tb.LowerTriangularToeplitz(dev(vc)), "upper": tb.UpperTriangularToeplitz(dev(vc)),
        "hankel": tb.Hankel(dev((0.9 ** k[::-1]).astype(dtype)), vr=dev(vr)),
    }


@pytest.mark.gpu
@pytest.mark.parametrize("n", [101, 2000])
def test_cuda_products_match_definitions_f64(tb, n):
    import torch
    _, _, x, xi = recipe(n)
    worst = 0.0
    for name, A in device_matrices(tb, n, np.float64).items():
        # matrix form (FFT path at every size) and vector form (direct N < 512 loop at n = 101)
        X2 = torch.from_numpy(np.stack([x, 2 * x])).cuda().t()
        e = max(rel(tb.mul(A, X2)[:, 0].cpu().numpy(), T["%s_%d" % (name, n)]),
                rel(tb.mul(A, dev(x)).cpu().numpy(), T["%s_%d" % (name, n)]))
        worst = max(worst, e)
        assert e <= TOL64, (name, n, e)
    vc, vr, _, _ = recipe(n)
    Ac = tb.Toeplitz(dev(vc.astype(complex)), dev(vr.astype(complex)))
    e = rel(tb.mul(Ac, dev(x + 1j * xi)).cpu().numpy(), T["toeplitz_cplx_%d" % n])
    assert e <= TOL64, e
    print("max rel err vs 200-bit definitions, n=%d: %.2e" % (n, max(worst, e)))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [101, 2000])
def test_cuda_products_match_definitions_f32(tb, n):
    """Float32 inputs are the Float64 recipes rounded; the truth of the rounded problem differs from the stored one by
    O(eps32), far inside 1e-5."""
    _, _, x, _ = recipe(n)
    for name, A in device_matrices(tb, n, np.float32).items():
        e = rel(tb.mul(A, dev(x.astype(np.float32))).cpu().numpy(), T["%s_%d" % (name, n)])
        assert e <= TOL32, (name, n, e)


@pytest.mark.gpu
def test_cuda_rectangular_match_definitions(tb):
    k2, k1 = np.arange(2000), np.arange(101)
    A = tb.Toeplitz(dev(0.9 ** k2), dev(0.4 ** k1))
    assert rel(tb.mul(A, dev(np.cos(0.37 * k1))).cpu().numpy(), T["rect_2000x101"]) <= TOL64
    A = tb.Toeplitz(dev(0.9 ** k1), dev(0.4 ** k2))
    assert rel(tb.mul(A, dev(np.cos(0.37 * k2))).cpu().numpy(), T["rect_101x2000"]) <= TOL64


@pytest.mark.gpu
@pytest.mark.parametrize("n", [101, 2000])
def test_cuda_circulant_mathematics_match_definitions(tb, n):
    vc, vr, x, _ = recipe(n)
    C = tb.Circulant(dev(vc))
    errs = {
        "eigvals": rel(tb.eigvals(C).cpu().numpy(), T["circ_eigvals_%d" % n]),
        "ldiv": rel(tb.ldiv(C, dev(x)).cpu().numpy(), T["circ_ldiv_%d" % n]),
        "inv": rel(tb.inv(C).vc.cpu().numpy(), T["circ_inv_vc_%d" % n]),
        "strang": rel(tb.strang(tb.Toeplitz(dev(vc), dev(vr))).vc.cpu().numpy(), T["strang_%d" % n]),
        "chan": rel(tb.chan(tb.Toeplitz(dev(vc), dev(vr))).vc.cpu().numpy(), T["chan_%d" % n]),
    }
    print("circulant n=%d rel errs vs definitions: %s" % (n, {k: "%.1e" % v for k, v in errs.items()}))
    assert max(errs.values()) <= TOL64, errs
    for m in (5, 6):
        assert rel(tb.eigvals(tb.Circulant(dev(np.arange(1.0, m + 1)))).cpu().numpy(), T["circ_eigvals_small_%d" % m]) <= TOL64


@pytest.mark.gpu
def test_cuda_levinson_family_match_definitions(tb):
    b = T["kernel_b"]
    for name in ("exp", "rbf"):
        r = T["kernel_%s_r" % name]
        c = np.linalg.cond(O.SymmetricToeplitz(np.concatenate([[1.0], r[:-1]])).dense())
        bound = max(TOL64, 8 * c * EPS)
        e_d = rel(tb.durbin(dev(r)).cpu().numpy(), T["durbin_" + name])
        e_l = rel(tb.levinson(dev(r[:-1]), dev(b)).cpu().numpy(), T["levinson_" + name])
        e_t = rel(tb.trench(dev(r[:-1])).cpu().numpy(), T["trench_" + name])
        print("kernel %s: cond %.1e, durbin %.1e levinson %.1e trench %.1e (bound %.1e)" % (name, c, e_d, e_l, e_t, bound))
        assert max(e_d, e_l, e_t) <= bound
    for n in (128, 256):
        assert rel(tb.levinson(dev(T["lev_r_%d" % n]), dev(T["lev_b_%d" % n])).cpu().numpy(), T["lev_x_%d" % n]) <= TOL64
        assert rel(tb.durbin(dev(T["dur_r_%d" % n])).cpu().numpy(), T["dur_y_%d" % n]) <= TOL64


@pytest.mark.gpu
def test_cuda_iterative_solves_match_definitions(tb):
    import torch
    tol = 100 * EPS
    for n in (600, 4096):
        A = tb.SymmetricToeplitz(dev(0.5 ** np.arange(n)))
        b = dev(T["kms_b_%d" % n])
        x, err, it, flag = tb.cgs(A, torch.zeros_like(b), b, tb.factorize(tb.strang(A)), 1000, tol)
        e = rel(x.cpu().numpy(), T["kms_x_%d" % n])
        print("cgs KMS n=%d: %d iterations, flag %d, rel err vs closed form %.2e (cond 9, tol %.1e)" % (n, it, flag, e, tol))
        assert flag == 0 and e <= 9 * tol * 10
        res = tb.cg(A, torch.zeros_like(b), b, tb.factorize(tb.strang(A)), 1000, tol)
        assert res is not None and res[3] == 0 and rel(res[0].cpu().numpy(), T["kms_x_%d" % n]) <= 9 * tol * 10
    vc, vr, x101, _ = recipe(101)
    for name, A, Ao in (("toeplitz", tb.Toeplitz(dev(vc), dev(vr)), O.Toeplitz(vc, vr)),
                        ("symmetric", tb.SymmetricToeplitz(dev(vc)), O.SymmetricToeplitz(vc)),
                        ("lower", tb.LowerTriangularToeplitz(dev(vc)), O.LowerTriangularToeplitz(vc)),
                        ("upper", tb.UpperTriangularToeplitz(dev(vc)), O.UpperTriangularToeplitz(vc))):
        c = np.linalg.cond(Ao.dense())
        e = rel(tb.ldiv(A, dev(x101)).cpu().numpy(), T["solve_%s_101" % name])
        bound = max(100 * c * EPS * 100, 1e-12)
        print("ldiv %s n=101: cond %.1e, rel err vs LU(200 bit) %.2e (bound %.1e)" % (name, c, e, bound))
        assert e <= bound
