"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the oracle on
the reference's deterministic recipes): the oracle must still reproduce them (CPU), and the CUDA path must match
them through the C ABI (GPU)."""
import os

import numpy as np
import pytest

import parity_tools as PT

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v1.npz"))


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def test_oracle_reproduces_golden():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fresh = mod.build()
    assert sorted(fresh) == sorted(G.files)
    for k in G.files:
        assert rel(fresh[k], G[k]) <= 1e-13, k
    assert np.array_equal(G["hankel_5x5"], np.array([15.0, 20, 25, 30, 35]))


@pytest.mark.gpu
def test_cuda_matches_golden():
    import torch
    import toeplitz_b200 as tb
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    host = lambda t: t.cpu().numpy()
    for n in (101, 2000):
        k = np.arange(n)
        x = np.cos(0.37 * k) + 0.25 * np.sin(1.3 * k)
        xc = x + 1j * np.sin(0.11 * k)
        vc, vr = 0.9 ** k, 0.4 ** k
        T = tb.Toeplitz(dev(vc), dev(vr))
        cases = {
            "toeplitz_real_%d" % n: tb.mul(T, dev(x)),
            "toeplitz_cplx_%d" % n: tb.mul(tb.Toeplitz(dev(vc.astype(complex)), dev(vr.astype(complex))), dev(xc)),
            "toeplitz_adj_%d" % n: tb.mul(tb.adjoint(T), dev(x)),
            "symmetric_%d" % n: tb.mul(tb.SymmetricToeplitz(dev(vc)), dev(x)),
            "circulant_%d" % n: tb.mul(tb.Circulant(dev(vc)), dev(x)),
            "upper_%d" % n: tb.mul(tb.UpperTriangularToeplitz(dev(vc)), dev(x)),
            "lower_%d" % n: tb.mul(tb.LowerTriangularToeplitz(dev(vc)), dev(x)),
            "hankel_%d" % n: tb.mul(tb.Hankel(dev(0.9 ** k[::-1].copy()), vr=dev(vr)), dev(x)),
            "circ_ldiv_%d" % n: tb.ldiv(tb.Circulant(dev(vc)), dev(x)),
            "circ_eigvals_%d" % n: tb.eigvals(tb.Circulant(dev(vc))),
            "circ_inv_vc_%d" % n: tb.inv(tb.Circulant(dev(vc))).vc,
            "strang_%d" % n: tb.strang(T).vc,
            "chan_%d" % n: tb.chan(T).vc,
        }
        kC = PT.cond_spectrum(np.fft.fft(vc))            # the three Circulant solves divide by the spectrum
        for name, val in cases.items():
            PT.check("golden " + name, host(val), G[name], np.float64, kC if name.startswith(("circ_ldiv", "circ_inv")) else 1.0)
    assert rel(host(tb.mul(tb.Toeplitz(dev(0.9 ** np.arange(2000)), dev(0.4 ** np.arange(101))), dev(np.cos(0.37 * np.arange(101))))), G["rect_2000x101"]) <= 1e-12
    assert rel(host(tb.mul(tb.Toeplitz(dev(0.9 ** np.arange(101)), dev(0.4 ** np.arange(2000))), dev(np.cos(0.37 * np.arange(2000))))), G["rect_101x2000"]) <= 1e-12
    n = 8
    xg = np.linspace(-1, 1, n + 1)
    for name, a in (("exp", np.exp(-np.abs(xg[0] - xg))), ("rbf", np.exp(-np.abs(xg[0] - xg) ** 2) / 2)):
        a = a.copy()
        a[0] = 1
        r = a[1:]
        b = np.cos(np.arange(n) + 0.5)
        from scipy.linalg import toeplitz as _toep
        kT = PT.cond2(_toep(np.concatenate([[1.0], r[:-1]])))          # exp / rbf kernels on 9 points: cond up to ~1e3
        PT.check("golden durbin_" + name, host(tb.durbin(dev(r))), G["durbin_" + name], np.float64, kT)
        PT.check("golden levinson_" + name, host(tb.levinson(dev(r[:-1].copy()), dev(b))), G["levinson_" + name], np.float64, kT)
        PT.check("golden trench_" + name, host(tb.trench(dev(r[:-1].copy()))), G["trench_" + name], np.float64, kT)
    n = 600
    A = tb.SymmetricToeplitz(dev(0.5 ** np.arange(n)))
    b = dev(np.cos(0.2 * np.arange(n)))
    M = tb.factorize(tb.strang(A))
    x3, e3, it3, f3 = tb.cgs(A, torch.zeros_like(b), b, M, 3, 0.0)
    assert (it3, f3) == (int(G["cgs3_meta"][1]), int(G["cgs3_meta"][2]))
    from scipy.linalg import toeplitz as _toep
    Dn = _toep(0.5 ** np.arange(n))
    kA, bh = PT.cond2(Dn), host(b)
    floor = (np.linalg.norm(bh - Dn @ host(x3)) + np.linalg.norm(bh - Dn @ G["cgs3_x"])) / np.linalg.norm(bh)
    PT.check("golden cgs 3 iterations", host(x3), G["cgs3_x"], np.float64, kA, extra=2.0 * kA * floor)
    xc_, ec, itc, fc = tb.cgs(A, torch.zeros_like(b), b, M, 1000, 100 * np.finfo(float).eps)
    assert fc == int(G["cgs_conv_meta"][2]) and abs(itc - int(G["cgs_conv_meta"][1])) <= 1
    floor = (np.linalg.norm(bh - Dn @ host(xc_)) + np.linalg.norm(bh - Dn @ G["cgs_conv_x"])) / np.linalg.norm(bh)
    PT.check("golden cgs to tolerance", host(xc_), G["cgs_conv_x"], np.float64, kA, extra=2.0 * kA * floor)
    H = tb.Hankel(dev(np.array([1.0, 2, 3, 4, 5])), vr=dev(np.array([5.0, 6, 7, 8, 9])))
    assert np.allclose(host(tb.mul(H, dev(np.ones(5)))), G["hankel_5x5"], rtol=0, atol=1e-13)
