"""Regenerate tests/golden/golden_v1.npz.

The reference is a Julia package that cannot be executed in this image (no `julia`, no FFTW), so these are NOT
outputs of the reference: they are outputs of the CPU oracle (oracle/toeplitz_oracle.py, itself pinned to the
reference's own deterministic test assertions by tests/test_oracle_pinning.py) on the reference's deterministic
input recipes (test/runtests.jl:23-48,113-116,156-161,181-214,797).  Their job is to freeze the oracle: an
accidental change of the oracle or of the CUDA path shows up against a committed file.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import toeplitz_oracle as O  # noqa: E402


def build():
    g = {}
    for n in (101, 2000):
        k = np.arange(n)
        x = np.cos(0.37 * k) + 0.25 * np.sin(1.3 * k)            # deterministic input, no RNG
        xc = x + 1j * np.sin(0.11 * k)
        vc, vr = 0.9 ** k, 0.4 ** k
        g["toeplitz_real_%d" % n] = O.matvec(O.Toeplitz(vc, vr), x)
        g["toeplitz_cplx_%d" % n] = O.matvec(O.Toeplitz(vc.astype(complex), vr.astype(complex)), xc)
        g["toeplitz_adj_%d" % n] = O.matvec(O.Toeplitz(vc, vr).adjoint(), x)
        g["symmetric_%d" % n] = O.matvec(O.SymmetricToeplitz(vc), x)
        g["circulant_%d" % n] = O.matvec(O.Circulant(vc), x)
        g["upper_%d" % n] = O.matvec(O.UpperTriangularToeplitz(vc), x)
        g["lower_%d" % n] = O.matvec(O.LowerTriangularToeplitz(vc), x)
        g["hankel_%d" % n] = O.matvec(O.Hankel(0.9 ** k[::-1], vr=0.4 ** k), x)
        g["circ_ldiv_%d" % n] = O.circ_ldiv(O.Circulant(vc), x.copy())
        g["circ_eigvals_%d" % n] = O.circ_eigvals(O.Circulant(vc))
        g["circ_inv_vc_%d" % n] = O.circ_inv(O.Circulant(vc)).vc
        g["strang_%d" % n] = O.strang(O.Toeplitz(vc, vr)).vc
        g["chan_%d" % n] = O.chan(O.Toeplitz(vc, vr)).vc
    g["rect_2000x101"] = O.matvec(O.Toeplitz(0.9 ** np.arange(2000), 0.4 ** np.arange(101)), np.cos(0.37 * np.arange(101)))
    g["rect_101x2000"] = O.matvec(O.Toeplitz(0.9 ** np.arange(101), 0.4 ** np.arange(2000)), np.cos(0.37 * np.arange(2000)))
    # Levinson / Durbin kernels of test/runtests.jl:113-116 (deterministic ones)
    n = 8
    xg = np.linspace(-1, 1, n + 1)
    for name, a in (("exp", np.exp(-np.abs(xg[0] - xg))), ("rbf", np.exp(-np.abs(xg[0] - xg) ** 2) / 2)):
        a = a.copy()
        a[0] = 1
        r = a[1:]
        b = np.cos(np.arange(n) + 0.5)
        g["durbin_" + name] = O.durbin(r)
        g["levinson_" + name] = O.levinson(r[:-1], b)
        g["trench_" + name] = O.trench(r[:-1])
    # cgs + Strang on the KMS matrix: iterates after 3 steps and the converged run
    n = 600
    A = O.SymmetricToeplitz(0.5 ** np.arange(n))
    b = np.cos(0.2 * np.arange(n))
    x3, e3, it3, f3 = O.cgs(A, np.zeros(n), b.copy(), O.factorize(O.strang(A)), 3, 0.0)
    xc_, ec, itc, fc = O.cgs(A, np.zeros(n), b.copy(), O.factorize(O.strang(A)), 1000, 100 * np.finfo(float).eps)
    g["cgs3_x"] = x3
    g["cgs3_meta"] = np.array([e3, it3, f3])
    g["cgs_conv_x"] = xc_
    g["cgs_conv_meta"] = np.array([ec, itc, fc])
    g["hankel_5x5"] = O.matvec(O.Hankel(np.array([1.0, 2, 3, 4, 5]), vr=np.array([5.0, 6, 7, 8, 9])), np.ones(5))
    return g


if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **build())
    print(out, os.path.getsize(out), "bytes")
