"""Regenerate tests/golden/truth_v1.npz: DEFINITION-level answers in 60-digit arithmetic (mpmath).

The reference (a Julia package) cannot run in this image and holds no golden vectors, so nothing here is an output of
the reference.  Unlike golden_v1.npz (frozen outputs of our own oracle) these values do not come from ANY fast
algorithm: every entry is the mathematical definition evaluated with mpmath at 200 bits -- dense products
sum_j A[i,j] x[j], dense solves Matrix(A) \\ b by LU, the DFT as sum_j v[j] exp(-2 pi i jk/n), closed-form inverses
checked by a 200-bit residual -- on the reference's deterministic input recipes (test/runtests.jl:23-48, 76-77,
113-116, 156-161, 181-214, 797) plus the BASELINE recipes of SURVEY 8d.  Inputs that the reference draws from its
RNG are replaced by closed-form sequences (stored in the file).  The oracle AND the CUDA path are asserted against
this file (tests/test_truth.py), which makes the parity chain  definition -> oracle -> CUDA  non self-referential.

    python tests/golden/make_truth.py        # ~5 minutes, pure Python
"""
import os
import sys
import time

import mpmath as mp
import numpy as np

mp.mp.prec = 200
HERE = os.path.dirname(os.path.abspath(__file__))


def M(a):
    """numpy real array -> list of mpf (exact: binary floats convert exactly)."""
    return [mp.mpf(float(v)) for v in np.asarray(a, dtype=np.float64).ravel()]


def toeplitz_row(vc, vr, i, n):
    """Row i of Toeplitz(vc, vr): A[i,j] = vc[i-j] (i >= j) else vr[j-i]   (src/toeplitz.jl:47-55)."""
    lo = vc[i::-1][: n] if i < n else vc[i:i - n:-1]
    row = lo + vr[1:n - i] if i < n else lo
    return row[:n]


def dense_matvec(rowfun, m, x):
    return np.array([float(mp.fdot(rowfun(i), x)) for i in range(m)])


def dense_matvec_c(rowfun, m, xre, xim):
    re = np.array([float(mp.fdot(rowfun(i), xre)) for i in range(m)])
    im = np.array([float(mp.fdot(rowfun(i), xim)) for i in range(m)])
    return re + 1j * im


def dft_mp(vre, vim, n, sign=-1):
    """X[k] = sum_j v[j] exp(sign 2 pi i jk/n), exact twiddles from an n-entry table."""
    tw = [mp.expjpi(mp.mpf(sign * 2 * j) / n) for j in range(n)]
    twr, twi = [t.real for t in tw], [t.imag for t in tw]
    out_r, out_i = [], []
    for k in range(n):
        idx = [(j * k) % n for j in range(n)]
        wr = [twr[t] for t in idx]
        wi = [twi[t] for t in idx]
        re = mp.fdot(vre, wr) - (mp.fdot(vim, wi) if vim else 0)
        im = mp.fdot(vre, wi) + (mp.fdot(vim, wr) if vim else 0)
        out_r.append(re)
        out_i.append(im)
    return out_r, out_i


def lu_solve_mp(rows, b):
    A = mp.matrix(rows)
    return mp.lu_solve(A, mp.matrix(b))


def tofloat(v):
    return np.array([float(t) for t in v])


def build():
    g = {}
    t0 = time.time()
    # ---- products on the 0.9^k / 0.4^k recipes (test/runtests.jl:23-48), ns = 101, nl = 2000 -----------------
    for n in (101, 2000):
        k = np.arange(n)
        x = np.cos(0.37 * k) + 0.25 * np.sin(1.3 * k)
        xi = np.sin(0.11 * k)
        vc_f, vr_f = 0.9 ** k, 0.4 ** k
        g["x_%d" % n], g["xi_%d" % n] = x, xi
        vc, vr, X, XI = M(vc_f), M(vr_f), M(x), M(xi)
        zero = [mp.mpf(0)] * n
        rows = {
            "toeplitz": lambda i: toeplitz_row(vc, vr, i, n),
            "toeplitz_adj": lambda i: toeplitz_row(vr, vc, i, n),
            "symmetric": lambda i: toeplitz_row(vc, vc, i, n),
            "circulant": lambda i: toeplitz_row(vc, [vc[0]] + vc[:0:-1], i, n),
            "lower": lambda i: toeplitz_row(vc, [vc[0]] + zero[1:], i, n),
            "upper": lambda i: toeplitz_row([vc[0]] + zero[1:], vc, i, n),
        }
        for name, rf in rows.items():
            g["%s_%d" % (name, n)] = dense_matvec(rf, n, X)
        g["toeplitz_cplx_%d" % n] = dense_matvec_c(rows["toeplitz"], n, X, XI)
        # Hankel(0.9^(n-1:-1:0), 0.4^(0:n-1)) (test/runtests.jl:181-214): H[i,j] = v[i+j], v = [vc; vr[2:end]]
        hv = M(np.concatenate([(0.9 ** k)[::-1], (0.4 ** k)[1:]]))
        g["hankel_%d" % n] = dense_matvec(lambda i: hv[i:i + n], n, X)
        print("products n=%d done, %.0f s" % (n, time.time() - t0), flush=True)
        # ---- Circulant mathematics: eigvals = DFT(vc) (src/linearalgebra.jl:184-192,286-287), inv, ldiv ------
        er, ei = dft_mp(vc, None, n)
        g["circ_eigvals_%d" % n] = tofloat(er) + 1j * tofloat(ei)
        # inv(C).vc = IDFT(1 ./ lambda), C \\ x = IDFT(DFT(x) ./ lambda) -- spectral definition, then verified below
        lam = [mp.mpc(a, b) for a, b in zip(er, ei)]
        inv_l = [1 / t for t in lam]
        ir, ii = dft_mp([t.real for t in inv_l], [t.imag for t in inv_l], n, sign=+1)
        inv_vc = [t / n for t in ir]
        g["circ_inv_vc_%d" % n] = tofloat(inv_vc)
        xr_, xi_ = dft_mp(X, None, n)
        q = [mp.mpc(a, b) / l for a, b, l in zip(xr_, xi_, lam)]
        sr, _ = dft_mp([t.real for t in q], [t.imag for t in q], n, sign=+1)
        sol = [t / n for t in sr]
        # residual of the DEFINITION C*sol = x at 200 bits
        res = max(abs(mp.fdot(rows["circulant"](i), sol) - X[i]) for i in range(0, n, max(1, n // 50)))
        assert res < mp.mpf(10) ** -45, res
        g["circ_ldiv_%d" % n] = tofloat(sol)
        print("circulant n=%d done, %.0f s" % (n, time.time() - t0), flush=True)
        # strang / chan (src/linearalgebra.jl:255-274), exact rational arithmetic
        st = [vc[i] if i <= n // 2 else vr[n - i] for i in range(n)]
        ch = [((n - i) * vc[i] + i * vr[min(n - i, n - 1)]) / n for i in range(n)]
        g["strang_%d" % n] = tofloat(st)
        g["chan_%d" % n] = tofloat(ch)
    # rectangular (test/runtests.jl:83-95)
    k2, k1 = np.arange(2000), np.arange(101)
    vcl, vrs = M(0.9 ** k2), M(0.4 ** k1)
    xs = M(np.cos(0.37 * k1))
    g["rect_2000x101"] = dense_matvec(lambda i: (vcl[i::-1] + vrs[1:101 - i])[:101] if i < 101 else vcl[i:i - 101:-1], 2000, xs)
    vcs, vrl = M(0.9 ** k1), M(0.4 ** k2)
    xl = M(np.cos(0.37 * k2))
    g["rect_101x2000"] = dense_matvec(lambda i: vcs[i::-1] + vrl[1:2000 - i], 101, xl)
    # small circulant sizes of the reference's tests (5, 6: test/runtests.jl:502-506,797)
    for n in (5, 6):
        v = M(np.arange(1, n + 1, dtype=float))
        er, ei = dft_mp(v, None, n)
        g["circ_eigvals_small_%d" % n] = tofloat(er) + 1j * tofloat(ei)
    # ---- dense solves Matrix(A) \\ b by LU (test/runtests.jl:56-63) at ns = 101 ------------------------------
    n = 101
    k = np.arange(n)
    vc, vr, X = M(0.9 ** k), M(0.4 ** k), M(g["x_101"])
    zero = [mp.mpf(0)] * n
    for name, (c_, r_) in {"toeplitz": (vc, vr), "symmetric": (vc, vc), "lower": (vc, [vc[0]] + zero[1:]),
                           "upper": ([vc[0]] + zero[1:], vc)}.items():
        rows = [toeplitz_row(c_, r_, i, n) for i in range(n)]
        g["solve_%s_101" % name] = tofloat(lu_solve_mp(rows, X))
    print("dense solves done, %.0f s" % (time.time() - t0), flush=True)
    # ---- Durbin / Levinson / Trench on the deterministic kernels of test/runtests.jl:112-151 -------------------
    n = 8
    xg = np.linspace(-1, 1, n + 1)
    for name, a in (("exp", np.exp(-np.abs(xg[0] - xg))), ("rbf", np.exp(-np.abs(xg[0] - xg) ** 2) / 2)):
        a = a.copy()
        a[0] = 1
        r = a[1:]                                          # n entries
        b = np.cos(np.arange(n) + 0.5)
        g["kernel_%s_r" % name], g["kernel_b"] = r, b
        col = M(np.concatenate([[1.0], r[:-1]]))           # T = SymmetricToeplitz([1; r[1:end-1]])
        rows = [toeplitz_row(col, col, i, n) for i in range(n)]
        g["durbin_" + name] = -tofloat(lu_solve_mp(rows, M(r)))
        g["levinson_" + name] = tofloat(lu_solve_mp(rows, M(b)))
        inv = mp.matrix(rows) ** -1
        g["trench_" + name] = np.array([[float(inv[i, j]) for j in range(n)] for i in range(n)])
    # ---- BASELINE config 4 recipe: r = u/n diagonally dominant, n = 128, 256; KMS rho^k -----------
    for n in (128, 256):
        u = (np.modf(np.sqrt(2.0) * (np.arange(1, n) ** 1.5))[0]) / n      # closed-form stand-in for U(0,1)/n
        b = np.cos(0.7 * np.arange(n)) + 0.1
        g["lev_r_%d" % n], g["lev_b_%d" % n] = u, b
        col = M(np.concatenate([[1.0], u]))
        rows = [toeplitz_row(col, col, i, n) for i in range(n)]
        g["lev_x_%d" % n] = tofloat(lu_solve_mp(rows, M(b)))
        rd = np.concatenate([u, [0.3 / n]])                                # durbin takes n entries
        g["dur_r_%d" % n] = rd
        g["dur_y_%d" % n] = -tofloat(lu_solve_mp(rows, M(rd)))
        print("levinson n=%d done, %.0f s" % (n, time.time() - t0), flush=True)
    # ---- KMS systems (config 5b recipe, SymmetricToeplitz(0.5^k)): closed-form inverse, residual-verified ------
    for n in (600, 4096):
        rho = mp.mpf("0.5")
        b = np.cos(0.2 * np.arange(n))
        B = M(b)
        d = 1 - rho * rho
        x = []
        for i in range(n):
            diag = 1 if i in (0, n - 1) else 1 + rho * rho
            acc = diag * B[i]
            if i > 0:
                acc -= rho * B[i - 1]
            if i < n - 1:
                acc -= rho * B[i + 1]
            x.append(acc / d)
        col = [rho ** i for i in range(n)]
        res = max(abs(mp.fdot(toeplitz_row(col, col, i, n), x) - B[i]) for i in range(0, n, max(1, n // 40)))
        assert res < mp.mpf(10) ** -45, res
        g["kms_b_%d" % n], g["kms_x_%d" % n] = b, tofloat(x)
    # ---- exact small cases (test/runtests.jl:76-77,156-161) ----------------------------------------------------
    g["exact_toeplitz_2"] = np.array([3.0, 3.0])
    g["exact_circulant_3"] = np.array([6.0, 6.0, 6.0])
    hv = M(np.array([1.0, 2, 3, 4, 5, 6, 7, 8, 9]))
    g["hankel_5x5"] = dense_matvec(lambda i: hv[i:i + 5], 5, M(np.ones(5)))
    print("total %.0f s" % (time.time() - t0), flush=True)
    return g


if __name__ == "__main__":
    out = os.path.join(HERE, "truth_v1.npz")
    np.savez_compressed(out, **build())
    print(out, os.path.getsize(out), "bytes")
