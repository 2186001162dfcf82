"""Parity bounds shared by the GPU tests.

north_star: ||y - y_ref|| / ||y_ref|| <= 1e-12 (Float64 / ComplexF64), <= 1e-5 (Float32 / ComplexF32).  That bound is
asserted as it stands wherever the compared quantity is a product or a well-conditioned solve.  Where the quantity is
the solution of a linear system (or a determinant) both sides carry a forward error of order cond * eps, so the
assertion is `max(north_star, c * cond * eps)` with `cond` COMPUTED in the test (2-norm condition number of the dense
matrix, or max|lambda| / min|lambda| of a circulant spectrum) and a stated constant c; for iterative solves stopped at
a residual tolerance the forward bound is cond * (residual_a + residual_b).  Every check records what was measured:
`pytest -s` prints the lines, and with a `gpurun_out/` directory present they are appended to
`gpurun_out/parity_measured.jsonl` (copied to `profiles/parity_measured_r02.jsonl`).
"""
import json
import os

import numpy as np

NS_TOL = {"float64": 1e-12, "complex128": 1e-12, "float32": 1e-5, "complex64": 1e-5}
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LOG = os.path.join(_ROOT, "gpurun_out", "parity_measured.jsonl")


def eps_of(dtype):
    return float(np.finfo(np.dtype(dtype)).eps)


def ns_tol(dtype):
    return NS_TOL[np.dtype(dtype).name]


def rel(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def cond2(M):
    """2-norm condition number of a dense matrix (float64 SVD)."""
    s = np.linalg.svd(np.asarray(M).astype(np.complex128 if np.iscomplexobj(M) else np.float64), compute_uv=False)
    return float(s[0] / s[-1])


def cond_spectrum(lam):
    a = np.abs(np.asarray(lam))
    return float(a.max() / a.min())


def bound(dtype, cond=1.0, c=2.0, floor=None):
    """max(north_star tolerance, c * cond * eps(dtype))."""
    return max(ns_tol(dtype) if floor is None else floor, c * cond * eps_of(dtype))


def check(name, got, ref, dtype, cond=1.0, c=2.0, extra=0.0):
    """Assert rel(got, ref) <= max(north_star, c * cond * eps, extra); record the measurement."""
    err = rel(got, ref)
    b = max(bound(dtype, cond, c), extra)
    rec = {"check": name, "dtype": np.dtype(dtype).name, "rel_err": err, "bound": b, "north_star": ns_tol(dtype),
           "cond": cond, "c": c}
    print("parity %-58s err %.3e  bound %.3e  (cond %.3g)" % (name, err, b, cond))
    if os.path.isdir(os.path.dirname(_LOG)):
        try:
            with open(_LOG, "a") as f:
                f.write(json.dumps(rec) + "\n")
        except OSError:
            pass
    assert err <= b, rec
    return err
