"""Build / load the C half of the oracle (TEST INFRASTRUCTURE).

``python oracle/build_oracle.py`` compiles ``levinson_oracle.c`` into
``oracle/_build/liboracle_levinson.so`` (git-ignored, travels to the GPU box
with the gpurun snapshot).  The Julia reference has no C sources, so there is
no ``oracle/_ref`` build: the reference is "unbuildable here" (no ``julia``,
no FFTW) and every CPU baseline is ``kind: "port"``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "levinson_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_OUT = os.path.join(_OUT_DIR, "liboracle_levinson.so")


def build(force: bool = False) -> str:
    os.makedirs(_OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(_OUT)
            and os.path.getmtime(_OUT) >= os.path.getmtime(_SRC)):
        return _OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _OUT, _SRC]
    subprocess.run(cmd, check=True)
    return _OUT


_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = _OUT if os.path.exists(_OUT) else build()
        lib = ctypes.CDLL(path)
        i64, dp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        lib.oracle_durbin.argtypes = [dp, dp, i64]
        lib.oracle_levinson.argtypes = [dp, dp, dp, dp, i64]
        lib.oracle_levinson_sym.argtypes = [dp, dp, dp, i64]
        lib.oracle_levinson_batched.argtypes = [dp, i64, dp, i64, dp, i64, i64, i64, ci]
        lib.oracle_durbin_batched.argtypes = [dp, i64, dp, i64, i64, i64, ci]
        lib.oracle_max_threads.restype = ci
        for f in ("oracle_durbin", "oracle_levinson", "oracle_levinson_sym",
                  "oracle_levinson_batched", "oracle_durbin_batched"):
            getattr(lib, f).restype = None
        _lib = lib
    return _lib


# numpy-facing helpers ------------------------------------------------------
def _p(a):
    return a.ctypes.data


def c_durbin(r):
    import numpy as np
    r = np.ascontiguousarray(r, dtype=np.float64)
    y = np.zeros_like(r)
    load().oracle_durbin(_p(y), _p(r), len(r))
    return y


def c_levinson(r, b):
    import numpy as np
    r = np.ascontiguousarray(r, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = len(b)
    x = np.zeros(n)
    y = np.zeros(n)
    load().oracle_levinson(_p(x), _p(r), _p(b), _p(y), n)
    return x


def c_levinson_sym(vc, b):
    import numpy as np
    vc = np.ascontiguousarray(vc, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(len(b))
    load().oracle_levinson_sym(_p(x), _p(vc), _p(b), len(b))
    return x


def c_levinson_batched(R, B, nthreads=0):
    """R: (n-1, nsys), B: (n, nsys), both Fortran-ordered float64 -> X (n, nsys)."""
    import numpy as np
    R = np.asfortranarray(R, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    n, nsys = B.shape
    X = np.zeros((n, nsys), order="F")
    load().oracle_levinson_batched(_p(X), n, _p(R), R.shape[0], _p(B), n, n, nsys, nthreads)
    return X


def c_durbin_batched(R, nthreads=0):
    import numpy as np
    R = np.asfortranarray(R, dtype=np.float64)
    n, nsys = R.shape
    Y = np.zeros((n, nsys), order="F")
    load().oracle_durbin_batched(_p(Y), n, _p(R), n, n, nsys, nthreads)
    return Y


def max_threads() -> int:
    return int(load().oracle_max_threads())


if __name__ == "__main__":
    print(build(force=True))
