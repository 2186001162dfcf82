"""ctypes binding of libtoeplitz_b200.so -- the same marshalling the Julia extension's
``ccall``s perform.  There is no fallback: if the library is missing the
import of any compute entry point raises."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TB200_LIB") or os.path.join(_HERE, "lib", "libtoeplitz_b200.so")   # TB200_LIB: experiment builds

OK, DIM_MISMATCH, BAD_ARG, UNSUPPORTED, CUDA_ERR, OOM, NCCL_ERR = range(7)
F32, F64, C32, C64 = range(4)
TOEPLITZ, SYMMETRIC, CIRCULANT, LOWER, UPPER, HANKEL = range(6)
OP_INV, OP_PINV, OP_SQRT, OP_CONJ, OP_COPY = range(5)

# every symbol include/toeplitz_b200.h declares
EXPORTS = [
    "tb200_version", "tb200_last_error", "tb200_factorize", "tb200_factorize_host", "tb200_destroy",
    "tb200_info", "tb200_mul", "tb200_mul_host", "tb200_mul_direct", "tb200_circ_ldiv",
    "tb200_circ_spectrum", "tb200_circ_map", "tb200_circ_mul", "tb200_circ_to_vc",
    "tb200_precond_vector", "tb200_levinson_batched", "tb200_durbin_batched", "tb200_cgs", "tb200_cg",
    "tb200_factorize_empty", "tb200_spectrum_buffer", "tb200_launch_count", "tb200_set_option",
    "tb200_pass_times", "tb200_pass_times_ex", "tb200_trench", "tb200_tri_smallinv", "tb200_levinson_multirhs",
    "tb200_cgs_batched", "tb200_cg_batched", "tb200_spectrum_ready",
    "tb200_comm_init", "tb200_comm_unique_id", "tb200_comm_init_rank", "tb200_comm_size", "tb200_bcast_spectrum",
    "tb200_comm_destroy", "tb200_vec_op", "tb200_vec_equal",
]


class DimensionMismatch(ValueError):
    """Julia's ``DimensionMismatch``."""


class ArgumentError(ValueError):
    """Julia's ``ArgumentError``."""


class ToeplitzB200Error(RuntimeError):
    """Julia's ``ErrorException`` (unsupported input, CUDA failure, out of memory)."""


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ToeplitzB200Error(
            "libtoeplitz_b200.so is not built (%s). Run `python toeplitzmatrices.jl_b200/build.py`; "
            "there is no CPU fallback." % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    i64, vp, ci, dbl = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    pd = ctypes.POINTER(ctypes.c_double)
    pi = ctypes.POINTER(ctypes.c_int)
    pi64 = ctypes.POINTER(ctypes.c_int64)
    pvp = ctypes.POINTER(ctypes.c_void_p)
    L.tb200_version.restype = ctypes.c_char_p
    L.tb200_last_error.restype = ctypes.c_char_p
    L.tb200_launch_count.restype = i64
    sig = {
        "tb200_factorize": [ci, ci, i64, i64, vp, vp, vp, pvp],
        "tb200_factorize_host": [ci, ci, i64, i64, vp, vp, vp, pvp],
        "tb200_factorize_empty": [ci, ci, i64, i64, vp, pvp],
        "tb200_destroy": [vp],
        "tb200_info": [vp, pi64, pi64, pi64, pi, pi],
        "tb200_mul": [vp, vp, i64, ci, vp, i64, ci, i64, pd, pd, vp],
        "tb200_mul_host": [vp, vp, i64, ci, vp, i64, ci, i64, pd, pd],
        "tb200_mul_direct": [ci, ci, i64, i64, vp, vp, vp, i64, ci, vp, i64, ci, i64, pd, pd, vp],
        "tb200_circ_ldiv": [vp, vp, i64, vp, i64, ci, i64, vp],
        "tb200_circ_spectrum": [vp, vp, vp],
        "tb200_circ_map": [vp, ci, dbl, vp, pvp],
        "tb200_circ_mul": [vp, vp, vp, pvp],
        "tb200_circ_to_vc": [vp, vp, ci, vp],
        "tb200_precond_vector": [ci, ci, ci, i64, vp, vp, vp, vp],
        "tb200_levinson_batched": [ci, i64, i64, vp, i64, vp, i64, vp, i64, vp, vp],
        "tb200_durbin_batched": [ci, i64, i64, vp, i64, vp, i64, vp],
        "tb200_cgs": [vp, vp, vp, vp, ci, i64, ci, dbl, pd, pi, pi, vp],
        "tb200_cg": [vp, vp, vp, vp, ci, i64, ci, dbl, pd, pi, pi, vp],
        "tb200_cgs_batched": [vp, vp, vp, i64, vp, i64, ci, i64, i64, ci, dbl, pd, pi, pi, vp],
        "tb200_cg_batched": [vp, vp, vp, i64, vp, i64, ci, i64, i64, ci, dbl, pd, pi, pi, vp],
        "tb200_spectrum_buffer": [vp, pvp, pi64],
        "tb200_spectrum_ready": [vp, vp],
        "tb200_comm_init": [ci, pi, pvp],
        "tb200_comm_unique_id": [vp],
        "tb200_comm_init_rank": [ci, ci, vp, pvp],
        "tb200_comm_size": [vp, pi, pi],
        "tb200_bcast_spectrum": [vp, pvp, ci, ci, pvp],
        "tb200_comm_destroy": [vp],
        "tb200_set_option": [ctypes.c_char_p, i64],
        "tb200_pass_times": [pd, pi64],
        "tb200_pass_times_ex": [pd, pi64, ctypes.c_int],
        "tb200_vec_op": [ci, i64, vp, ci, ci, vp, ci, ci, vp, ci, pd, pd, vp],
        "tb200_vec_equal": [i64, vp, ci, vp, ci, pi, vp],
        "tb200_trench": [ci, i64, vp, vp, i64, dbl, vp],
        "tb200_tri_smallinv": [ci, i64, vp, vp, vp],
        "tb200_levinson_multirhs": [ci, i64, i64, vp, vp, i64, vp, i64, vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = ci
    _lib = L
    return L


def check(status: int) -> None:
    """Map a C status to the Julia exception type the reference raises."""
    if status == OK:
        return
    msg = lib().tb200_last_error().decode("utf-8", "replace")
    if status == DIM_MISMATCH:
        raise DimensionMismatch(msg)
    if status == BAD_ARG:
        raise ArgumentError(msg)
    if status == OOM:
        raise MemoryError(msg)
    raise ToeplitzB200Error(msg)


def dbl2(z) -> ctypes.Array:
    z = complex(z)
    return (ctypes.c_double * 2)(z.real, z.imag)
