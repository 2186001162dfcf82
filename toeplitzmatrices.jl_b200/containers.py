"""Container algebra on device-resident coefficient vectors.

What callers do *between* hot-path calls -- ``tril/triu``, ``+ - scalar*``, ``conj/real/imag``, ``copyto!``,
``fill!``, ``lmul!/rmul!``, ``==``, ``transpose/adjoint`` of Hankel, ``reverse`` (Hankel <-> Toeplitz) -- is O(n)
elementwise work on the stored vectors.  The reference writes these generically (broadcasts on ``A.vc``/``A.vr``/
``A.v``: src/toeplitz.jl:66-140, src/special.jl:13-33,77-84,223-248, src/hankel.jl:75-121), so in Julia they already
run on ``CuVector`` coefficients through the array package; here device vectors go through the library's own elementwise kernel (``tb200_vec_op``: axpby with per-operand dtypes and
optional index reversal, conj, real, imag, fill; ``tb200_vec_equal`` for ``==``/``iszero``/``isdiag``) -- nothing round-trips
to the host between a ``factorize`` and the next ``mul!``; host tensors (CPU tests of the container logic) use torch.
Result TYPES follow the reference: single-vector types keep their type under unary ops, same-type ``+``/``-`` and
scalar ``*``; mixed types and ``tril``/``triu`` of Symmetric/Circulant give a general ``Toeplitz``.
"""
from __future__ import annotations

import torch

from . import api as A_
from .api import (ArgumentError, Circulant, DimensionMismatch, Hankel, LowerTriangularToeplitz, SymmetricToeplitz,
                  Toeplitz, UpperTriangularToeplitz, _SingleVector)

_TRI = (LowerTriangularToeplitz, UpperTriangularToeplitz)

# ---- elementwise work on the stored vectors: the library's own kernel (tb200_vec_op / tb200_vec_equal) for device
# ---- vectors; host tensors (CPU-side tests of the container logic) use the array package ----------------------------
_AXPBY, _CONJ, _REAL, _IMAG, _FILL = range(5)
_REAL_OF = {torch.float32: torch.float32, torch.float64: torch.float64, torch.complex64: torch.float32,
            torch.complex128: torch.float64}
_CPLX_OF = {torch.float32: torch.complex64, torch.float64: torch.complex128, torch.complex64: torch.complex64,
            torch.complex128: torch.complex128}


def _on_device(*vs) -> bool:
    return all(v is None or (v.is_cuda and v.dtype in A_._DT) for v in vs)


def _flat(v):
    v = v.resolve_conj()
    return v if (v.dim() == 1 and (v.shape[0] <= 1 or v.stride(0) == 1)) else v.contiguous()


def _vop(op, a, b=None, alpha=1.0, beta=0.0, out=None, out_dtype=None, rev_a=False, rev_b=False, n=None):
    """out[i] = op(a[i or n-1-i], b[...]) through tb200_vec_op; allocates `out` unless given (in-place is fine)."""
    ref = a if a is not None else out
    n = (a.shape[0] if a is not None else out.shape[0]) if n is None else n
    if out is None:
        out = torch.empty(n, dtype=out_dtype or ref.dtype, device=ref.device)
    a = None if a is None else _flat(a)
    b = None if b is None else _flat(b)
    A_.check(A_.L.lib().tb200_vec_op(op, n, a.data_ptr() if a is not None else None, A_._DT[a.dtype] if a is not None else 0,
                                     int(rev_a), b.data_ptr() if b is not None else None,
                                     A_._DT[b.dtype] if b is not None else 0, int(rev_b), out.data_ptr(), A_._DT[out.dtype],
                                     A_.L.dbl2(alpha), A_.L.dbl2(beta), A_._stream()))
    return out


def _scalar_dtype(v: torch.Tensor, s):
    """eltype of s * v: a complex scalar promotes a real vector (Julia's promote_type)."""
    if isinstance(s, torch.Tensor):
        s = s.item()
    return _CPLX_OF[v.dtype] if isinstance(s, complex) and not v.is_complex() else v.dtype


def _own(v: torch.Tensor) -> torch.Tensor:
    if _on_device(v):
        return _vop(_AXPBY, v)
    return v.resolve_conj().clone()


def _zero_(v: torch.Tensor) -> None:
    """v .= 0 on a (contiguous) slice."""
    if v.shape[0] == 0:
        return
    if _on_device(v) and (v.shape[0] <= 1 or v.stride(0) == 1):
        _vop(_FILL, None, out=v, alpha=0.0)
    else:
        v.zero_()


def _all_zero(v: torch.Tensor) -> bool:
    if v.shape[0] == 0:
        return True
    if _on_device(v):
        eq = A_.ctypes.c_int(0)
        vf = _flat(v)
        A_.check(A_.L.lib().tb200_vec_equal(vf.shape[0], vf.data_ptr(), A_._DT[vf.dtype], None, 0, A_.ctypes.byref(eq), A_._stream()))
        return bool(eq.value)
    return not bool(v.any())


def _vec_equal(a: torch.Tensor, b: torch.Tensor) -> bool:
    if a.shape != b.shape:
        return False
    if _on_device(a, b):
        eq = A_.ctypes.c_int(0)
        af, bf = _flat(a), _flat(b)
        A_.check(A_.L.lib().tb200_vec_equal(af.shape[0], af.data_ptr(), A_._DT[af.dtype], bf.data_ptr(), A_._DT[bf.dtype],
                                            A_.ctypes.byref(eq), A_._stream()))
        return bool(eq.value)
    return bool(torch.equal(a, b))


def _raw_toeplitz(vc: torch.Tensor, vr: torch.Tensor) -> Toeplitz:
    """Toeplitz(vc, vr) without the vc[1] == vr[1] check: in-place results of tril!/triu! with |k| > 0 legitimately
    have a zeroed corner in one vector only (the reference mutates the fields directly, src/toeplitz.jl:66-95)."""
    T = Toeplitz.__new__(Toeplitz)
    dt = torch.promote_types(vc.dtype, vr.dtype)
    T.vc, T.vr = vc.to(dt), vr.to(dt)
    return T


def copy(A):
    """``copy(A)`` (src/toeplitz.jl:108-110, src/special.jl:31-33, src/hankel.jl:88-90)."""
    return _unary(A, _own)


def _unary(A, f):
    if isinstance(A, Hankel):
        return Hankel(f(A.v), A.shape)
    if isinstance(A, _SingleVector):
        return type(A)(f(A.v))
    if isinstance(A, Toeplitz):
        return _raw_toeplitz(f(A.vc), f(A.vr))
    raise TypeError(type(A))


def zero(A):
    return _unary(A, lambda v: _vop(_FILL, None, out=torch.empty_like(v), alpha=0.0) if _on_device(v) else torch.zeros_like(v))


def conj(A):
    return _unary(A, lambda v: _vop(_CONJ, v) if _on_device(v) else v.conj().resolve_conj().clone())


def neg(A):
    return _unary(A, lambda v: _vop(_AXPBY, v, alpha=-1.0) if _on_device(v) else torch.neg(v))


def real(A):
    return _unary(A, lambda v: _vop(_REAL, v, out_dtype=_REAL_OF[v.dtype]) if _on_device(v)
                  else (v.real if v.is_complex() else v).clone())


def imag(A):
    return _unary(A, lambda v: _vop(_IMAG, v, out_dtype=_REAL_OF[v.dtype]) if _on_device(v)
                  else (v.imag.clone() if v.is_complex() else torch.zeros_like(v)))


def _binary(A, B, op):
    if isinstance(A, Hankel) or isinstance(B, Hankel):
        if not (isinstance(A, Hankel) and isinstance(B, Hankel)):
            raise TypeError("Hankel +/- Toeplitz is a dense operation in the reference")
        if A.shape != B.shape:                      # promote_shape, src/hankel.jl:93
            raise DimensionMismatch("dimensions must match: a has dims %r, b has dims %r" % (A.shape, B.shape))
        if A.v.shape != B.v.shape:                  # the broadcast on A.v, B.v fails in the reference too
            raise DimensionMismatch("arrays could not be broadcast to a common size: %d and %d"
                                    % (A.v.shape[0], B.v.shape[0]))
        return Hankel(op(A.v, B.v), A.shape)
    if A.shape != B.shape:
        raise DimensionMismatch("dimensions must match: a has dims %r, b has dims %r" % (A.shape, B.shape))
    if type(A) is type(B) and isinstance(A, _SingleVector):      # src/special.jl:77-79
        return type(A)(op(A.v, B.v))
    return _raw_toeplitz(op(A.vc, B.vc), op(A.vr, B.vr))        # src/toeplitz.jl:111-113


def _axpby(a, b, beta):
    if _on_device(a, b) and a.shape == b.shape:
        return _vop(_AXPBY, a, b, alpha=1.0, beta=beta, out_dtype=torch.promote_types(a.dtype, b.dtype))
    return a + beta * b if beta != 1.0 else a + b


def add(A, B):
    return _binary(A, B, lambda a, b: _axpby(a, b, 1.0))


def sub(A, B):
    return _binary(A, B, lambda a, b: _axpby(a, b, -1.0))


def _pyscalar(s):
    return s.item() if isinstance(s, torch.Tensor) else s


def scale(A, s):
    """``s * A`` / ``A * s`` (src/toeplitz.jl:127-128, src/special.jl:24-25, src/hankel.jl:117-118)."""
    return _unary(A, lambda v: _vop(_AXPBY, v, alpha=_pyscalar(s), out_dtype=_scalar_dtype(v, s)) if _on_device(v) else v * s)


def lmul_(s, A):
    """``lmul!(s, A)`` (src/toeplitz.jl:130-135, src/special.jl:13-16, src/hankel.jl:99-104)."""
    for v in _fields(A):
        if _on_device(v) and (v.shape[0] <= 1 or v.stride(0) == 1) and (v.is_complex() or not isinstance(_pyscalar(s), complex)):
            _vop(_AXPBY, v, alpha=_pyscalar(s), out=v)
        else:
            v.mul_(s)
    return A


def rmul_(A, s):
    """``rmul!(A, s)`` (src/toeplitz.jl:136-141, src/special.jl:17-20, src/hankel.jl:105-110)."""
    return lmul_(s, A)


def fill_(A, s):
    """``fill!(A, s)`` (src/toeplitz.jl:122-126, src/hankel.jl:105-110)."""
    if isinstance(A, _SingleVector):
        raise TypeError("fill! is defined for Toeplitz and Hankel only")
    for v in _fields(A):
        if _on_device(v) and (v.shape[0] <= 1 or v.stride(0) == 1) and (v.is_complex() or not isinstance(_pyscalar(s), complex)):
            _vop(_FILL, None, out=v, alpha=_pyscalar(s))
        else:
            v.fill_(s)
    return A


def copyto_(A, B):
    """``copyto!(A, B)`` (src/toeplitz.jl:114-119, src/special.jl:71-74, src/hankel.jl:94-98)."""
    if isinstance(A, Hankel):
        if not isinstance(B, Hankel) or A.shape != B.shape:
            raise DimensionMismatch("dimensions must match")
        _copy_checked(A.v, B.v)
        return A
    if isinstance(A, _SingleVector):
        if type(A) is not type(B):
            raise TypeError("copyto! between different single-vector types")
        _copy_checked(A.v, B.v)
        return A
    _copy_checked(A.vc, B.vc)
    _copy_checked(A.vr, B.vr)
    return A


def _copy_checked(dst, src):
    if dst.shape[0] < src.shape[0]:
        raise DimensionMismatch("destination has fewer elements than required")     # Base.copyto! BoundsError analogue
    d = dst[: src.shape[0]]
    if _on_device(d, src) and (d.shape[0] <= 1 or d.stride(0) == 1) and (d.is_complex() or not src.is_complex()):
        _vop(_AXPBY, src, out=d)
    else:
        d.copy_(src)


def _fields(A):
    if isinstance(A, Toeplitz):
        return (A.vc, A.vr)
    return (A.v,)


def equal(A, B) -> bool:
    """``A == B`` (src/toeplitz.jl:120, src/special.jl:67, src/hankel.jl:116): one device reduction, one bool back."""
    if isinstance(A, Hankel) or isinstance(B, Hankel):
        return (isinstance(A, Hankel) and isinstance(B, Hankel) and A.shape == B.shape and A.v.shape == B.v.shape
                and _vec_equal(A.v, B.v))
    if A.shape != B.shape:
        return False
    if type(A) is type(B) and isinstance(A, _SingleVector):
        return _vec_equal(A.v, B.v)
    return _vec_equal(A.vr, B.vr) and _vec_equal(A.vc, B.vc)


def iszero(A) -> bool:
    return all(_all_zero(v) for v in _fields(A))


def isdiag(A) -> bool:
    """src/special.jl:245-246."""
    if isinstance(A, _SingleVector):
        return _all_zero(A.v[1:])
    return _all_zero(A.vc[1:]) and _all_zero(A.vr[1:])


# ---- tril / triu ---------------------------------------------------------------------------------------------
def tril_(A, k: int = 0):
    """``tril!(A, k)``: Toeplitz src/toeplitz.jl:66-80; triangular src/special.jl:223-241."""
    if isinstance(A, Toeplitz):
        if k >= 0:
            _zero_(A.vr[k + 1:])
        else:
            _zero_(A.vr)
            _zero_(A.vc[: min(-k, A.vc.shape[0])])
        return A
    if isinstance(A, UpperTriangularToeplitz):      # _tridiff!(A, k)
        return _tridiff_(A, k)
    if isinstance(A, LowerTriangularToeplitz):      # _trisame!(A, k)
        return _trisame_(A, k)
    raise TypeError("tril! is defined for Toeplitz and triangular Toeplitz")


def triu_(A, k: int = 0):
    """``triu!(A, k)``: Toeplitz src/toeplitz.jl:81-95; triangular src/special.jl:223-241."""
    if isinstance(A, Toeplitz):
        if k <= 0:
            _zero_(A.vc[-k + 1:])
        else:
            _zero_(A.vc)
            _zero_(A.vr[: min(k, A.vr.shape[0])])
        return A
    if isinstance(A, LowerTriangularToeplitz):
        return _tridiff_(A, -k)
    if isinstance(A, UpperTriangularToeplitz):
        return _trisame_(A, -k)
    raise TypeError("triu! is defined for Toeplitz and triangular Toeplitz")


def _tridiff_(A, k):
    if k >= 0:
        _zero_(A.v[k + 1:])
    else:
        _zero_(A.v)
    return A


def _trisame_(A, k):
    if k < 0:
        _zero_(A.v[: min(-k, A.v.shape[0])])
    return A


def tril(A, k: int = 0):
    """``tril(A, k)`` (src/toeplitz.jl:97, src/special.jl:242): triangular types keep their type, the others
    become a general Toeplitz."""
    if isinstance(A, _TRI):
        return tril_(copy(A), k)
    return tril_(_raw_toeplitz(_own(A.vc), _own(A.vr)), k)


def triu(A, k: int = 0):
    if isinstance(A, _TRI):
        return triu_(copy(A), k)
    return triu_(_raw_toeplitz(_own(A.vc), _own(A.vr)), k)


# ---- Hankel <-> Toeplitz -------------------------------------------------------------------------------------
def hankel_vc(H: Hankel) -> torch.Tensor:           # src/hankel.jl:26-36
    return H.v[: H.shape[0]]


def hankel_vr(H: Hankel) -> torch.Tensor:
    return H.v[H.shape[0] - 1:]


def reverse(A, dims: int = 1):
    """``reverse(A; dims)`` (src/hankel.jl:112-130).  The vectors are built exactly as the reference builds them,
    including ``reverse(::AbstractToeplitz, dims=2) = Hankel(vcat(reverse(vr), vc), size)`` whose corner entry
    appears twice (the reference only type-checks this method, test/runtests.jl:227-229,379-381)."""
    if dims not in (1, 2):
        raise ArgumentError("invalid dimension %d in reverse" % dims)
    if isinstance(A, Hankel):
        n = A.shape[1]
        if dims == 1:
            return _raw_toeplitz(_rev(hankel_vc(A)), _own(hankel_vr(A)))
        return _raw_toeplitz(_own(A.v[n - 1:]), _rev(A.v[:n]))
    if dims == 1:
        return Hankel(_rev(A.vc), vr=A.vr)
    nr, nc = A.vr.shape[0], A.vc.shape[0]
    dt = torch.promote_types(A.vr.dtype, A.vc.dtype)
    if _on_device(A.vr, A.vc):
        v = torch.empty(nr + nc, dtype=dt, device=A.vr.device)
        _vop(_AXPBY, A.vr, out=v[:nr], rev_a=True)
        _vop(_AXPBY, A.vc, out=v[nr:])
        return Hankel(v, A.shape)
    return Hankel(torch.cat([A.vr.flip(0), A.vc]), A.shape)


def _rev(v: torch.Tensor) -> torch.Tensor:
    """reverse(v) (index reversal fused into the copy kernel)."""
    return _vop(_AXPBY, v, rev_a=True) if _on_device(v) else v.flip(0)


transpose = A_.transpose
adjoint = A_.adjoint


def getindex(A, i: int, j: int):
    """1-based ``A[i, j]`` (src/toeplitz.jl:42-51, src/special.jl:133-150, src/hankel.jl:83-86); returns a 0-dim
    device tensor (no synchronisation)."""
    m, n = A.shape
    if not (1 <= i <= m and 1 <= j <= n):
        raise IndexError("BoundsError: attempt to access %dx%d matrix at index [%d, %d]" % (m, n, i, j))
    if isinstance(A, Hankel):
        return A.v[i + j - 2]
    d = i - j
    if isinstance(A, SymmetricToeplitz):
        return A.v[abs(d)]
    if isinstance(A, Circulant):
        return A.v[d + m if d < 0 else d]
    if isinstance(A, LowerTriangularToeplitz):
        return A.v[d] if d >= 0 else torch.zeros((), dtype=A.dtype, device=A.device)
    if isinstance(A, UpperTriangularToeplitz):
        return A.v[-d] if d <= 0 else torch.zeros((), dtype=A.dtype, device=A.device)
    return A.vc[d] if d >= 0 else A.vr[-d]


def dense(A) -> torch.Tensor:
    """``Matrix(A)`` on the device (a gather; for tests and small debugging only)."""
    m, n = A.shape
    i = torch.arange(m, device=A.device)[:, None]
    j = torch.arange(n, device=A.device)[None, :]
    if isinstance(A, Hankel):
        return A.v[i + j]
    d = i - j
    vc, vr = A.vc, A.vr
    return torch.where(d >= 0, vc[d.clamp(min=0)], vr[(-d).clamp(min=0)])


def _install_operators():
    """Python operator sugar for the algebra above (``A + B``, ``-A``, ``2 * A``, ``A == B``)."""
    for cls in (A_.AbstractToeplitz, Hankel):
        cls.__add__ = lambda self, other: add(self, other)
        cls.__sub__ = lambda self, other: sub(self, other)
        cls.__neg__ = lambda self: neg(self)
        cls.__mul__ = lambda self, s: (A_.mul(self, s) if isinstance(s, torch.Tensor) and s.dim() >= 1 else scale(self, s))
        cls.__rmul__ = lambda self, s: scale(self, s)
        cls.__eq__ = lambda self, other: equal(self, other) if isinstance(other, (A_.AbstractToeplitz, Hankel)) else NotImplemented
        cls.__hash__ = lambda self: id(self)
        cls.conj = lambda self: conj(self)
        cls.copy = lambda self: copy(self)


_install_operators()
