"""Host-side mirror of the reference's operator interface for the hot path.

Same names, argument meaning and error behaviour as ToeplitzMatrices.jl (v0.8.5), with
Julia's ``f!`` spelled ``f_``: ``factorize``, ``mul_`` (``mul!``), ``mul`` (``*``),
``ldiv_`` (``ldiv!``), ``ldiv`` (``\\``), ``eigvals``, ``det``, ``inv``, ``pinv``,
``sqrt``, ``adjoint``, ``strang``, ``chan``, ``levinson``, ``levinson_``, ``durbin``,
``durbin_``, ``cgs``, ``cg``.  Vectors and matrices are CUDA ``torch`` tensors (device
memory + streams only -- every flop happens in libtoeplitz_b200.so through the C ABI of
``include/toeplitz_b200.h``).  Matrices follow Julia's column-major convention: a
``(n, l)`` tensor with ``stride(0) == 1`` (use :func:`colmajor`).

Reference citations are relative to /root/reference.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib as L
from ._lib import ArgumentError, DimensionMismatch, ToeplitzB200Error, check

_DT = {torch.float32: L.F32, torch.float64: L.F64, torch.complex64: L.C32, torch.complex128: L.C64}
_REAL_OF = {torch.complex64: torch.float32, torch.complex128: torch.float64,
            torch.float32: torch.float32, torch.float64: torch.float64}
_CPLX_OF = {torch.float32: torch.complex64, torch.float64: torch.complex128,
            torch.complex64: torch.complex64, torch.complex128: torch.complex128}


# ------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------
def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _floatify(v: torch.Tensor) -> torch.Tensor:
    """``float(T)``: integer/bool element types become Float64 (src/linearalgebra.jl:125)."""
    if v.dtype in _DT:
        return v
    return v.to(torch.float64)


def _as_vec(v, device=None) -> torch.Tensor:
    if not isinstance(v, torch.Tensor):
        v = torch.as_tensor(v)
    if device is not None and v.device != torch.device(device):
        v = v.to(device)
    if not v.is_cuda:
        v = v.cuda()
    if v.dim() != 1:
        raise ArgumentError("expected a vector")
    return _floatify(v).contiguous()


def colmajor(n: int, l: int, dtype=torch.float64, device="cuda") -> torch.Tensor:
    """An uninitialised Julia-style ``Matrix`` (n rows, l columns, column-major)."""
    return torch.empty((l, n), dtype=dtype, device=device).t()


def to_colmajor(X: torch.Tensor) -> torch.Tensor:
    """Dense Julia layout: unit stride down a column (any StridedVector / StridedMatrix is accepted by the
    reference, src/linearalgebra.jl:5,43; the kernels address columns densely, so strided views are packed)."""
    if X.dim() == 1:
        return X if (X.stride(0) == 1 or X.shape[0] <= 1) else X.contiguous()
    if X.stride(0) == 1 and X.stride(1) >= X.shape[0]:
        return X
    out = colmajor(X.shape[0], X.shape[1], X.dtype, X.device)
    out.copy_(X)
    return out


def _is_colmajor(X: torch.Tensor) -> bool:
    if X.dim() == 1:
        return X.stride(0) == 1 or X.shape[0] <= 1
    return (X.stride(0) == 1 or X.shape[0] <= 1) and (X.stride(1) >= X.shape[0] or X.shape[1] <= 1)


def _ld(X: torch.Tensor) -> int:
    if X.dim() == 1:
        return max(X.shape[0], 1)
    return max(X.stride(1), X.shape[0], 1) if X.shape[1] > 1 else max(X.shape[0], 1)


def _ncols(X: torch.Tensor) -> int:
    return 1 if X.dim() == 1 else X.shape[1]


def _is_real(dt) -> bool:
    return dt in (torch.float32, torch.float64)


def _prec64(dt) -> bool:
    return dt in (torch.float64, torch.complex128)


def _with_prec(dt, like):
    """dtype ``dt`` moved to the precision of ``like`` keeping real/complex."""
    if _prec64(like):
        return torch.float64 if _is_real(dt) else torch.complex128
    return torch.float32 if _is_real(dt) else torch.complex64


# ------------------------------------------------------------------------------------
# matrix types  (src/toeplitz.jl:7-18, src/special.jl:44-117, src/hankel.jl:2-36)
# ------------------------------------------------------------------------------------
class AbstractToeplitz:
    kind = -1

    @property
    def dtype(self):
        return self.vc.dtype

    @property
    def device(self):
        return self.vc.device

    @property
    def shape(self):
        return (self.vc.shape[0], self.vr.shape[0])

    def size(self, i=None):
        return self.shape if i is None else self.shape[i - 1]

    def __matmul__(self, x):
        return mul(self, x)

    def adjoint(self):
        return adjoint(self)


class Toeplitz(AbstractToeplitz):
    kind = L.TOEPLITZ

    def __init__(self, vc, vr):
        vc = _as_vec(vc)
        vr = _as_vec(vr, vc.device)
        T = torch.promote_types(vc.dtype, vr.dtype)
        self.vc, self.vr = vc.to(T), vr.to(T)
        if bool(self.vc[0] != self.vr[0]):        # src/toeplitz.jl:13-15
            raise ToeplitzB200Error("First element of the vectors must be the same")


class _SingleVector(AbstractToeplitz):
    def __init__(self, v):
        self.v = _as_vec(v)

    @property
    def dtype(self):
        return self.v.dtype

    @property
    def device(self):
        return self.v.device

    @property
    def shape(self):
        return (self.v.shape[0], self.v.shape[0])


class SymmetricToeplitz(_SingleVector):
    kind = L.SYMMETRIC
    vc = property(lambda self: self.v)
    vr = property(lambda self: self.v)


class Circulant(_SingleVector):
    kind = L.CIRCULANT
    vc = property(lambda self: self.v)

    @property
    def vr(self):                                   # src/special.jl:119  [v1, vn, ..., v2]
        return torch.cat([self.v[:1], self.v[1:].flip(0)])


class LowerTriangularToeplitz(_SingleVector):
    kind = L.LOWER
    uplo = "L"
    vc = property(lambda self: self.v)

    @property
    def vr(self):                                   # src/special.jl:120-124
        w = torch.zeros_like(self.v)
        w[0] = self.v[0]
        return w


class UpperTriangularToeplitz(_SingleVector):
    kind = L.UPPER
    uplo = "U"
    vr = property(lambda self: self.v)

    @property
    def vc(self):
        w = torch.zeros_like(self.v)
        w[0] = self.v[0]
        return w


class Hankel:
    """``H[i,j] = v[i+j-1]`` (src/hankel.jl:2-12)."""
    kind = L.HANKEL

    def __init__(self, v, size=None, vr=None):
        if vr is not None:                          # Hankel(vc, vr)  src/hankel.jl:20-25
            vc = _as_vec(v)
            vr = _as_vec(vr, vc.device)
            if bool(vc[-1] != vr[0]):
                raise ArgumentError("vc[end] != vr[1]")
            T = torch.promote_types(vc.dtype, vr.dtype)
            v = torch.cat([vc.to(T), vr[1:].to(T)])
            size = (vc.shape[0], vr.shape[0])
        self.v = _as_vec(v)
        if size is None:
            l = self.v.shape[0]
            size = ((l + 1) // 2, (l + 1) // 2)
        m, n = size
        if m < 0 or n < 0:
            raise ArgumentError("negative size: %r" % (size,))
        if self.v.shape[0] < m + n - 1:             # src/hankel.jl:9
            raise ArgumentError("inconsistency between size and number of anti-diagonals")
        self.shape = (m, n)

    dtype = property(lambda self: self.v.dtype)
    device = property(lambda self: self.v.device)

    def size(self, i=None):
        return self.shape if i is None else self.shape[i - 1]

    def __matmul__(self, x):
        return mul(self, x)


def adjoint(A):
    """Eager adjoints (src/toeplitz.jl:100-101, src/special.jl:127-130, src/linearalgebra.jl:214-215)."""
    if isinstance(A, ToeplitzFactorization):
        return A.adjoint()
    if isinstance(A, Toeplitz):
        return Toeplitz(A.vr.conj().resolve_conj(), A.vc.conj().resolve_conj())
    if isinstance(A, SymmetricToeplitz):
        return SymmetricToeplitz(A.v.conj().resolve_conj())
    if isinstance(A, Circulant):
        return Circulant(A.vr.conj().resolve_conj())
    if isinstance(A, LowerTriangularToeplitz):
        return UpperTriangularToeplitz(A.v.conj().resolve_conj())
    if isinstance(A, UpperTriangularToeplitz):
        return LowerTriangularToeplitz(A.v.conj().resolve_conj())
    if isinstance(A, Hankel):                       # src/hankel.jl:113
        return Hankel(A.v.conj().resolve_conj(), (A.shape[1], A.shape[0]))
    raise TypeError(type(A))


def transpose(A):
    if isinstance(A, Toeplitz):
        return Toeplitz(A.vr, A.vc)
    if isinstance(A, SymmetricToeplitz):
        return A
    if isinstance(A, Circulant):
        return Circulant(A.vr)
    if isinstance(A, LowerTriangularToeplitz):
        return UpperTriangularToeplitz(A.v)
    if isinstance(A, UpperTriangularToeplitz):
        return LowerTriangularToeplitz(A.v)
    if isinstance(A, Hankel):                       # src/hankel.jl:112
        return Hankel(A.v, (A.shape[1], A.shape[0]))
    raise TypeError(type(A))


# ------------------------------------------------------------------------------------
# factorization  (src/ToeplitzMatrices.jl:97-101)
# ------------------------------------------------------------------------------------
class ToeplitzFactorization:
    """Opaque device factorization (spectrum of the circulant embedding + plan)."""

    def __init__(self, handle: int, kind: int, dtype, shape, keep=()):
        self._h = ctypes.c_void_p(handle)
        self.kind = kind
        self.dtype = dtype          # eltype T of the matrix
        self.shape = shape
        self._keep = keep
        self.device_index = torch.cuda.current_device() if torch.cuda.is_available() else 0

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                L.lib().tb200_destroy(h)
            except Exception:
                pass
            self._h = None

    def size(self, i=None):         # src/special.jl:262-263
        return self.shape if i is None else self.shape[i - 1]

    @property
    def is_circulant(self):
        return self.kind == L.CIRCULANT

    def _derived(self, handle):
        return ToeplitzFactorization(handle, self.kind, self.dtype, self.shape)

    def adjoint(self):
        """``adjoint(C::CirculantFactorization)`` -- conj spectrum (src/linearalgebra.jl:214-215)."""
        return _circ_map(self, L.OP_CONJ)

    def __matmul__(self, x):
        if isinstance(x, (ToeplitzFactorization, Circulant)):
            return circ_mul(self, x)
        return mul(self, x)


def factorize(A) -> ToeplitzFactorization:
    """``factorize`` for the five layouts + Hankel (src/linearalgebra.jl:122-131,140-150,184-192,415-435)."""
    if isinstance(A, ToeplitzFactorization):
        return A
    out = ctypes.c_void_p()
    m, n = A.shape
    dt = _DT[A.dtype]
    if isinstance(A, Toeplitz):
        vc, vr = A.vc, A.vr
        check(L.lib().tb200_factorize(L.TOEPLITZ, dt, m, n, vc.data_ptr(), vr.data_ptr(), _stream(), ctypes.byref(out)))
    elif isinstance(A, Hankel):
        if A.v.shape[0] != m + n - 1:
            # reverse(A, dims=2) keeps the surplus tail (src/hankel.jl:117) and mul! then fails its size check
            raise DimensionMismatch(
                "first dimension of A, %d, does not match length of y, %d" % (A.v.shape[0] - n + 1, m))
        check(L.lib().tb200_factorize(L.HANKEL, dt, m, n, A.v.data_ptr(), None, _stream(), ctypes.byref(out)))
    else:
        check(L.lib().tb200_factorize(A.kind, dt, n, n, A.v.data_ptr(), None, _stream(), ctypes.byref(out)))
    return ToeplitzFactorization(out.value, A.kind, A.dtype, (m, n))


_ONE2, _ZERO2 = L.dbl2(1.0), L.dbl2(0.0)


def _mul_factorization_for(A) -> ToeplitzFactorization:
    """Factorization used for products.  A Circulant whose size has no exact device DFT plan
    is multiplied through its Toeplitz embedding (same product)."""
    if isinstance(A, ToeplitzFactorization):
        return A
    if isinstance(A, Circulant):
        n = A.shape[0]
        if n < 16 or (n & (n - 1)) != 0:
            return factorize(Toeplitz(A.vc, A.vr))
    return factorize(A)


# ------------------------------------------------------------------------------------
# mul!  (src/linearalgebra.jl:4-99, src/hankel.jl:133-137)
# ------------------------------------------------------------------------------------
def _scalar_is_real(s) -> bool:
    if isinstance(s, (bool, int, float)):
        return True
    if isinstance(s, torch.Tensor):
        return not s.is_complex()
    return complex(s).imag == 0 and not isinstance(s, complex)


def _promote_eltype(y, A, x, alpha, beta):
    """``Base.promote_eltype(y, A, x, alpha, beta)`` is real iff everything is (:55)."""
    real = _is_real(y.dtype) and _is_real(A.dtype) and _is_real(x.dtype)
    return real and _scalar_is_real(alpha) and _scalar_is_real(beta)


def _check_mul_dims(y, shape, x):
    m, n = shape
    if y.shape[0] != m:
        raise DimensionMismatch("first dimension of A, %d, does not match length of y, %d" % (m, y.shape[0]))
    if x.shape[0] != n:
        raise DimensionMismatch("second dimension of A, %d, does not match length of x, %d" % (n, x.shape[0]))


def mul_(y: torch.Tensor, A, x: torch.Tensor, alpha=True, beta=False) -> torch.Tensor:
    """``mul!(y, A, x, alpha, beta)``: ``y = alpha*A*x + beta*y`` for vectors and matrices."""
    # the plain call of an iterative solver -- dense vectors of the factorization's own element type, alpha = 1, beta = 0 --
    # goes straight to the C ABI: every check below holds by construction (a few microseconds of Python per product
    # otherwise, which is what bounds single-vector products up to n ~ 2^17)
    if (alpha is True and beta is False and type(A) is ToeplitzFactorization and x.dim() == 1 and y.dim() == 1
            and x.dtype == A.dtype and y.dtype == A.dtype and x.is_cuda and y.is_cuda
            and y.shape[0] == A.shape[0] and x.shape[0] == A.shape[1] and x.stride(0) == 1 and y.stride(0) == 1):
        dt = _DT[A.dtype]
        check(L.lib().tb200_mul(A._h, y.data_ptr(), A.shape[0], dt, x.data_ptr(), A.shape[1], dt, 1, _ONE2, _ZERO2, _stream()))
        return y
    if y.dim() != x.dim():
        raise DimensionMismatch("y and x must both be vectors or both be matrices")
    if x.dim() == 2 and y.shape[1] != x.shape[1]:
        raise DimensionMismatch("input and output matrices must have same number of columns")
    is_fact = isinstance(A, ToeplitzFactorization)
    if is_fact:
        # src/linearalgebra.jl:49-53
        m, n = A.shape
        if y.shape[0] != m or x.shape[0] != n:
            raise DimensionMismatch("Toeplitz factorization does not match size of input and output vector")
    else:
        if isinstance(A, Hankel) and A.v.shape[0] != A.shape[0] + A.shape[1] - 1:
            # reverse(A, dims=2) keeps the surplus tail of v (src/hankel.jl:117), so the Toeplitz it builds is
            # taller than y and the reference's mul! fails its first size check (src/linearalgebra.jl:8-12)
            raise DimensionMismatch("first dimension of A, %d, does not match length of y, %d"
                                    % (A.v.shape[0] - A.shape[1] + 1, y.shape[0]))
        _check_mul_dims(y, A.shape, x)
    if not _promote_eltype(y, A, x, alpha, beta) and _is_real(y.dtype):
        raise ArgumentError("InexactError: cannot store a complex product into a real y")
    Adt = A.dtype
    work_x = _with_prec(x.dtype, Adt)
    work_y = _with_prec(y.dtype, Adt)
    xin = x if x.dtype == work_x else x.to(work_x)
    if not _is_colmajor(xin):
        xin = to_colmajor(xin)
    direct_y = (y.dtype == work_y) and _is_colmajor(y)
    m, n = A.shape
    # small case: vectors only, N = m+n-1 < 512 (src/linearalgebra.jl:19-34)
    small = (not is_fact) and x.dim() == 1 and (m + n - 1) < 512
    if direct_y:
        yout, a, b = y, alpha, beta
    else:
        yout = colmajor(y.shape[0], _ncols(y), work_y, y.device) if y.dim() == 2 else torch.empty(y.shape[0], dtype=work_y, device=y.device)
        a, b = True, False
    aa, bb = L.dbl2(a), L.dbl2(b)
    if small:
        if isinstance(A, Hankel):
            vcp, vrp = A.v.data_ptr(), None
        elif isinstance(A, Toeplitz):
            vcp, vrp = A.vc.data_ptr(), A.vr.data_ptr()
        else:
            vcp, vrp = A.v.data_ptr(), None
        check(L.lib().tb200_mul_direct(A.kind, _DT[Adt], m, n, vcp, vrp, yout.data_ptr(), _ld(yout), _DT[yout.dtype],
                                       xin.data_ptr(), _ld(xin), _DT[xin.dtype], 1, aa, bb, _stream()))
    else:
        F = _mul_factorization_for(A)
        check(L.lib().tb200_mul(F._h, yout.data_ptr(), _ld(yout), _DT[yout.dtype], xin.data_ptr(), _ld(xin),
                                _DT[xin.dtype], _ncols(xin), aa, bb, _stream()))
    if not direct_y:
        if beta == 0:
            y.copy_(alpha * yout if alpha is not True else yout)
        else:
            y.copy_(alpha * yout.to(y.dtype) + beta * y)
    return y


def mul(A, x: torch.Tensor) -> torch.Tensor:
    """``A * x`` (vector or matrix)."""
    if isinstance(x, (ToeplitzFactorization, Circulant)) and isinstance(A, (ToeplitzFactorization, Circulant)):
        return circ_mul(A, x)
    x = _floatify(x)
    T = torch.promote_types(A.dtype, x.dtype)
    m = A.shape[0]
    y = torch.empty(m, dtype=T, device=x.device) if x.dim() == 1 else colmajor(m, x.shape[1], T, x.device)
    return mul_(y, A, x, True, False)


# ------------------------------------------------------------------------------------
# Circulant spectral family  (src/linearalgebra.jl:183-315)
# ------------------------------------------------------------------------------------
def _circ_fact(C) -> ToeplitzFactorization:
    if isinstance(C, ToeplitzFactorization):
        if not C.is_circulant:
            raise ArgumentError("expected a Circulant factorization")
        return C
    if not isinstance(C, Circulant):
        raise ArgumentError("expected a Circulant")
    return factorize(C)


def _circ_map(F: ToeplitzFactorization, op: int, tol: float = 0.0) -> ToeplitzFactorization:
    out = ctypes.c_void_p()
    check(L.lib().tb200_circ_map(F._h, op, float(tol), _stream(), ctypes.byref(out)))
    return F._derived(out.value)


def _to_circulant(F: ToeplitzFactorization, T) -> "Circulant":
    """``Circulant(maybereal(T, dft \\ spectrum))``."""
    n = F.shape[0]
    vc = torch.empty(n, dtype=T, device="cuda")
    check(L.lib().tb200_circ_to_vc(F._h, vc.data_ptr(), _DT[T], _stream()))
    return Circulant(vc)


def circ_ldiv_(F: ToeplitzFactorization, b: torch.Tensor, out: torch.Tensor = None) -> torch.Tensor:
    """``ldiv!(C::CirculantFactorization, b)`` (:225-242); ``out`` gives the 3-arg form."""
    n = b.shape[0]
    if F.shape[0] != n:
        raise DimensionMismatch("size of Toeplitz factorization does not match the length of the output vector")
    x = b if out is None else out
    if x.shape != b.shape:
        raise DimensionMismatch("x and b must have the same size")
    work = _with_prec(b.dtype, F.dtype)
    bin_ = b if b.dtype == work else b.to(work)
    if not _is_colmajor(bin_):
        bin_ = to_colmajor(bin_)
    direct = (x.dtype == work) and _is_colmajor(x)
    xo = x if direct else (bin_ if bin_ is not b else torch.empty_like(bin_))
    if _is_real(work) and not _is_real(F.dtype):
        raise ArgumentError("InexactError: real b with a complex Circulant")
    check(L.lib().tb200_circ_ldiv(F._h, xo.data_ptr(), _ld(xo), bin_.data_ptr(), _ld(bin_), _DT[work], _ncols(bin_), _stream()))
    if _is_real(F.dtype) and not _is_real(work):
        # `b .= maybereal.(T, tmp)` keeps only the real part for a real Circulant (:239-240)
        xo.imag.zero_()
    if not direct:
        x.copy_(xo)
    return x


def ldiv_(A, b: torch.Tensor) -> torch.Tensor:
    """``ldiv!(A, b)`` for vectors and (column loop, :102-110) matrices."""
    if isinstance(A, ToeplitzFactorization):
        if not A.is_circulant:
            raise ArgumentError("ldiv! is only defined for Circulant factorizations")
        return circ_ldiv_(A, b)
    m, n = A.shape
    if b.dim() == 2 and m != n:
        raise ToeplitzB200Error("Division: Rectangular case is not supported.")
    if isinstance(A, Circulant):
        return circ_ldiv_(factorize(A), b)              # :224
    tol = 100 * torch.finfo(torch.float64).eps
    FA = factorize(A)          # cached spectrum instead of the reference's per-product re-factorization
    if isinstance(A, Toeplitz):                         # :133-136  cgs + Strang
        pre = factorize(strang(A))
        solver = lambda col: cgs(FA, torch.zeros_like(col), col, pre, 1000, tol)[0]
    elif isinstance(A, SymmetricToeplitz):              # :152      cg + Strang, Float64 start vector
        if b.dtype != torch.float64:
            raise ArgumentError("MethodError: cg needs x, b of the same BlasReal type (zeros(length(b)) is Float64)")
        pre = factorize(strang(A))

        def solver(col):
            res = cg(FA, torch.zeros_like(col), col, pre, 1000, tol)
            if res is None:
                raise ToeplitzB200Error("MethodError: cg returned nothing (initial residual below tol, "
                                        "src/iterativeLinearSolvers.jl:57-59)")
            return res[0]
    elif isinstance(A, (LowerTriangularToeplitz, UpperTriangularToeplitz)):   # :399-402 cgs + T. Chan
        pre = factorize(chan(A))
        solver = lambda col: cgs(FA, torch.zeros_like(col), col, pre, 1000, tol)[0]
    else:
        raise TypeError(type(A))
    if b.dim() == 1:
        b.copy_(solver(b.contiguous()))
        return b
    # matrix right-hand side (:102-110): the reference loops over the columns; here they iterate together
    Bc = b if _is_colmajor(b) else to_colmajor(b)
    Xs = colmajor(Bc.shape[0], Bc.shape[1], Bc.dtype, Bc.device)
    Xs.zero_()
    if isinstance(A, SymmetricToeplitz):
        _, _, _, flags = cg_batched(FA, Xs, Bc, pre, 1000, tol)
        if any(f == 2 for f in flags):
            raise ToeplitzB200Error("MethodError: cg returned nothing (initial residual below tol, "
                                    "src/iterativeLinearSolvers.jl:57-59)")
    else:
        cgs_batched(FA, Xs, Bc, pre, 1000, tol)
    b.copy_(Xs)
    return b


def ldiv(A, b: torch.Tensor) -> torch.Tensor:
    """``A \\ b`` (:112-120): refuses eltype promotion of the matrix."""
    T = torch.promote_types(A.dtype, b.dtype)
    if T != A.dtype:
        raise ArgumentError("promotion of Toeplitz matrices not handled yet")
    bb = b.to(T).clone() if b.dim() == 1 else to_colmajor(b.to(T)).clone()
    return ldiv_(A, bb)


def eigvals(C) -> torch.Tensor:
    """``eigvals(C)`` (:286-287): the spectrum in DFT order, always complex."""
    F = _circ_fact(C)
    out = torch.empty(F.shape[0], dtype=_CPLX_OF[F.dtype], device="cuda")
    check(L.lib().tb200_circ_spectrum(F._h, out.data_ptr(), _stream()))
    return out


def det(C):
    """``det`` (:288-290)."""
    d = torch.prod(eigvals(C))
    return d.real if _is_real(C.dtype) else d


def _tri_inv_lower(v: torch.Tensor) -> torch.Tensor:
    """First column of ``inv(LowerTriangularToeplitz(v))`` by recursive doubling (:337-365); every product
    is a device ``mul`` (direct kernel below N = 512, FFT above), the n <= 64 base case one small kernel."""
    n = v.shape[0]
    if n <= 64:
        out = torch.empty_like(v)
        check(L.lib().tb200_tri_smallinv(_DT[v.dtype], n, v.data_ptr(), out.data_ptr(), _stream()))
        return out
    np2 = 1 << (n - 1).bit_length()
    if n != np2:
        vp = torch.zeros(np2, dtype=v.dtype, device=v.device)
        vp[:n] = v
        return _tri_inv_lower(vp)[:n].contiguous()
    nd2 = n // 2
    a1 = _tri_inv_lower(v[:nd2].contiguous())
    T = Toeplitz(v[nd2:].contiguous(), v[1:nd2 + 1].flip(0).contiguous())
    t = mul(T, a1)
    return torch.cat([a1, -mul(LowerTriangularToeplitz(a1), t)])


def inv(C):
    """``inv(C::Circulant)`` -> Circulant (:244-249); ``inv(F)`` -> factorization (:250-253);
    ``inv(::Lower/UpperTriangularToeplitz)`` -> triangular Toeplitz (:337-396).  The upper inverse uses the
    coefficients of the lower one (inv(U(v)) = U(inv(L(v)).v)); the reference's own upper branch is
    ``@test_broken`` (test/runtests.jl:660), so it is matched against the dense inverse instead."""
    if isinstance(C, LowerTriangularToeplitz):
        return LowerTriangularToeplitz(_tri_inv_lower(C.v))
    if isinstance(C, UpperTriangularToeplitz):
        return UpperTriangularToeplitz(_tri_inv_lower(C.v))
    if isinstance(C, ToeplitzFactorization):
        return _circ_map(_circ_fact(C), L.OP_INV)
    return _to_circulant(_circ_map(_circ_fact(C), L.OP_INV), C.dtype)


def pinv(C: Circulant, tol=None):
    """``pinv(C, tol)`` (:276-284)."""
    if tol is None:
        tol = torch.finfo(_REAL_OF[C.dtype]).eps
    return _to_circulant(_circ_map(_circ_fact(C), L.OP_PINV, tol), C.dtype)


def sqrt(C):
    """``sqrt(C)`` (:311-315)."""
    F = _circ_fact(C)
    return _to_circulant(_circ_map(F, L.OP_SQRT), F.dtype)


def circ_mul(A, B) -> Circulant:
    """``A * B`` for Circulants / factorizations (:194-211, 216-222)."""
    FA, FB = _circ_fact(A), _circ_fact(B)
    if FA.shape[0] != FB.shape[0]:
        m, n = FA.shape[0], FB.shape[0]
        raise DimensionMismatch("size of matrix A, %dx%d, does not match size of matrix B, %dx%d" % (m, m, n, n))
    out = ctypes.c_void_p()
    check(L.lib().tb200_circ_mul(FA._h, FB._h, _stream(), ctypes.byref(out)))
    T = torch.promote_types(FA.dtype, FB.dtype)
    F = ToeplitzFactorization(out.value, L.CIRCULANT, T, FA.shape)
    return _to_circulant(F, T)


def _precond(which: int, A) -> Circulant:
    m, n = A.shape
    if isinstance(A, Hankel):
        raise ArgumentError("strang/chan need a Toeplitz-structured matrix")
    v = torch.empty(n, dtype=A.dtype, device=A.device)
    if isinstance(A, Toeplitz):
        vcp, vrp = A.vc.data_ptr(), A.vr.data_ptr()
    else:
        vcp, vrp = A.v.data_ptr(), None
    check(L.lib().tb200_precond_vector(which, A.kind, _DT[A.dtype], n, vcp, vrp, v.data_ptr(), _stream()))
    return Circulant(v)


def strang(A) -> Circulant:
    """``strang(A)`` (:255-266)."""
    return _precond(0, A)


def chan(A) -> Circulant:
    """``chan(A)`` (:267-274)."""
    return _precond(1, A)


# ------------------------------------------------------------------------------------
# direct solvers  (src/directLinearSolvers.jl)
# ------------------------------------------------------------------------------------
def durbin_(y: torch.Tensor, r: torch.Tensor) -> torch.Tensor:
    """``durbin!(y, r)`` (:15-28); matrices = one system per column (batched)."""
    n = r.shape[0]
    if y.shape[0] != n:
        raise DimensionMismatch("length(y) = %d != %d = length(r)" % (y.shape[0], n))
    if _ncols(y) != _ncols(r):
        raise DimensionMismatch("y and r must have the same number of columns")
    if not (_is_colmajor(y) and _is_colmajor(r)) or y.dtype != r.dtype:
        raise ArgumentError("durbin!: y and r must be column-major and of the same real dtype")
    check(L.lib().tb200_durbin_batched(_DT[r.dtype], n, _ncols(r), r.data_ptr(), _ld(r), y.data_ptr(), _ld(y), _stream()))
    return y


def durbin(r: torch.Tensor) -> torch.Tensor:
    """``durbin(r)`` (:8)."""
    r = to_colmajor(_floatify(r))
    return durbin_(torch.zeros_like(r) if r.dim() == 1 else colmajor(r.shape[0], r.shape[1], r.dtype, r.device), r)


def levinson_(x: torch.Tensor, r: torch.Tensor, b: torch.Tensor, diag: torch.Tensor = None) -> torch.Tensor:
    """``levinson!(x, r, b)`` (:100-121); matrices = one system per column (batched)."""
    n = r.shape[0] + 1
    if x.shape[0] != n:
        raise DimensionMismatch("length(x) = %d != %d = length(r) + 1" % (x.shape[0], n))
    if b.shape[0] != n:
        raise DimensionMismatch("length(b) = %d != %d = length(r) + 1" % (b.shape[0], n))
    if not (_ncols(x) == _ncols(b) == _ncols(r)):
        raise DimensionMismatch("x, r and b must have the same number of columns")
    if not (_is_colmajor(x) and _is_colmajor(r) and _is_colmajor(b)) or not (x.dtype == r.dtype == b.dtype):
        raise ArgumentError("levinson!: x, r, b must be column-major and of the same real dtype")
    dp = None if diag is None else diag.data_ptr()
    check(L.lib().tb200_levinson_batched(_DT[r.dtype], n, _ncols(b), r.data_ptr(), _ld(r), b.data_ptr(), _ld(b),
                                         x.data_ptr(), _ld(x), dp, _stream()))
    return x


def levinson(r_or_T, b: torch.Tensor) -> torch.Tensor:
    """``levinson(r, b)`` (:93) and ``levinson(T::SymmetricToeplitz, b)`` (:123-134; many RHS
    columns as in ext/ToeplitzMatricesStatsBaseExt.jl:15-24)."""
    b = to_colmajor(_floatify(b))
    x = torch.zeros_like(b) if b.dim() == 1 else colmajor(b.shape[0], b.shape[1], b.dtype, b.device)
    if isinstance(r_or_T, SymmetricToeplitz) and b.dim() == 2 and 2 <= b.shape[0] <= 512 and b.shape[1] >= 2:
        # many right-hand sides, one matrix: the y-recursion is shared (ext/ToeplitzMatricesStatsBaseExt.jl:15-24)
        vc = r_or_T.vc.to(b.dtype).contiguous()
        if vc.shape[0] != b.shape[0]:
            raise DimensionMismatch("length(b) = %d != %d = size(T, 1)" % (b.shape[0], vc.shape[0]))
        check(L.lib().tb200_levinson_multirhs(_DT[b.dtype], b.shape[0], b.shape[1], vc.data_ptr(), b.data_ptr(), _ld(b),
                                              x.data_ptr(), _ld(x), _stream()))
        return x
    if isinstance(r_or_T, SymmetricToeplitz):
        T = r_or_T
        l = _ncols(b)
        r = T.vc[1:].to(b.dtype)
        diag = T.vc[:1].to(b.dtype)
        if b.dim() == 2:
            r = r.unsqueeze(1).expand(-1, l)
            r = to_colmajor(r.clone() if l == 1 else r)
            diag = diag.expand(l).contiguous()
        else:
            r = r.contiguous()
        return levinson_(x, r, b, diag)
    r = to_colmajor(_floatify(r_or_T).to(b.dtype))
    return levinson_(x, r, b)


def trench(r_or_T) -> torch.Tensor:
    """``trench(T::SymmetricToeplitz)`` / ``trench(r)`` (:36-85): dense inverse of the SPD matrix, O(n^2)
    (Durbin + the anti-diagonal recurrence).  Returns the full symmetric ``n x n`` matrix (column-major)."""
    if isinstance(r_or_T, SymmetricToeplitz):
        vc = r_or_T.vc
        r0 = float(vc[0].item())
        r = (vc[1:] / r0).contiguous() if r0 != 1.0 else vc[1:].contiguous()
    else:
        r = _floatify(r_or_T).contiguous()
        r0 = 1.0
    if not _is_real(r.dtype):
        raise ArgumentError("trench needs a real symmetric positive definite Toeplitz matrix")
    n = r.shape[0] + 1
    B = colmajor(n, n, r.dtype, r.device)
    check(L.lib().tb200_trench(_DT[r.dtype], n, r.data_ptr() if n > 1 else None, B.data_ptr(), _ld(B), r0, _stream()))
    return B


# ------------------------------------------------------------------------------------
# iterative solvers  (src/iterativeLinearSolvers.jl)
# ------------------------------------------------------------------------------------
def _krylov(fn, A, x, b, M, max_it, tol):
    FA = A if isinstance(A, ToeplitzFactorization) else factorize(A)
    if M is None:
        Mh = None
    else:
        FM = _circ_fact(M)
        Mh = FM._h
    if x.dtype != b.dtype:
        raise ArgumentError("MethodError: x and b must have the same element type")
    if x.dim() != 1 or b.dim() != 1 or not x.is_contiguous() or not b.is_contiguous():
        raise ArgumentError("x and b must be contiguous vectors")
    if _with_prec(b.dtype, FA.dtype) != b.dtype:
        raise ArgumentError("vector precision must match the matrix precision")
    err = ctypes.c_double()
    it = ctypes.c_int()
    flag = ctypes.c_int()
    check(fn(FA._h, Mh, x.data_ptr(), b.data_ptr(), _DT[b.dtype], b.shape[0], int(max_it), float(tol),
             ctypes.byref(err), ctypes.byref(it), ctypes.byref(flag), _stream()))
    return x, err.value, it.value, flag.value


def cgs(A, x, b, M, max_it, tol):
    """``IterativeLinearSolvers.cgs(A, x, b, M, max_it, tol)`` -> ``(x, error, iter, flag)`` (:99-215)."""
    return _krylov(L.lib().tb200_cgs, A, x, b, M, max_it, tol)


def cg(A, x, b, M, max_it, tol):
    """``IterativeLinearSolvers.cg`` (:9-97).  Returns ``None`` where the reference's early
    ``return`` yields ``nothing`` (:57-59)."""
    if not _is_real(b.dtype):
        raise ArgumentError("MethodError: cg is defined for BlasReal element types")
    res = _krylov(L.lib().tb200_cg, A, x, b, M, max_it, tol)
    if res[3] == 2:
        return None
    return res


def _krylov_batched(fn, A, X, B, M, max_it, tol):
    FA = A if isinstance(A, ToeplitzFactorization) else factorize(A)
    Mh = None if M is None else _circ_fact(M)._h
    if X.dtype != B.dtype:
        raise ArgumentError("MethodError: X and B must have the same element type")
    if X.dim() != 2 or B.dim() != 2 or X.shape != B.shape:
        raise DimensionMismatch("X and B must be matrices of the same size")
    if not _is_colmajor(X) or not _is_colmajor(B):
        raise ArgumentError("X and B must be column-major (tb.colmajor / tb.to_colmajor)")
    if _with_prec(B.dtype, FA.dtype) != B.dtype:
        raise ArgumentError("vector precision must match the matrix precision")
    l = B.shape[1]
    err = (ctypes.c_double * max(l, 1))()
    it = (ctypes.c_int * max(l, 1))()
    flag = (ctypes.c_int * max(l, 1))()
    check(fn(FA._h, Mh, X.data_ptr(), _ld(X), B.data_ptr(), _ld(B), _DT[B.dtype], B.shape[0], l, int(max_it), float(tol),
             err, it, flag, _stream()))
    return X, list(err)[:l], list(it)[:l], list(flag)[:l]


def cgs_batched(A, X, B, M, max_it, tol):
    """All columns of ``B`` through ``cgs`` at once (the column loop of ``ldiv!(A, B::Matrix)``, src/linearalgebra.jl:
    102-110): ``(X, errors, iters, flags)`` with one entry per column."""
    return _krylov_batched(L.lib().tb200_cgs_batched, A, X, B, M, max_it, tol)


def cg_batched(A, X, B, M, max_it, tol):
    """All columns of ``B`` through ``cg`` at once; ``flags[j] == 2`` marks the reference's value-less early return."""
    if not _is_real(B.dtype):
        raise ArgumentError("MethodError: cg is defined for BlasReal element types")
    return _krylov_batched(L.lib().tb200_cg_batched, A, X, B, M, max_it, tol)


# ------------------------------------------------------------------------------------
# host-buffer entry point (the end-to-end path bench.py times) and multi-GPU plumbing
# ------------------------------------------------------------------------------------
def mul_host_(Y: torch.Tensor, F: ToeplitzFactorization, X: torch.Tensor, alpha=True, beta=False) -> torch.Tensor:
    """``mul!`` on HOST (ideally pinned) column-major matrices through ``tb200_mul_host``."""
    if X.is_cuda or Y.is_cuda:
        raise ArgumentError("mul_host_ takes host tensors")
    if not (_is_colmajor(X) and _is_colmajor(Y)):
        raise ArgumentError("host matrices must be column-major")
    m, n = F.shape
    if Y.shape[0] != m or X.shape[0] != n:
        raise DimensionMismatch("Toeplitz factorization does not match size of input and output vector")
    if _ncols(X) != _ncols(Y):
        raise DimensionMismatch("input and output matrices must have same number of columns")
    check(L.lib().tb200_mul_host(F._h, Y.data_ptr(), _ld(Y), _DT[Y.dtype], X.data_ptr(), _ld(X), _DT[X.dtype],
                                 _ncols(X), L.dbl2(alpha), L.dbl2(beta)))
    return Y


def factorize_host(kind_cls, vc: torch.Tensor, vr: torch.Tensor = None, size=None) -> ToeplitzFactorization:
    """``factorize`` from HOST coefficient vectors (``tb200_factorize_host``)."""
    out = ctypes.c_void_p()
    vc = _floatify(vc).contiguous()
    dt = _DT[vc.dtype]
    if kind_cls is Toeplitz:
        vr = vr.to(vc.dtype).contiguous()
        m, n = vc.shape[0], vr.shape[0]
        check(L.lib().tb200_factorize_host(L.TOEPLITZ, dt, m, n, vc.data_ptr(), vr.data_ptr(), _stream(), ctypes.byref(out)))
    elif kind_cls is Hankel:
        m, n = size
        check(L.lib().tb200_factorize_host(L.HANKEL, dt, m, n, vc.data_ptr(), None, _stream(), ctypes.byref(out)))
    else:
        m = n = vc.shape[0]
        check(L.lib().tb200_factorize_host(kind_cls.kind, dt, n, n, vc.data_ptr(), None, _stream(), ctypes.byref(out)))
    return ToeplitzFactorization(out.value, kind_cls.kind, vc.dtype, (m, n))


class _RawCuda:
    """``__cuda_array_interface__`` view of a raw device buffer (for torch.distributed)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def spectrum_tensor(F: ToeplitzFactorization) -> torch.Tensor:
    """The handle's spectrum buffer as a uint8 CUDA tensor (no copy) -- what NCCL broadcasts."""
    ptr = ctypes.c_void_p()
    nbytes = ctypes.c_int64()
    check(L.lib().tb200_spectrum_buffer(F._h, ctypes.byref(ptr), ctypes.byref(nbytes)))
    return torch.as_tensor(_RawCuda(ptr.value, nbytes.value), device="cuda")


def factorize_empty_like(kind: int, dtype, shape) -> ToeplitzFactorization:
    out = ctypes.c_void_p()
    m, n = shape
    check(L.lib().tb200_factorize_empty(kind, _DT[dtype], m, n, _stream(), ctypes.byref(out)))
    return ToeplitzFactorization(out.value, kind, dtype, (m, n))


def spectrum_ready(F: ToeplitzFactorization) -> None:
    """Tell the handle that its spectrum buffer was (re)written on the current stream (after the host's own
    collective filled :func:`spectrum_tensor`): products on other streams then wait for it by event."""
    check(L.lib().tb200_spectrum_ready(F._h, _stream()))


def launch_count() -> int:
    return int(L.lib().tb200_launch_count())


def set_option(name: str, value: int) -> None:
    check(L.lib().tb200_set_option(name.encode(), int(value)))


def pass_times():
    """Drain the per-pass CUDA-event timers: ({name: ms}, {name: launches})."""
    ms = (ctypes.c_double * 5)()
    cnt = (ctypes.c_int64 * 5)()
    check(L.lib().tb200_pass_times_ex(ms, cnt, 5))
    names = ("cols_A", "rows_B", "cols_C", "rows_single", "cols_CA")
    return {k: ms[i] for i, k in enumerate(names)}, {k: cnt[i] for i, k in enumerate(names)}
