"""NumPy restatement of ToeplitzMatrices.jl's FFT hot path (TEST INFRASTRUCTURE).

Every function cites the reference ``file:line`` (relative to /root/reference)
whose algorithm it follows, including the reference's FFT length
``N = m + n - 1`` (no power-of-two padding), its operation order and its
error behaviour.  The FFT backend is ``scipy.fft`` (pocketfft) standing in for
FFTW.jl, which is a *test-only* dependency of the reference
(``Project.toml:30-38``) reached through ``AbstractFFTs.plan_fft!``:
unnormalised forward ``X_k = sum_j x_j exp(-2 pi i jk/N)``, inverse scaled by
``1/N``.  Parity unpinned (see ``oracle/__init__.py``).

Indices are 0-based here; comments give the reference's 1-based form.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as _fft


class DimensionMismatch(ValueError):
    """Mirror of Julia's ``DimensionMismatch``."""


# ----------------------------------------------------------------------------
# dtype rules
# ----------------------------------------------------------------------------
def spectrum_dtype(T):
    """``S = promote_type(float(T), Complex{Float32})`` (src/linearalgebra.jl:125)."""
    T = np.dtype(T)
    if T in (np.dtype(np.float32), np.dtype(np.complex64)):
        return np.dtype(np.complex64)
    return np.dtype(np.complex128)


def maybereal(T, x):
    """src/ToeplitzMatrices.jl:114-115."""
    if np.dtype(T).kind != "c":
        return np.real(x)
    return x


def _fwd(a, workers=None):
    return _fft.fft(a, workers=workers)


def _inv(a, workers=None):
    return _fft.ifft(a, workers=workers)


# ----------------------------------------------------------------------------
# containers (only what the hot path reads: .vc, .vr, size, getindex)
# ----------------------------------------------------------------------------
class AbstractToeplitz:
    kind = "abstract"

    @property
    def dtype(self):
        return self.vc.dtype

    @property
    def shape(self):
        return (len(self.vc), len(self.vr))

    def getindex(self, i, j):
        """0-based; src/toeplitz.jl:47-55."""
        d = i - j
        return self.vc[d] if d >= 0 else self.vr[-d]

    def dense(self):
        m, n = self.shape
        i = np.arange(m)[:, None]
        j = np.arange(n)[None, :]
        d = i - j
        vc = np.asarray(self.vc)
        vr = np.asarray(self.vr)
        return np.where(d >= 0, vc[np.clip(d, 0, m - 1)], vr[np.clip(-d, 0, n - 1)])


class Toeplitz(AbstractToeplitz):
    """src/toeplitz.jl:7-18."""
    kind = "toeplitz"

    def __init__(self, vc, vr):
        vc = np.asarray(vc)
        vr = np.asarray(vr)
        T = np.result_type(vc, vr)
        self.vc = vc.astype(T)
        self.vr = vr.astype(T)
        if self.vc[0] != self.vr[0]:
            raise ValueError("First element of the vectors must be the same")

    def adjoint(self):
        """src/toeplitz.jl:100-101: eager ``transpose(conj(A))``."""
        return Toeplitz(np.conj(self.vr), np.conj(self.vc))

    def transpose(self):
        return Toeplitz(self.vr, self.vc)


class SymmetricToeplitz(AbstractToeplitz):
    """src/special.jl:44-76,79-85: vc === vr === v."""
    kind = "symmetric"

    def __init__(self, v):
        self.v = np.asarray(v)
        self.vc = self.v
        self.vr = self.v

    def adjoint(self):
        return SymmetricToeplitz(np.conj(self.v)) if self.v.dtype.kind == "c" else self

    def transpose(self):
        return self


class Circulant(AbstractToeplitz):
    """src/special.jl:86-94,119: vr = first row = [v1, vn, vn-1, ..., v2]."""
    kind = "circulant"

    def __init__(self, v):
        self.v = np.asarray(v)
        self.vc = self.v
        self.vr = np.concatenate([self.v[:1], self.v[:0:-1]])

    def adjoint(self):
        return Circulant(np.conj(self.vr))

    def transpose(self):
        return Circulant(self.vr)


class LowerTriangularToeplitz(AbstractToeplitz):
    """src/special.jl:95-105,120-124."""
    kind = "lower"
    uplo = "L"

    def __init__(self, v):
        self.v = np.asarray(v)
        self.vc = self.v
        w = np.zeros_like(self.v)
        w[0] = self.v[0]
        self.vr = w

    def adjoint(self):
        return UpperTriangularToeplitz(np.conj(self.v))

    def transpose(self):
        return UpperTriangularToeplitz(self.v)


class UpperTriangularToeplitz(AbstractToeplitz):
    """src/special.jl:106-117."""
    kind = "upper"
    uplo = "U"

    def __init__(self, v):
        self.v = np.asarray(v)
        self.vr = self.v
        w = np.zeros_like(self.v)
        w[0] = self.v[0]
        self.vc = w

    def adjoint(self):
        return LowerTriangularToeplitz(np.conj(self.v))

    def transpose(self):
        return LowerTriangularToeplitz(self.v)


class Hankel:
    """src/hankel.jl:2-36: H[i,j] = v[i+j-1] (1-based)."""
    kind = "hankel"

    def __init__(self, v, size=None, vr=None):
        if vr is not None:  # Hankel(vc, vr)  (src/hankel.jl:20-25)
            vc = np.asarray(v)
            vr = np.asarray(vr)
            if vc[-1] != vr[0]:
                raise ValueError("vc[end] != vr[1]")
            T = np.result_type(vc, vr)
            v = np.concatenate([vc.astype(T), vr[1:].astype(T)])
            size = (len(vc), len(vr))
        self.v = np.asarray(v)
        if size is None:
            l = len(self.v)
            size = ((l + 1) // 2, (l + 1) // 2)
        m, n = size
        if len(self.v) < m + n - 1:
            raise ValueError("inconsistency between size and number of anti-diagonals")
        self.shape = (m, n)

    @property
    def dtype(self):
        return self.v.dtype

    def dense(self):
        m, n = self.shape
        i = np.arange(m)[:, None]
        j = np.arange(n)[None, :]
        return self.v[i + j]

    def reverse(self, dims=1):
        """src/hankel.jl:112-121."""
        m, n = self.shape
        if dims == 1:
            return Toeplitz(self.v[:m][::-1], self.v[m - 1:])
        if dims == 2:
            return Toeplitz(self.v[n - 1:], self.v[n - 1::-1])
        raise ValueError("invalid dimension %r in reverse" % (dims,))


# ----------------------------------------------------------------------------
# factorize (src/linearalgebra.jl:122-131,140-150,184-192,415-435)
# ----------------------------------------------------------------------------
class ToeplitzFactorization:
    """src/ToeplitzMatrices.jl:97-101 -- vcvr_dft (+ the eltype/kind type params)."""

    def __init__(self, vcvr_dft, T, kind, workers=None):
        self.vcvr_dft = vcvr_dft
        self.T = np.dtype(T)
        self.kind = kind
        self.workers = workers

    @property
    def shape(self):
        """src/special.jl:262-263 (Circulant only in the reference)."""
        s = len(self.vcvr_dft)
        return (s, s)

    def adjoint(self):
        """src/linearalgebra.jl:214-215 (Circulant factorization)."""
        return ToeplitzFactorization(np.conj(self.vcvr_dft), self.T, self.kind, self.workers)


def embedding_vector(A):
    """The length-N vector whose DFT ``factorize`` stores, as type S."""
    S = spectrum_dtype(A.dtype)
    if A.kind == "toeplitz":                      # :126-128
        m, n = A.shape
        tmp = np.empty(m + n - 1, dtype=S)
        tmp[:m] = A.vc
        tmp[m:] = A.vr[:0:-1]                     # vr[n], vr[n-1], ..., vr[2]
    elif A.kind == "symmetric":                   # :144-147
        m = len(A.vc)
        tmp = np.empty(2 * m, dtype=S)
        tmp[:m] = A.vc
        tmp[m] = 0
        tmp[m + 1:] = A.vc[:0:-1]
    elif A.kind == "circulant":                   # :188-189
        tmp = np.array(A.vc, dtype=S)
    elif A.kind == "lower":                       # :420-421
        n = len(A.v)
        tmp = np.zeros(2 * n - 1, dtype=S)
        tmp[:n] = A.v
    elif A.kind == "upper":                       # :430-432
        n = len(A.v)
        tmp = np.zeros(2 * n - 1, dtype=S)
        tmp[0] = A.v[0]
        tmp[n:] = A.v[:0:-1]
    else:
        raise TypeError(A.kind)
    return tmp


def factorize(A, workers=None):
    tmp = embedding_vector(A)
    return ToeplitzFactorization(_fwd(tmp, workers), A.dtype, A.kind, workers)


# ----------------------------------------------------------------------------
# mul!  (src/linearalgebra.jl:4-99, src/hankel.jl:133-137)
# ----------------------------------------------------------------------------
def _promote(*objs):
    ds = []
    for o in objs:
        if isinstance(o, (ToeplitzFactorization,)):
            ds.append(o.T)
        elif hasattr(o, "dtype"):
            ds.append(np.dtype(o.dtype))
        elif isinstance(o, bool):
            ds.append(np.dtype(np.bool_))
        elif isinstance(o, int):
            ds.append(np.dtype(np.int64))
        elif isinstance(o, float):
            ds.append(np.dtype(np.float64))
        elif isinstance(o, complex):
            ds.append(np.dtype(np.complex128))
        else:
            ds.append(np.dtype(type(o)))
    # Julia never widens Float32 because of a Bool/Int/Float64 *literal* used as
    # alpha/beta in the tests; numpy's result_type on dtypes does.  Callers pass
    # typed scalars, so plain promotion is right for arrays and typed scalars.
    return np.result_type(*ds)


def mul_factorization(y, F, x, alpha=True, beta=False):
    """``mul!(y, A::ToeplitzFactorization, x, alpha, beta)`` src/linearalgebra.jl:42-80."""
    n = len(x)
    m = len(y)
    vcvr_dft = F.vcvr_dft
    N = len(vcvr_dft)
    if m > N or n > N:
        raise DimensionMismatch(
            "Toeplitz factorization does not match size of input and output vector")
    T = _promote(y, F, x, alpha, beta)
    tmp = np.zeros(N, dtype=vcvr_dft.dtype)
    tmp[:n] = x                                   # :59-62
    tmp = _fwd(tmp, F.workers)                    # :63
    tmp *= vcvr_dft                               # :64-66
    tmp = _inv(tmp, F.workers)                    # :67
    t = maybereal(T, tmp[:m])
    if beta == 0:                                 # iszero(beta): y is *not* read  (:68-71)
        y[:] = alpha * t
    else:
        y[:] = alpha * t + beta * y               # muladd(alpha, t, beta*y)  (:72-76)
    return y


def mul_vec(y, A, x, alpha=True, beta=False, workers=None):
    """``mul!(y, A::AbstractToeplitz, x, alpha, beta)`` src/linearalgebra.jl:4-41."""
    if isinstance(A, Hankel):                     # src/hankel.jl:134
        return mul_vec(y, A.reverse(dims=2), x[::-1], alpha, beta, workers)
    m, n = A.shape
    if len(y) != m:
        raise DimensionMismatch(
            "first dimension of A, %d, does not match length of y, %d" % (m, len(y)))
    if len(x) != n:
        raise DimensionMismatch(
            "second dimension of A, %d, does not match length of x, %d" % (n, len(x)))
    N = m + n - 1
    if N < 512:                                   # :19-34 direct O(mn) loop
        if beta == 0:
            y[:] = 0
        else:
            y *= beta
        vc, vr = A.vc, A.vr
        for j in range(n):
            tmp = alpha * x[j]
            # column j of A: A[i,j] = vc[i-j] (i>=j) else vr[j-i]
            col = np.empty(m, dtype=np.result_type(vc, vr))
            lo = min(j, m)
            col[:lo] = vr[j:j - lo:-1] if lo > 0 else []
            if m > j:
                col[j:] = vc[:m - j]
            y += tmp * col
        return y
    return mul_factorization(y, factorize(A, workers), x, alpha, beta)


def mul_mat(C, A, B, alpha=True, beta=False, workers=None):
    """``mul!(C, A, B, alpha, beta)`` src/linearalgebra.jl:83-99 (serial column loop)."""
    if isinstance(A, Hankel):                     # src/hankel.jl:137
        return mul_mat(C, A.reverse(dims=2), B[::-1, :], alpha, beta, workers)
    F = A if isinstance(A, ToeplitzFactorization) else factorize(A, workers)
    l = B.shape[1]
    if C.shape[1] != l:
        raise DimensionMismatch("input and output matrices must have same number of columns")
    for j in range(l):
        mul_factorization(C[:, j], F, B[:, j], alpha, beta)
    return C


def _out_dtype(A, x):
    return np.result_type(A.dtype, x.dtype)


def matvec(A, x, workers=None):
    """``A * x`` (vector or matrix) -> new array."""
    m = A.shape[0]
    T = _out_dtype(A, x)
    if x.ndim == 1:
        return mul_vec(np.zeros(m, dtype=T), A, x, True, False, workers)
    return mul_mat(np.zeros((m, x.shape[1]), dtype=T, order="F"), A, x, True, False, workers)


# ----------------------------------------------------------------------------
# Circulant spectral family (src/linearalgebra.jl:183-315)
# ----------------------------------------------------------------------------
def circ_ldiv_factorization(F, b):
    """``ldiv!(C::CirculantFactorization, b)`` :225-242 (in place)."""
    n = len(b)
    if len(F.vcvr_dft) != n:
        raise DimensionMismatch(
            "size of Toeplitz factorization does not match the length of the output vector")
    tmp = np.array(b, dtype=F.vcvr_dft.dtype)
    tmp = _fwd(tmp, F.workers)
    tmp /= F.vcvr_dft
    tmp = _inv(tmp, F.workers)
    b[:] = maybereal(F.T, tmp)
    return b


def circ_ldiv(C, b, workers=None):
    """``ldiv!(C::Circulant, b)`` :224 -- re-factorizes per call."""
    return circ_ldiv_factorization(factorize(C, workers), b)


def ldiv_mat(A, B, solver):
    """``ldiv!(A::AbstractToeplitz, B::StridedMatrix)`` :102-110."""
    if A.shape[0] != A.shape[1]:
        raise RuntimeError("Division: Rectangular case is not supported.")
    for j in range(B.shape[1]):
        solver(A, B[:, j])
    return B


def circ_mul(FA, FB):
    """``A::CirculantFactorization * B::CirculantFactorization`` :197-211."""
    m, n = len(FA.vcvr_dft), len(FB.vcvr_dft)
    if m != n:
        raise DimensionMismatch(
            "size of matrix A, %dx%d, does not match size of matrix B, %dx%d" % (m, m, n, n))
    vc = _inv(FA.vcvr_dft * FB.vcvr_dft)
    return Circulant(maybereal(np.result_type(FA.T, FB.T), vc))


def circ_inv(C):
    """``inv(C::Circulant)`` :244-249."""
    F = factorize(C)
    vc = _inv(1.0 / F.vcvr_dft)
    return Circulant(maybereal(C.dtype, vc))


def circ_inv_factorization(F):
    """``inv(C::CirculantFactorization)`` :250-253."""
    return ToeplitzFactorization(1.0 / F.vcvr_dft, F.T, F.kind, F.workers)


def circ_pinv(C, tol=None):
    """``pinv(C::Circulant, tol)`` :276-284."""
    if tol is None:
        tol = np.finfo(np.zeros(1, C.dtype).real.dtype if C.dtype.kind in "fc" else np.float64).eps
    F = factorize(C)
    s = F.vcvr_dft
    with np.errstate(divide="ignore", invalid="ignore"):
        z = 1.0 / s
    z = np.where(np.abs(s) < tol, 0, z).astype(s.dtype)
    return Circulant(maybereal(C.dtype, _inv(z)))


def circ_sqrt(C):
    """``sqrt(C::Circulant)`` :311-315."""
    F = factorize(C)
    return Circulant(maybereal(C.dtype, _inv(np.sqrt(F.vcvr_dft))))


def circ_eigvals(C):
    """``eigvals(C)`` :286-287 -- a copy of the spectrum, unsorted, always complex."""
    F = C if isinstance(C, ToeplitzFactorization) else factorize(C)
    return F.vcvr_dft.copy()


def circ_det(C):
    """``det`` :288-290."""
    d = np.prod(circ_eigvals(C))
    return d.real if C.dtype.kind != "c" else d


def strang(A):
    """``strang(A)`` :255-266."""
    n = A.shape[0]
    v = np.empty(n, dtype=A.dtype)
    n2 = n // 2
    for i in range(1, n2 + 2):                    # v[i] = A[i,1]
        v[i - 1] = A.getindex(i - 1, 0)
    for i in range(n2 + 2, n + 1):                # v[i] = A[1, n-i+2]
        v[i - 1] = A.getindex(0, n - i + 1)
    return Circulant(v)


def chan(A):
    """``chan(A)`` :267-274."""
    n = A.shape[0]
    v = np.empty(n, dtype=np.result_type(A.dtype, np.float32) if A.dtype.kind in "iub" else A.dtype)
    for i in range(1, n + 1):
        v[i - 1] = ((n - i + 1) * A.getindex(i - 1, 0)
                    + (i - 1) * A.getindex(0, min(n - i + 2, n) - 1)) / n
    return Circulant(v)


# ----------------------------------------------------------------------------
# iterative solvers (src/iterativeLinearSolvers.jl)
# ----------------------------------------------------------------------------
def _apply_prec(M, z):
    """``ldiv!(M, z)`` for M a Circulant (re-factorized, :224) or a factorization."""
    if isinstance(M, ToeplitzFactorization):
        return circ_ldiv_factorization(M, z)
    return circ_ldiv(M, z)


def _norm(v):
    return float(np.sqrt(np.sum(np.abs(v.astype(np.result_type(v.dtype, np.float32))) ** 2)))


def cgs(A, x, b, M, max_it, tol):
    """``IterativeLinearSolvers.cgs`` src/iterativeLinearSolvers.jl:99-215.

    Returns (x, error, iter, flag); ``dot`` conjugates its first argument as
    Julia's does.
    """
    T = x.dtype
    it = 0
    flag = 0
    n = len(b)
    bnrm2 = _norm(b)
    if bnrm2 == 0:
        bnrm2 = 1.0
    u = np.zeros(n, T); p = np.zeros(n, T); ph = np.zeros(n, T)
    q = np.zeros(n, T); uh = np.zeros(n, T); vh = np.zeros(n, T)
    one = T.type(1)
    rho = T.type(0)
    rho1 = rho
    r = b.astype(T).copy()
    mul_vec(r, A, x, -one, one)                   # :150
    error = _norm(r) / bnrm2
    if error < tol:
        return x, error, it, flag
    r_tld = r.copy()
    for it in range(1, max_it + 1):
        rho = np.vdot(r_tld, r)                   # :163
        if rho == 0:
            break
        if it > 1:
            beta = rho / rho1
            u[:] = r + beta * q                   # :171
            p[:] = u + beta * (q + beta * p)      # :172
        else:
            u[:] = r
            p[:] = u
        ph[:] = p
        _apply_prec(M, ph)                        # :180
        mul_vec(vh, A, ph, one, T.type(0))        # :182
        alpha = rho / np.vdot(r_tld, vh)          # :184
        q[:] = u - alpha * vh                     # :186
        uh[:] = u + q                             # :187
        _apply_prec(M, uh)                        # :189
        x += alpha * uh                           # :193
        mul_vec(r, A, uh, -alpha, one)            # :196
        error = _norm(r) / bnrm2
        if error <= tol:
            break
        rho1 = rho
    if error <= tol:
        flag = 0
    elif rho == 0:
        flag = -1
    else:
        flag = 1
    return x, error, it, flag


def cg(A, x, b, M, max_it, tol):
    """``IterativeLinearSolvers.cg`` src/iterativeLinearSolvers.jl:9-97.

    Returns ``None`` on the reference's value-less early ``return`` (:57-59).
    """
    T = x.dtype
    n = len(b)
    flag = 0
    it = 0
    bnrm2 = _norm(b)
    if bnrm2 == 0.0:
        bnrm2 = 1.0
    z = np.zeros(n, T); q = np.zeros(n, T); p = np.zeros(n, T)
    r = b - matvec(A, x)                          # :55
    error = _norm(r) / bnrm2
    if error < tol:
        return None
    rho1 = None
    for it in range(1, max_it + 1):
        z[:] = r
        _apply_prec(M, z)                         # :65
        rho = np.vdot(r, z)
        if it > 1:
            beta = rho / rho1
            p[:] = z + beta * p
        else:
            p[:] = z
        q[:] = matvec(A, p)                       # :79
        alpha = rho / np.vdot(p, q)
        x += alpha * p
        r -= alpha * q
        error = _norm(r) / bnrm2
        if error <= tol:
            break
        rho1 = rho
    if error > tol:
        flag = 1
    return x, error, it, flag


def _eps100():
    return 100 * np.finfo(np.float64).eps


def ldiv_toeplitz(A, b):
    """``ldiv!(A::Toeplitz, b)`` src/linearalgebra.jl:133-136 (cgs + Strang)."""
    pre = factorize(strang(A))
    b[:] = cgs(A, np.zeros(len(b), b.dtype), b, pre, 1000, _eps100())[0]
    return b


def ldiv_symmetric(A, b):
    """``ldiv!(A::SymmetricToeplitz, b)`` :152 (cg + unfactorized Strang, Float64 x0)."""
    res = cg(A, np.zeros(len(b)), b, strang(A), 1000, _eps100())
    if res is None:
        raise TypeError("cg returned nothing (reference indexes `nothing[1]`)")
    b[:] = res[0]
    return b


def ldiv_triangular(A, b):
    """``ldiv!(A::TriangularToeplitz, b)`` :399-402 (cgs + T. Chan)."""
    pre = factorize(chan(A))
    b[:] = cgs(A, np.zeros(len(b), b.dtype), b, pre, 1000, _eps100())[0]
    return b


def ldiv(A, b):
    """Dispatch of ``ldiv!(A, b::Vector)`` by matrix type."""
    if A.kind == "circulant":
        return circ_ldiv(A, b)
    if A.kind == "toeplitz":
        return ldiv_toeplitz(A, b)
    if A.kind == "symmetric":
        return ldiv_symmetric(A, b)
    if A.kind in ("lower", "upper"):
        return ldiv_triangular(A, b)
    raise TypeError(A.kind)


# ----------------------------------------------------------------------------
# direct solvers (src/directLinearSolvers.jl) -- pure-Python, small sizes only;
# the C restatement in oracle/levinson_oracle.c is the batched/timed one.
# ----------------------------------------------------------------------------
def reverse_dot(x, y):
    """:137-145  ``sum_i x[i]*y[n-i+1]`` accumulated left to right."""
    n = len(x)
    if n != len(y):
        raise DimensionMismatch()
    d = np.result_type(x.dtype, y.dtype).type(0)
    for i in range(n):
        d += x[i] * y[n - 1 - i]
    return d


def reverse_increment(x, y=None, alpha=1):
    """:148-168  ``x += alpha*reverse(y)``, alias-safe when ``x is y``."""
    if y is None:
        y = x
    n = len(x)
    if n != len(y):
        raise DimensionMismatch()
    if x is y or (np.shares_memory(x, y) and x.ctypes.data == y.ctypes.data):
        for i in range(n // 2):
            y_i = y[i]
            z_i = y[n - 1 - i]
            x[i] += alpha * z_i
            x[n - 1 - i] += alpha * y_i
        if n % 2 == 1:
            i = n // 2
            x[i] += alpha * y[i]
    else:
        for i in range(n):
            x[i] += alpha * y[n - 1 - i]
    return x


def durbin_(y, r):
    """``durbin!(y, r)`` :15-28."""
    n = len(r)
    if len(y) != n:
        raise DimensionMismatch("length(y) = %d != %d = length(r)" % (len(y), n))
    y[0] = -r[0]
    alpha, beta = -r[0], r.dtype.type(1)
    for k in range(1, n):
        beta *= (1 - alpha ** 2)
        r_k, y_k = r[:k], y[:k]
        alpha = -(r[k] + reverse_dot(r_k, y_k)) / beta
        reverse_increment(y_k, y_k, alpha)
        y[k] = alpha
    return y


def durbin(r):
    """``durbin(r)`` :8."""
    return durbin_(np.zeros_like(r), r)


def levinson_(x, r, b, y=None):
    """``levinson!(x, r, b, y)`` :100-121."""
    n = len(r) + 1
    if len(x) != n:
        raise DimensionMismatch("length(x) = %d != %d = length(r) + 1" % (len(x), n))
    if len(b) != n:
        raise DimensionMismatch("length(b) = %d != %d = length(r) + 1" % (len(b), n))
    if y is None:
        y = np.zeros_like(x)
    y[0] = -r[0]
    x[0] = b[0]
    alpha, beta = -r[0], r.dtype.type(1)
    for k in range(1, n):
        beta *= (1 - alpha ** 2)
        r_k, x_k, y_k = r[:k], x[:k], y[:k]
        mu = (b[k] - reverse_dot(r_k, x_k)) / beta
        reverse_increment(x_k, y_k, mu)
        x[k] = mu
        if k < n - 1:
            alpha = -(r[k] + reverse_dot(r_k, y_k)) / beta
            reverse_increment(y_k, y_k, alpha)
            y[k] = alpha
    return x


def levinson(r_or_T, b):
    """``levinson(r, b)`` :93 and ``levinson(T::SymmetricToeplitz, b)`` :123-134."""
    if isinstance(r_or_T, SymmetricToeplitz):
        T = r_or_T
        r_0 = T.vc[0]
        r = T.vc[1:]
        if r_0 != 1:
            r = r / r_0
        x = levinson_(np.zeros_like(b, dtype=np.result_type(b.dtype, r.dtype)), np.asarray(r), b)
        if r_0 != 1:
            x /= r_0
        return x
    r = np.asarray(r_or_T)
    return levinson_(np.zeros_like(b, dtype=np.result_type(b.dtype, r.dtype)), r, b)


def trench(r_or_T):
    """``trench`` :36-85 (upper triangle filled, returned symmetrised)."""
    if isinstance(r_or_T, SymmetricToeplitz):
        T = r_or_T
        r_0 = T.vc[0]
        r = T.vc[1:]
        if r_0 != 1:
            r = r / r_0
        B = trench(np.asarray(r))
        if r_0 != 1:
            B = B / r_0
        return B
    r = np.asarray(r_or_T)
    n = len(r) + 1
    B = np.zeros((n, n), dtype=r.dtype)
    y = durbin(r)
    gamma = 1 / (1 + np.dot(r, y))
    nu = gamma * y[::-1]
    B[0, 0] = gamma
    B[0, 1:] = gamma * y
    for j in range(2, n + 1):
        for i in range(2, j + 1):
            B[i - 1, j - 1] = B[i - 2, j - 2] + (nu[n - j] * nu[n - i] - nu[i - 2] * nu[j - 2]) / gamma
    return np.triu(B) + np.triu(B, 1).T


# ----------------------------------------------------------------------------
# triangular inverse by recursive doubling (src/linearalgebra.jl:337-396), SURVEY 8f rank 2
# ----------------------------------------------------------------------------
def tri_smallinv_lower(v):
    """``smallinv(A::LowerTriangularToeplitz)`` :337-349 -> first column of the inverse."""
    n = len(v)
    b = np.zeros(n, dtype=v.dtype)
    b[0] = 1 / v[0]
    for k in range(1, n):
        tmp = v.dtype.type(0)
        for i in range(k):
            tmp += v[k - i] * b[i]
        b[k] = -tmp / v[0]
    return b


def tri_inv_lower(v):
    """``inv(A::LowerTriangularToeplitz)`` :350-365 -> first column of the inverse."""
    n = len(v)
    if n <= 64:
        return tri_smallinv_lower(v)
    np2 = 1 << (n - 1).bit_length()
    if n != np2:
        return tri_inv_lower(np.concatenate([v, np.zeros(np2 - n, dtype=v.dtype)]))[:n]
    nd2 = n // 2
    a1 = tri_inv_lower(v[:nd2])
    T = Toeplitz(v[nd2:], v[nd2:0:-1])
    t = matvec(T, a1)
    return np.concatenate([a1, -matvec(LowerTriangularToeplitz(a1), t)])
