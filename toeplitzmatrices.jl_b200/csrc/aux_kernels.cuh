// O(n) helper kernels around the FFT passes: circulant-embedding builder (K4),
// spectrum maps for the Circulant family, spectrum un-scramble (eigvals), Strang /
// T. Chan circulant builders, the N < 512 direct product and the fused BLAS-1
// kernels of the cgs / cg iterations (K6).
#pragma once
#include "fft_core.cuh"

namespace tb {

enum Kind : int { K_TOEPLITZ = 0, K_SYMMETRIC = 1, K_CIRCULANT = 2, K_LOWER = 3, K_UPPER = 4, K_HANKEL = 5 };
enum SpecOp : int { OP_INV = 0, OP_PINV = 1, OP_SQRT = 2, OP_CONJ = 3, OP_COPY = 4 };

template <typename R>
__device__ __forceinline__ cx<R> load_a(const void* p, int is_cplx, long long i) {
    if (is_cplx) return reinterpret_cast<const cx<R>*>(p)[i];
    return mk<R>(reinterpret_cast<const R*>(p)[i], 0);
}

// Entry (i, j) of the structured matrix, 0-based (src/toeplitz.jl:47-55, src/hankel.jl:70-73).
template <typename R>
__device__ __forceinline__ cx<R> struct_entry(int kind, const void* vc, const void* vr, int a_cplx,
                                              long long m, long long n, long long i, long long j) {
    switch (kind) {
    case K_HANKEL: return load_a<R>(vc, a_cplx, i + j);
    case K_CIRCULANT: { long long d = i - j; if (d < 0) d += n; return load_a<R>(vc, a_cplx, d); }
    case K_SYMMETRIC: { long long d = i - j; if (d < 0) d = -d; return load_a<R>(vc, a_cplx, d); }
    case K_LOWER: { long long d = i - j; return d >= 0 ? load_a<R>(vc, a_cplx, d) : mk<R>(0, 0); }
    case K_UPPER: { long long d = j - i; return d >= 0 ? load_a<R>(vc, a_cplx, d) : mk<R>(0, 0); }
    default: { long long d = i - j; return d >= 0 ? load_a<R>(vc, a_cplx, d) : load_a<R>(vr, a_cplx, -d); }
    }
}

// K4: c[idx], idx in [0, Np): first column of the circulant the matrix is embedded in.
// Reference: factorize (src/linearalgebra.jl:126-128, 144-147, 188-189, 420-421, 430-432)
// with zeros inserted between the column part and the reversed row part;
// Hankel: the Toeplitz of reverse(H, dims=2) (src/hankel.jl:116-117) built directly from v.
template <typename R>
__global__ void k_embed(cx<R>* __restrict__ c, long long Np, int kind, const void* vc, const void* vr,
                        int a_cplx, long long m, long long n) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < Np;
         idx += (long long)gridDim.x * blockDim.x) {
        cx<R> v = mk<R>(0, 0);
        const long long d = Np - idx;   // row offset for the wrapped part
        switch (kind) {
        case K_TOEPLITZ:
            if (idx < m) v = load_a<R>(vc, a_cplx, idx);
            else if (d < n) v = load_a<R>(vr, a_cplx, d);
            break;
        case K_SYMMETRIC:
            if (idx < m) v = load_a<R>(vc, a_cplx, idx);
            else if (d < n) v = load_a<R>(vc, a_cplx, d);
            break;
        case K_CIRCULANT:
            v = load_a<R>(vc, a_cplx, idx);
            break;
        case K_LOWER:
            if (idx < n) v = load_a<R>(vc, a_cplx, idx);
            break;
        case K_UPPER:
            if (idx == 0) v = load_a<R>(vc, a_cplx, 0);
            else if (d < n) v = load_a<R>(vc, a_cplx, d);
            break;
        case K_HANKEL:                  // vc[i] = v[n-1+i], vr[d] = v[n-1-d]
            if (idx < m) v = load_a<R>(vc, a_cplx, n - 1 + idx);
            else if (d < n) v = load_a<R>(vc, a_cplx, n - 1 - d);
            break;
        }
        c[idx] = v;
    }
}

template <typename R>
__device__ __forceinline__ cx<R> cinv(cx<R> z) {
    // Smith's algorithm (what Julia's inv(::Complex) does, up to scaling details)
    R a = z.x, b = z.y;
    if (fabs(a) >= fabs(b)) {
        R r = b / a, den = a + b * r;
        return mk<R>((R)1 / den, -r / den);
    }
    R r = a / b, den = a * r + b;
    return mk<R>(r / den, (R)-1 / den);
}

template <typename R>
__device__ __forceinline__ cx<R> csqrt_(cx<R> z) {
    // principal square root (Julia sqrt(::Complex) semantics: branch cut on the negative real axis)
    R a = z.x, b = z.y;
    if (a == 0 && b == 0) return mk<R>(0, b);
    R r = hypot(a, b);
    R t = sqrt((r + fabs(a)) / 2);
    if (a >= 0) return mk<R>(t, b / (2 * t));
    return mk<R>(fabs(b) / (2 * t), copysign(t, b));
}

// Elementwise map on a spectrum (layout-agnostic).  src/linearalgebra.jl:246,251,278-281,313,215.
template <typename R>
__global__ void k_spec_map(cx<R>* __restrict__ out, const cx<R>* __restrict__ in, long long N, int op, R tol) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x) {
        cx<R> z = in[i], o;
        switch (op) {
        case OP_INV: o = cinv<R>(z); break;
        case OP_PINV: o = (hypot(z.x, z.y) < tol) ? mk<R>(0, 0) : cinv<R>(z); break;
        case OP_SQRT: o = csqrt_<R>(z); break;
        case OP_CONJ: o = mk<R>(z.x, -z.y); break;
        default: o = z; break;
        }
        out[i] = o;
    }
}

template <typename R>
__global__ void k_spec_mul(cx<R>* __restrict__ out, const cx<R>* __restrict__ a, const cx<R>* __restrict__ b, long long N) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x)
        out[i] = cmul(a[i], b[i]);
}

// Natural-order spectrum out[k], k in [0, Np), from the private layout (eigvals,
// src/linearalgebra.jl:286-287).  two_level: k = k1 + N1*k2, row p = pos1(k1); three-level (log2l0 > 0): k = k0 + N0*kr,
// outer row pos0(k0), and kr addressed inside the row by the child plan's two levels.
template <typename R>
__global__ void k_unscramble(cx<R>* __restrict__ out, const cx<R>* __restrict__ spec, long long Np,
                             int log2l0, int log2l1, int log2l2, int le) {
    PlanDims d2{log2l2, le};
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < Np;
         k += (long long)gridDim.x * blockDim.x) {
        long long off = 0, kr = k;
        if (log2l0 > 0) {
            PlanDims d0{log2l0, le};
            const int k0 = (int)(k & ((1LL << log2l0) - 1));
            kr = k >> log2l0;
            off = (long long)d0.pos_of_freq(k0) << (log2l1 + log2l2);
        }
        if (log2l1 == 0) {
            off += d2.layout_of_pos(d2.pos_of_freq((int)kr));
        } else {
            PlanDims d1{log2l1, le};
            const int k1 = (int)(kr & ((1LL << log2l1) - 1));
            const int k2 = (int)(kr >> log2l1);
            const long long p = d1.pos_of_freq(k1);
            off += (p << log2l2) + d2.layout_of_pos(d2.pos_of_freq(k2));
        }
        out[k] = spec[off];
    }
}

// ---- Bluestein (chirp-z) glue for Circulants whose size is not a power of two ------------------
// DFT_n(x)[k] = c[k] * sum_j T[k-j] (c[j] x[j]),  c[j] = exp(-i pi j^2/n),  T[d] = conj(c[d]):
// a symmetric-Toeplitz product, done by the same FFT-convolution kernels.  These kernels are the
// O(n) diagonal scalings around it.  (needed by factorize(strang(A)) at any n.)
template <typename R>
__global__ void k_bs_pre(cx<R>* __restrict__ u, long long ldu, const void* x, long long ldx, int x_cplx, int conj_x,
                         const cx<R>* __restrict__ chirp, long long n, long long ncols) {
    const long long total = n * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long col = e / n, j = e - col * n;
        cx<R> v = x_cplx ? reinterpret_cast<const cx<R>*>(x)[col * ldx + j]
                         : mk<R>(reinterpret_cast<const R*>(x)[col * ldx + j], 0);
        if (conj_x) v.y = -v.y;
        u[col * ldu + j] = cmul(chirp[j], v);
    }
}
// u2 = chirp .* conj((chirp .* v) .* spec)   -- spectrum multiply + hand-over to the inverse transform
template <typename R>
__global__ void k_bs_mid(cx<R>* __restrict__ u2, const cx<R>* __restrict__ v, long long ld, const cx<R>* __restrict__ chirp,
                         const cx<R>* __restrict__ spec, long long n, long long ncols) {
    const long long total = n * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long col = e / n, j = e - col * n;
        cx<R> X = cmul(cmul(chirp[j], v[col * ld + j]), spec[j]);
        X.y = -X.y;
        u2[col * ld + j] = cmul(chirp[j], X);
    }
}
// y = alpha * maybereal(scale * (conj_out ? conj(chirp .* v) : chirp .* v)) + beta * y
template <typename R>
__global__ void k_bs_post(void* y, long long ldy, int y_cplx, const cx<R>* __restrict__ v, long long ldv,
                          const cx<R>* __restrict__ chirp, long long n, long long ncols, R scale, int conj_out,
                          int take_real, cx<R> alpha, cx<R> beta, int has_beta) {
    const long long total = n * ncols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long col = e / n, j = e - col * n;
        cx<R> t = cmul(chirp[j], v[col * ldv + j]);
        if (conj_out) t.y = -t.y;
        t.x *= scale; t.y *= scale;
        if (take_real) t.y = 0;
        if (y_cplx) {
            cx<R>* q = reinterpret_cast<cx<R>*>(y) + col * ldy + j;
            cx<R> o = cmul(alpha, t);
            if (has_beta) { const cx<R> b = cmul(beta, *q); o.x += b.x; o.y += b.y; }
            *q = o;
        } else {
            R* q = reinterpret_cast<R*>(y) + col * ldy + j;
            R o = alpha.x * t.x;
            if (has_beta) o = fma(beta.x, *q, o);
            *q = o;
        }
    }
}

// smallinv(::LowerTriangularToeplitz) (src/linearalgebra.jl:337-349): first column b of inv(L(v)) by forward
// substitution, n <= 64 (the base case of the recursive-doubling inverse).  One thread: the recurrence is
// sequential and tiny.  Same accumulation order as the reference.
template <typename R>
__global__ void k_tri_smallinv(void* bout, const void* v, int cplx, int n) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    cx<R> b[64];
    const cx<R> v0 = load_a<R>(v, cplx, 0);
    const R den = v0.x * v0.x + v0.y * v0.y;
    const cx<R> iv0 = cplx ? mk<R>(v0.x / den, -v0.y / den) : mk<R>((R)1 / v0.x, 0);
    b[0] = iv0;
    for (int k = 1; k < n; ++k) {
        cx<R> tmp = mk<R>(0, 0);
        for (int i = 0; i < k; ++i) {
            const cx<R> pr = cmul(load_a<R>(v, cplx, k - i), b[i]);
            tmp.x += pr.x; tmp.y += pr.y;
        }
        const cx<R> q = cplx ? cmul(tmp, iv0) : mk<R>(tmp.x / v0.x, 0);
        b[k] = mk<R>(-q.x, -q.y);
    }
    for (int k = 0; k < n; ++k) {
        if (cplx) reinterpret_cast<cx<R>*>(bout)[k] = b[k];
        else reinterpret_cast<R*>(bout)[k] = b[k].x;
    }
}

// trench!(B, r) fill (src/directLinearSolvers.jl:71-85): with y = durbin(r), gamma = 1/(1 + r.y),
// nu = gamma*reverse(y):  B[1,1] = gamma, B[1,2:n] = gamma*y,
// B[i,j] = B[i-1,j-1] + (nu[n+1-j] nu[n+1-i] - nu[i-1] nu[j-1]) / gamma  (2 <= i <= j).
// One thread per diagonal d = j - i walks down its diagonal (the recurrence is sequential along it);
// both triangles are written so the result is the full symmetric inverse; scaled by 1/diag.
template <typename R>
__global__ void k_trench_fill(R* __restrict__ B, long long ldb, const R* __restrict__ y, const double* __restrict__ dots,
                              long long n, R inv_diag) {
    const long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (d >= n) return;
    const R gamma = (R)(1.0 / (1.0 + dots[0]));
    // nu[k] (1-based, k = 1..n-1) = gamma * y[n-k] (1-based y) = gamma * y0[n-1-k] (0-based)
    auto nu = [&](long long k) { return gamma * y[n - 1 - k]; };
    R b = (d == 0) ? gamma : gamma * y[d - 1];                         // B[1, 1+d]
    B[0 + d * ldb] = b * inv_diag;
    B[d + 0 * ldb] = b * inv_diag;
    for (long long i = 2; i + d <= n; ++i) {                            // 1-based i, j = i + d
        const long long j = i + d;
        b += (nu(n + 1 - j) * nu(n + 1 - i) - nu(i - 1) * nu(j - 1)) / gamma;
        B[(i - 1) + (j - 1) * ldb] = b * inv_diag;
        B[(j - 1) + (i - 1) * ldb] = b * inv_diag;
    }
}

// strang / chan (src/linearalgebra.jl:255-274), elementwise on device.
template <typename R>
__global__ void k_strang_chan(void* vout, int which, int kind, const void* vc, const void* vr, int a_cplx, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        cx<R> v;
        if (which == 0) {      // strang: v[i] = A[i,1] for i <= n/2 (0-based), else A[1, n-i+1]
            if (i <= n / 2) v = struct_entry<R>(kind, vc, vr, a_cplx, n, n, i, 0);
            else v = struct_entry<R>(kind, vc, vr, a_cplx, n, n, 0, n - i);
        } else {               // chan: ((n-i)*A[i,1] + i*A[1, min(n-i+1, n)]) / n   (0-based i)
            cx<R> c0 = struct_entry<R>(kind, vc, vr, a_cplx, n, n, i, 0);
            long long jj = n - i; if (jj > n - 1) jj = n - 1;      // 0-based column
            cx<R> r0 = struct_entry<R>(kind, vc, vr, a_cplx, n, n, 0, jj);
            const R w0 = (R)(n - i), w1 = (R)i, inv = (R)n;
            v = mk<R>((w0 * c0.x + w1 * r0.x) / inv, (w0 * c0.y + w1 * r0.y) / inv);
        }
        if (a_cplx) reinterpret_cast<cx<R>*>(vout)[i] = v;
        else reinterpret_cast<R*>(vout)[i] = v.x;
    }
}

// Direct O(mn) product for N = m+n-1 < 512 (src/linearalgebra.jl:19-34): one thread per
// output row, same accumulation order over j as the reference.  grid = (ceil(m/128), ncols).
template <typename R>
__global__ void k_direct_mul(void* y, long long ldy, int y_cplx, const void* x, long long ldx, int x_cplx,
                             int kind, const void* vc, const void* vr, int a_cplx, long long m, long long n,
                             cx<R> alpha, cx<R> beta, int has_beta) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long col = blockIdx.y;
    if (i >= m) return;
    cx<R> acc = mk<R>(0, 0);
    if (has_beta) {
        cx<R> y0 = y_cplx ? reinterpret_cast<cx<R>*>(y)[col * ldy + i]
                          : mk<R>(reinterpret_cast<R*>(y)[col * ldy + i], 0);
        acc = cmul(beta, y0);
    }
    for (long long j = 0; j < n; ++j) {
        cx<R> xj = x_cplx ? reinterpret_cast<const cx<R>*>(x)[col * ldx + j]
                          : mk<R>(reinterpret_cast<const R*>(x)[col * ldx + j], 0);
        cx<R> tmp = cmul(alpha, xj);
        cx<R> aij = (kind == K_HANKEL) ? struct_entry<R>(kind, vc, vr, a_cplx, m, n, i, j)
                                       : struct_entry<R>(kind, vc, vr, a_cplx, m, n, i, j);
        cx<R> pr = cmul(tmp, aij);
        acc.x += pr.x; acc.y += pr.y;
    }
    if (y_cplx) reinterpret_cast<cx<R>*>(y)[col * ldy + i] = acc;
    else reinterpret_cast<R*>(y)[col * ldy + i] = acc.x;
}

// Container algebra on the stored coefficient vectors (src/toeplitz.jl:66-140, src/special.jl:13-33,77-84,
// 223-248, src/hankel.jl:75-130): one elementwise kernel for every dtype combination.  Values travel as complex double
// (exact for every supported input type); dtype codes are the TB200_* ones (0 f32, 1 f64, 2 c32, 3 c64).
enum VecOp : int { VOP_AXPBY = 0, VOP_CONJ = 1, VOP_REAL = 2, VOP_IMAG = 3, VOP_FILL = 4 };

__device__ __forceinline__ double2 vec_load(const void* p, int dt, long long i) {
    switch (dt) {
    case 0: return make_double2((double)reinterpret_cast<const float*>(p)[i], 0.0);
    case 1: return make_double2(reinterpret_cast<const double*>(p)[i], 0.0);
    case 2: { const float2 v = reinterpret_cast<const float2*>(p)[i]; return make_double2((double)v.x, (double)v.y); }
    default: return reinterpret_cast<const double2*>(p)[i];
    }
}
__device__ __forceinline__ void vec_store(void* p, int dt, long long i, double2 v) {
    switch (dt) {
    case 0: reinterpret_cast<float*>(p)[i] = (float)v.x; break;
    case 1: reinterpret_cast<double*>(p)[i] = v.x; break;
    case 2: reinterpret_cast<float2*>(p)[i] = make_float2((float)v.x, (float)v.y); break;
    default: reinterpret_cast<double2*>(p)[i] = v; break;
    }
}
// out[i] = op(a[ia], b[ib]); rev_a / rev_b read the operand back to front (reverse(v), src/hankel.jl:112-121)
__global__ void k_vec_op(int op, long long n, const void* a, int da, int rev_a, const void* b, int db, int rev_b, void* out,
                         int dout, double2 alpha, double2 beta) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 r = make_double2(0.0, 0.0);
        const double2 va = a ? vec_load(a, da, rev_a ? n - 1 - i : i) : r;
        switch (op) {
        case VOP_AXPBY: {
            r.x = alpha.x * va.x - alpha.y * va.y;
            r.y = alpha.x * va.y + alpha.y * va.x;
            if (b) {
                const double2 vb = vec_load(b, db, rev_b ? n - 1 - i : i);
                r.x += beta.x * vb.x - beta.y * vb.y;
                r.y += beta.x * vb.y + beta.y * vb.x;
            }
            break;
        }
        case VOP_CONJ: r = make_double2(va.x, -va.y); break;
        case VOP_REAL: r = make_double2(va.x, 0.0); break;
        case VOP_IMAG: r = make_double2(va.y, 0.0); break;
        default: r = alpha; break;            // VOP_FILL
        }
        vec_store(out, dout, i, r);
    }
}
// number of positions where a != b (==, isequal on the stored vectors); result accumulated into *count
__global__ void k_vec_mismatch(long long n, const void* a, int da, const void* b, int db, unsigned long long* count) {
    unsigned long long c = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 va = vec_load(a, da, i), vb = b ? vec_load(b, db, i) : make_double2(0.0, 0.0);     // b == NULL: iszero
        c += (va.x != vb.x || va.y != vb.y) ? 1ull : 0ull;
    }
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}

}  // namespace tb
