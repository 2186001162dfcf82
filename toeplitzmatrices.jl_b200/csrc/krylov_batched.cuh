// Batched Krylov solvers: the reference's matrix `ldiv!` is a serial column loop over single-vector cgs / cg
// solves (/root/reference/src/linearalgebra.jl:102-110 calling :133-136, :152, :399-402).  Here all columns of a chunk
// iterate together: the structured products and preconditioner solves run on the batched FFT path (two real columns
// per transform), the vector updates are single sweeps over the whole chunk, and every per-column scalar
// (rho, alpha, beta, error, flags) lives on the device.  After every iteration `k_compact` partitions the column indices
// into "still iterating" | "frozen" (stable), and the next iteration's products, preconditioner solves and sweeps run on
// the first group only (column indirection in the FFT kernels, fft_kernels.cuh) -- a straggler column costs one column's
// work, not the whole chunk's.  The host reads one integer (columns still iterating) per iteration, or every third
// iteration for small problems; single-vector cgs / cg are the one-column case of the same code.  Per column the arithmetic is the reference's
// (/root/reference/src/iterativeLinearSolvers.jl:99-215 cgs, :9-97 cg): a column that converges, breaks down or
// returns early is frozen with its own (error, iter, flag) while the others go on.
#pragma once
#include "blas1.cuh"

namespace tb {

struct KState {
    double rho[2], rho1[2], rhon[2], alpha[2], beta[2], dot[4];
    double bnrm2, error;
    double scale, iscale;          // power of two that brings ||b_j|| into [0.5, 1), and its inverse (see KS_BNRM)
    int iter, flag, active, pad;
};

__device__ __forceinline__ double2 zdiv(const double* a, const double* b) {
    const double d = b[0] * b[0] + b[1] * b[1];
    return make_double2((a[0] * b[0] + a[1] * b[1]) / d, (a[1] * b[0] - a[0] * b[1]) / d);
}

// dot[0..1] = dot(a1_j, b1_j), dot[2..3] = dot(a2_j, b2_j) for every column j = blockIdx.y (all four matrices share the
// leading dimension ld unless ld1 is given for the first pair).  Deterministic: per-CTA partials, the last CTA of a
// column sums them in index order.
template <typename T>
__global__ void __launch_bounds__(256)
k_bdot2(const T* __restrict__ a1, const T* __restrict__ b1, long long ld1, const T* __restrict__ a2, const T* __restrict__ b2,
        long long ld2, long long n, double* __restrict__ partials, unsigned* __restrict__ counters, KState* __restrict__ st,
        int only_active, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    if (only_active && !st[j].active) return;
    const T* pa1 = a1 + j * ld1; const T* pb1 = b1 + j * ld1;
    const T* pa2 = a2 ? a2 + j * ld2 : nullptr; const T* pb2 = b2 ? b2 + j * ld2 : nullptr;
    double acc[4] = {0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        dot_acc(pa1[i], pb1[i], acc[0], acc[1]);
        if (pa2) dot_acc(pa2[i], pb2[i], acc[2], acc[3]);
    }
    __shared__ double red[4][8];
    __shared__ bool is_last;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], o);
        if ((threadIdx.x & 31) == 0) red[v][threadIdx.x >> 5] = acc[v];
    }
    __syncthreads();
    double* mine = partials + ((size_t)j * gridDim.x) * 4;
    if (threadIdx.x == 0) {
        for (int v = 0; v < 4; ++v) {
            double s = 0;
            for (int w = 0; w < 8; ++w) s += red[v][w];
            mine[(size_t)blockIdx.x * 4 + v] = s;
        }
        __threadfence();
        const unsigned done = atomicAdd(counters + j, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        // the last CTA of the column sums the per-CTA partials in a fixed order (deterministic), all 256 threads at it:
        // one thread walking 2368 partials through L2 took longer than the sweep itself at n = 2^24
        const double4 s = final_sum4(mine, gridDim.x, red);
        if (threadIdx.x == 0) {
            st[j].dot[0] = s.x; st[j].dot[1] = s.y; st[j].dot[2] = s.z; st[j].dot[3] = s.w;
            counters[j] = 0;
        }
    }
}

// ---- per-column scalar steps (one thread per column) ---------------------------------------------------------
enum KStep : int { KS_BNRM = 0, KS_INIT_CGS, KS_INIT_CG, KS_RHO0, KS_CGS_BEGIN, KS_CGS_ALPHA, KS_CGS_END, KS_CG_BEGIN, KS_CG_ALPHA, KS_CG_END };

__global__ void k_kstep(KState* __restrict__ st, int ncols, int step, int it, int max_it, double tol) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    KState& s = st[j];
    switch (step) {
    case KS_BNRM:                                   // :136 / :44-47
        s.bnrm2 = sqrt(s.dot[0]);
        if (s.bnrm2 == 0.0) s.bnrm2 = 1.0;
        // Every column iterates in units of a power of two near its own ||b||: real columns share complex transforms two
        // by two, and a product's rounding error is relative to the larger column of a pair, so a right-hand side 10^k
        // times smaller than its partner would stall k digits above the tolerance.  Scaling by a power of two is exact:
        // the iterates are those of the unscaled column, bit for bit, and x is scaled back at the end.
        s.scale = 1.0; s.iscale = 1.0;
        if (s.bnrm2 > 0.0 && s.bnrm2 < 1.0e300) {
            int e = 0;
            frexp(s.bnrm2, &e);
            const int emax = it > 0 ? it : 900;          // KS_BNRM: `it` carries the exponent range of the vectors' precision
            e = e < -emax ? -emax : (e > emax ? emax : e);
            s.scale = ldexp(1.0, -e);
            s.iscale = ldexp(1.0, e);
            s.bnrm2 *= s.scale;
        }
        s.iter = 0; s.flag = 0; s.active = 1; s.pad = 0;
        s.rho[0] = s.rho[1] = s.rho1[0] = s.rho1[1] = 0.0;
        break;
    case KS_INIT_CGS:                               // :152-156
    case KS_INIT_CG:                                // :55-59 (flag 2 = the reference's value-less early return)
        s.error = sqrt(s.dot[0]) / s.bnrm2;
        if (s.error < tol) { s.active = 0; s.flag = (step == KS_INIT_CG) ? 2 : 0; }
        else if (max_it <= 0) { s.active = 0; s.flag = (step == KS_INIT_CG) ? 1 : -1; }    // empty loop: :205-213 with rho == 0 / :93-95
        else if (s.error != s.error) {              // NaN right-hand side: frozen before it shares a transform (see KS_CGS_END)
            s.active = 0; s.flag = (step == KS_INIT_CG) ? 0 : 1; s.pad = 1; s.iter = max_it;
        }
        break;
    case KS_RHO0:
        s.rhon[0] = s.dot[0]; s.rhon[1] = s.dot[1];
        break;
    case KS_CGS_BEGIN:                              // :163-169
        if (!s.active) break;
        s.iter = it;
        s.rho[0] = s.rhon[0]; s.rho[1] = s.rhon[1];
        if (s.rho[0] == 0.0 && s.rho[1] == 0.0) { s.active = 0; s.flag = (s.error <= tol) ? 0 : -1; break; }
        if (it > 1) { const double2 b = zdiv(s.rho, s.rho1); s.beta[0] = b.x; s.beta[1] = b.y; }
        else { s.beta[0] = s.beta[1] = 0.0; }
        break;
    case KS_CGS_ALPHA:                              // :184
    case KS_CG_ALPHA: {                             // :80
        if (!s.active) break;
        const double2 a = zdiv(s.rho, s.dot);
        s.alpha[0] = a.x; s.alpha[1] = a.y;
        break;
    }
    case KS_CGS_END:                                // :198-203, flag logic :205-213
        if (!s.active) break;
        s.error = sqrt(s.dot[0]) / s.bnrm2;
        s.rhon[0] = s.dot[2]; s.rhon[1] = s.dot[3];
        if (s.error <= tol) { s.active = 0; s.flag = 0; }
        else if (s.error != s.error) {                      // NaN: the reference iterates on to max_it with NaN everywhere
            s.active = 0; s.flag = 1; s.pad = (it < max_it); s.iter = max_it;   // (rho is NaN, not 0: flag 1); frozen here so that
        }                                                   // the column leaves the pair it shares a transform with
        else {
            s.rho1[0] = s.rho[0]; s.rho1[1] = s.rho[1];
            if (it == max_it) { s.active = 0; s.flag = 1; }
        }
        break;
    case KS_CG_BEGIN:                               // :67-68
        if (!s.active) break;
        s.iter = it;
        s.rho[0] = s.dot[0]; s.rho[1] = s.dot[1];
        if (it > 1) { const double2 b = zdiv(s.rho, s.rho1); s.beta[0] = b.x; s.beta[1] = b.y; }
        else { s.beta[0] = s.beta[1] = 0.0; }
        break;
    case KS_CG_END:                                 // :86-91
        if (!s.active) break;
        s.error = sqrt(s.dot[0]) / s.bnrm2;
        if (s.error <= tol) { s.active = 0; s.flag = 0; }
        else if (s.error != s.error) {                      // NaN: as above; `error > tol` is false for NaN, so cg reports flag 0 (:95)
            s.active = 0; s.flag = 0; s.pad = (it < max_it); s.iter = max_it;
        }
        else {
            s.rho1[0] = s.rho[0]; s.rho1[1] = s.rho[1];
            if (it == max_it) { s.active = 0; s.flag = 1; }
        }
        break;
    }
}

// Stable partition of the column indices: colmap[0 .. nact) = columns still iterating (ascending), colmap[nact .. ncols)
// = frozen columns (ascending); *nactive = nact.  One CTA; the tail matters because a launch sized with a stale (larger)
// count must still see every column at most once.
__global__ void __launch_bounds__(1024) k_compact(const KState* __restrict__ st, int ncols, int* __restrict__ colmap,
                                                  int* __restrict__ nactive) {
    __shared__ int wsum[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int pass = 0; pass < 2; ++pass) {                 // pass 0: active columns, pass 1: the rest
        if (pass == 0) { if (threadIdx.x == 0) base = 0; __syncthreads(); }
        for (int j0 = 0; j0 < ncols; j0 += 1024) {
            const int j = j0 + threadIdx.x;
            const bool sel = (j < ncols) && ((st[j].active != 0) == (pass == 0));
            const unsigned m = __ballot_sync(0xffffffffu, sel);
            if (lane == 0) wsum[w] = __popc(m);
            __syncthreads();
            int off = base;
            for (int i = 0; i < w; ++i) off += wsum[i];
            if (sel) colmap[off + __popc(m & ((1u << lane) - 1u))] = j;
            __syncthreads();
            if (threadIdx.x == 0) { int t = 0; for (int i = 0; i < 32; ++i) t += wsum[i]; base += t; }
            __syncthreads();
        }
        if (pass == 0 && threadIdx.x == 0) *nactive = base;
        __syncthreads();
    }
}

// ---- vector sweeps over a chunk: grid (x over n, y = logical column); frozen columns are skipped ------------------
#define TB_KCOL_LOOP(nn)                                                                                         \
    if (!st[j].active) return;                                                                                   \
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (nn); i += (long long)gridDim.x * blockDim.x)

// cgs :168-177
template <typename T>
__global__ void k_bcgs_dir(T* __restrict__ u, T* __restrict__ p, const T* __restrict__ r, const T* __restrict__ q,
                           const KState* __restrict__ st, int first, long long n, long long ld, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    const double2 beta = make_double2(st[j].beta[0], st[j].beta[1]);
    const long long o = j * ld;
    TB_KCOL_LOOP(n) {
        if (first) { const T rv = r[o + i]; u[o + i] = rv; p[o + i] = rv; }
        else {
            const T qv = q[o + i];
            const T uv = v_add(r[o + i], s_mul<T>(beta, qv));
            u[o + i] = uv;
            p[o + i] = v_add(uv, s_mul<T>(beta, v_add(qv, s_mul<T>(beta, p[o + i]))));
        }
    }
}
// cgs :185-188
template <typename T>
__global__ void k_bcgs_q(T* __restrict__ q, T* __restrict__ uh, const T* __restrict__ u, const T* __restrict__ vh,
                         const KState* __restrict__ st, long long n, long long ld, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    const double2 alpha = make_double2(st[j].alpha[0], st[j].alpha[1]);
    const long long o = j * ld;
    TB_KCOL_LOOP(n) {
        const T uv = u[o + i];
        const T qv = v_sub(uv, s_mul<T>(alpha, vh[o + i]));
        q[o + i] = qv;
        uh[o + i] = v_add(uv, qv);
    }
}
// cg :81-84:  x += alpha*d ; r -= alpha*Ad
template <typename T>
__global__ void k_bxr(T* __restrict__ x, long long ldx, T* __restrict__ r, const T* __restrict__ d, const T* __restrict__ Ad,
                      const KState* __restrict__ st, long long n, long long ld, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    const double2 alpha = make_double2(st[j].alpha[0], st[j].alpha[1]);
    const long long o = j * ld, ox = j * ldx;
    TB_KCOL_LOOP(n) {
        x[ox + i] = v_add(x[ox + i], s_mul<T>(alpha, d[o + i]));
        r[o + i] = v_sub(r[o + i], s_mul<T>(alpha, Ad[o + i]));
    }
}
// cgs :192-194:  x += alpha*d  (the residual update r -= alpha*A*uh, :196, is the epilogue of the product itself)
template <typename T>
__global__ void k_bx(T* __restrict__ x, long long ldx, const T* __restrict__ d, const KState* __restrict__ st, long long n,
                     long long ld, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    const double2 alpha = make_double2(st[j].alpha[0], st[j].alpha[1]);
    const long long o = j * ld, ox = j * ldx;
    TB_KCOL_LOOP(n) { x[ox + i] = v_add(x[ox + i], s_mul<T>(alpha, d[o + i])); }
}
// x_j *= scale_j (dir = 0, before the first residual) or iscale_j (dir = 1, at the end); every column, frozen or not
template <typename T>
__global__ void k_bscale(T* __restrict__ x, long long ldx, const KState* __restrict__ st, long long n, int dir) {
    const int j = blockIdx.y;
    const double f = dir ? st[j].iscale : st[j].scale;
    if (f == 1.0) return;
    const double2 f2 = make_double2(f, 0.0);
    const long long ox = j * ldx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[ox + i] = s_mul<T>(f2, x[ox + i]);
}
// r_j = scale_j * b_j
template <typename T>
__global__ void k_bcopy_scaled(T* __restrict__ r, long long ld, const T* __restrict__ b, long long ldb, const KState* __restrict__ st,
                               long long n) {
    const int j = blockIdx.y;
    const double2 f2 = make_double2(st[j].scale, 0.0);
    const long long o = j * ld, ob = j * ldb;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        r[o + i] = s_mul<T>(f2, b[ob + i]);
}
// columns frozen on a NaN residual before max_it: the reference's remaining iterations would have left x all NaN
template <typename T>
__global__ void k_bnanfill(T* __restrict__ x, long long ldx, const KState* __restrict__ st, long long n) {
    const int j = blockIdx.y;
    if (!st[j].pad) return;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const double2 nan2 = make_double2(qnan, qnan);
    const long long ox = j * ldx;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[ox + i] = s_mul<T>(nan2, x[ox + i]);
}
// cg :69-76
template <typename T>
__global__ void k_bcg_dir(T* __restrict__ p, const T* __restrict__ z, const KState* __restrict__ st, int first, long long n,
                          long long ld, const int* __restrict__ colmap) {
    const int j = colmap ? colmap[blockIdx.y] : blockIdx.y;
    const double2 beta = make_double2(st[j].beta[0], st[j].beta[1]);
    const long long o = j * ld;
    TB_KCOL_LOOP(n) { p[o + i] = first ? z[o + i] : v_add(z[o + i], s_mul<T>(beta, p[o + i])); }
}

}  // namespace tb
