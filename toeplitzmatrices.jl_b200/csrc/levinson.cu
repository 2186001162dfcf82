// Batched Levinson / Durbin recursions (K5): dispatch, the one-warp-per-system instantiations, the shared-memory
// fallback for n > 2048 and the one-matrix / many-right-hand-sides kernels.  Shared pieces: levinson_kernels.cuh.
#include "levinson_kernels.cuh"

namespace tb {

int g_lev_gen = 1;


// Fallback for n > 512: one CTA per system, iterates in shared memory (double-buffered y).
template <typename R, bool DURBIN>
__global__ void __launch_bounds__(256)
k_levinson_block(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm,
                 int64_t ldb, R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = (int)n64;
    const int tid = threadIdx.x, nth = blockDim.x;
    R* r_s = reinterpret_cast<R*>(smraw);
    R* b_s = r_s + n;
    R* x_s = b_s + n;
    R* y0 = x_s + n;
    R* y1 = y0 + n;
    __shared__ R red[2][8];
    __shared__ R sc[2];
    for (int64_t sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
        const int nr = DURBIN ? n : n - 1;
        const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
        const bool scale = (dg != (R)1);
        for (int i = tid; i < n; i += nth) {
            r_s[i] = (i < nr) ? (scale ? Rm[sys * ldr + i] / dg : Rm[sys * ldr + i]) : (R)0;
            b_s[i] = DURBIN ? (R)0 : Bm[sys * ldb + i];
            x_s[i] = 0; y0[i] = 0; y1[i] = 0;
        }
        __syncthreads();
        if (tid == 0) { y0[0] = -r_s[0]; x_s[0] = b_s[0]; }
        R alpha = -r_s[0], beta = 1;
        R* yc = y0; R* yn = y1;
        __syncthreads();
        for (int k = 1; k < n; ++k) {
            beta *= ((R)1 - alpha * alpha);
            R d1 = 0, d2 = 0;
            for (int i = tid; i < k; i += nth) {
                const R rr = r_s[k - 1 - i];
                if (!DURBIN) d1 = fma(rr, x_s[i], d1);
                d2 = fma(rr, yc[i], d2);
            }
            d1 = warp_sum(d1); d2 = warp_sum(d2);
            if ((tid & 31) == 0) { red[0][tid >> 5] = d1; red[1][tid >> 5] = d2; }
            __syncthreads();
            if (tid == 0) {
                R s1 = 0, s2 = 0;
                for (int w = 0; w < (nth >> 5); ++w) { s1 += red[0][w]; s2 += red[1][w]; }
                sc[0] = s1; sc[1] = s2;
            }
            __syncthreads();
            d1 = sc[0]; d2 = sc[1];
            if (!DURBIN) {
                const R mu = (b_s[k] - d1) / beta;
                for (int i = tid; i < k; i += nth) x_s[i] = fma(mu, yc[k - 1 - i], x_s[i]);
                if (tid == 0) x_s[k] = mu;
            }
            if (DURBIN || k < n - 1) {
                alpha = -(r_s[k] + d2) / beta;
                for (int i = tid; i < k; i += nth) yn[i] = fma(alpha, yc[k - 1 - i], yc[i]);
                if (tid == 0) yn[k] = alpha;
                R* t = yc; yc = yn; yn = t;
            }
            __syncthreads();
        }
        for (int i = tid; i < n; i += nth) {
            R v = DURBIN ? yc[i] : x_s[i];
            if (scale) v /= dg;
            Xm[sys * ldx + i] = v;
        }
        __syncthreads();
    }
}

// Any n: one CTA per system, generator form (levinson_gen.cu) with the vectors in shared memory (GLOBAL = false) or in
// a global-memory workspace that lives in L2 (GLOBAL = true, n beyond ~6000).  The shifting vector B costs nothing
// here: B_k[s] = buf[n - k + s] in a buffer of 2n slots, so "shift up by one and insert alpha" is a moving base.
// One barrier per step: the thread that owns slot k+1 publishes the next pivots (A[k+1], X[k+1]) in a two-slot
// mailbox while it updates them.
template <typename R, bool DURBIN, bool GLOBAL>
__global__ void __launch_bounds__(512)
k_levinson_gen_block(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm,
                     int64_t ldb, R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag, R* __restrict__ work) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ R piv[2][2];
    const int n = (int)n64;
    const int tid = threadIdx.x, nth = blockDim.x;
    const int nr = DURBIN ? n : n - 1;
    R* base = GLOBAL ? work + (size_t)blockIdx.x * (size_t)(4 * (size_t)n) : reinterpret_cast<R*>(smraw);
    R* A = base;                    // n
    R* Bb = A + n;                  // 2n
    R* X = Bb + 2 * (size_t)n;      // n (levinson only)
    for (int64_t sys = blockIdx.x; sys < nsys; sys += gridDim.x) {
        const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
        const bool scale = (dg != (R)1);
        for (int i = tid; i < n; i += nth) {
            R ri = 0, rim1 = (R)1;                     // r[i+1] and r[i] with r[0] = 1
            if (i < nr) { ri = Rm[sys * ldr + i]; if (scale) ri /= dg; }
            if (i >= 1) { rim1 = Rm[sys * ldr + i - 1]; if (scale) rim1 /= dg; }
            A[i] = ri;
            Bb[n + i] = rim1;
            Bb[i] = 0;
            if (!DURBIN) X[i] = -Bm[sys * ldb + i];
        }
        if (tid == 0) {
            R r1 = 0;
            if (nr > 0) { r1 = Rm[sys * ldr]; if (scale) r1 /= dg; }
            piv[0][0] = r1;                            // A[0]
            piv[0][1] = DURBIN ? (R)0 : -Bm[sys * ldb];
        }
        R alpha = 0, beta = 1;
        __syncthreads();
        for (int k = 0; k < n; ++k) {
            beta *= ((R)1 - alpha * alpha);
            const R ibeta = (R)1 / beta;
            const R ak = piv[k & 1][0], xk = piv[k & 1][1];
            const bool need_alpha = DURBIN || k < n - 1;
            alpha = need_alpha ? -ak * ibeta : (R)0;
            const R mu = -xk * ibeta;
            R* B = Bb + (n - k);                       // B_k[s] = B[s]
            for (int sidx = tid; sidx < n; sidx += nth) {
                const R bo = B[sidx];
                R xn = 0, an = 0;
                if (!DURBIN) { xn = (sidx == k) ? mu : fma(mu, bo, X[sidx]); X[sidx] = xn; }
                if (need_alpha) {
                    const R ao = A[sidx];
                    an = (sidx == k) ? alpha : fma(alpha, bo, ao);
                    A[sidx] = an;
                    B[sidx] = fma(alpha, ao, bo);
                }
                if (sidx == k + 1) { piv[(k + 1) & 1][0] = an; piv[(k + 1) & 1][1] = xn; }
            }
            if (need_alpha && tid == 0) Bb[n - k - 1] = alpha;     // head of B_{k+1}
            __syncthreads();
        }
        for (int i = tid; i < n; i += nth) {
            R v = DURBIN ? A[i] : X[i];
            if (scale) v /= dg;
            Xm[sys * ldx + i] = v;
        }
        __syncthreads();
    }
}

template <typename R, bool DURBIN>
static int launch_levinson_gen_block(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                     const R* diag, cudaStream_t s) {
    const size_t per_sys = (size_t)4 * n * sizeof(R);
    if (per_sys <= 200 * 1024) {
        auto kern = k_levinson_gen_block<R, DURBIN, false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)per_sys);
        if (e != cudaSuccess) return cuda_status(e, "levinson smem attribute");
        const unsigned grid = (unsigned)(nsys < 148 * 4 ? nsys : 148 * 4);
        kern<<<grid, 512, per_sys, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag, nullptr);
        count_launch();
        return cuda_status(cudaGetLastError(), "levinson generator-form block kernel");
    }
    const unsigned grid = (unsigned)(nsys < 148 * 2 ? nsys : 148 * 2);
    R* work = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&work), per_sys * grid, s);
    if (e != cudaSuccess) return cuda_status(e, "levinson workspace");
    k_levinson_gen_block<R, DURBIN, true><<<grid, 512, 0, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag, work);
    count_launch();
    e = cudaGetLastError();
    cudaFreeAsync(work, s);
    return cuda_status(e, "levinson generator-form block kernel (global workspace)");
}

template <typename R, bool DURBIN>
static int launch_levinson(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x,
                           int64_t ldx, const R* diag, cudaStream_t s) {
    if (nsys <= 0 || n <= 0) return TB200_OK;
    if (n <= 4 * LSPAN && g_lev_gen) {         // generator form: no reductions (levinson_gen.cu)
        const cudaError_t e = launch_levinson_gen<R, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        count_launch();
        return cuda_status(e, "levinson generator-form kernel");
    }
    if (n <= 4 * 16 * LB) {                   // small orders: 8 (n <= 64), 4 (n <= 128) or 2 (n <= 256) systems per warp
        const cudaError_t e = launch_levinson_small<R, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        count_launch();
        return cuda_status(e, "levinson warp kernel (several systems per warp)");
    }
    if (n <= 4 * LSPAN) {
        const int S = (int)((n + LSPAN - 1) / LSPAN);
        const int wpc = 4;
        const unsigned grid = (unsigned)((nsys + wpc - 1) / wpc);
        const size_t smem = (size_t)wpc * 2 * S * LSPAN * sizeof(R);
#define TB_LEV_CASE(SV)                                                                          \
    case SV:                                                                                     \
        k_levinson_warp<R, SV, DURBIN><<<grid, wpc * 32, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag); \
        break;
        switch (S) {
            TB_LEV_CASE(1) TB_LEV_CASE(2) TB_LEV_CASE(3) TB_LEV_CASE(4)
        }
#undef TB_LEV_CASE
        count_launch();
        return cuda_status(cudaGetLastError(), "levinson warp kernel");
    }
    if (n <= 16 * LSPAN) {                    // 2 (n <= 1024) or 4 (n <= 2048) warps per system
        const cudaError_t e = g_lev_gen ? launch_levinson_gen_mw<R, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s)
                                        : launch_levinson_mw<R, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        count_launch();
        return cuda_status(e, "levinson multi-warp kernel");
    }
    if (n >= (1ll << 30)) return set_error(TB200_UNSUPPORTED, "levinson/durbin: n must be below 2^30");
    const size_t smem = (size_t)5 * n * sizeof(R);
    if (g_lev_gen || smem > 200 * 1024)        // any n: generator form, shared memory or an L2-resident global workspace
        return launch_levinson_gen_block<R, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    auto kern = k_levinson_block<R, DURBIN>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e, "levinson smem attribute");
    const unsigned grid = (unsigned)(nsys < 148 * 8 ? nsys : 148 * 8);
    kern<<<grid, 256, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);
    count_launch();
    return cuda_status(cudaGetLastError(), "levinson block kernel");
}


// ---------------------------------------------------------------------------------------------------------
// Many right-hand sides, ONE matrix (the shape of ext/ToeplitzMatricesStatsBaseExt.jl:15-24): the y-recursion
// (Durbin part of levinson!, src/directLinearSolvers.jl:114-118) does not depend on b, so it runs once and
// tabulates, for every order k, 1/beta_k and reverse(y_k); each right-hand side then only needs
//   mu = (b[k+1] - reverse_dot(r_k, x_k)) / beta ;  x_k += mu*reverse(y_k) ;  x[k+1] = mu      (:111-113)
// i.e. 2k instead of 5k FMAs per step, one warp per right-hand side, x in registers as above.
// ---------------------------------------------------------------------------------------------------------
template <typename R, int S, int SC, int J>
__device__ __forceinline__ void ytab_step(LevState<R, S>& st, int k, int n, int lane, int lb, const R* r_s,
                                          R* __restrict__ ytab, R* __restrict__ ibeta_tab) {
    constexpr int ACT = (SC + 1) * LB;
    constexpr int PH = (J + LB - 1) & (LB - 1);
    st.beta *= ((R)1 - st.alpha * st.alpha);
    const R ibeta = fast_rcp(st.beta);
    if (lane == 0) ibeta_tab[k] = ibeta;
    // reverse(y_k) in natural order, zero padded to the active super-rows, at row k of a (n x S*128) table:
    // rows are 16-byte aligned so the consumers read them with unconditional 128-bit loads
    R* row = ytab + (long long)k * (S * LSPAN);
#pragma unroll
    for (int i = 0; i < ACT; ++i) row[(i / LB) * LSPAN + lane * LB + (i % LB)] = st.yh[phys(i, PH)];
    if (k < n - 1) {
        R p2[LB];
#pragma unroll
        for (int q = 0; q < LB; ++q) p2[q] = 0;
#pragma unroll
        for (int i = 0; i < ACT; ++i) p2[i % LB] = fma(st.rh[phys(i, PH)], st.y[i], p2[i % LB]);
        const R d2 = warp_sum((p2[0] + p2[1]) + (p2[2] + p2[3]));
        const R rk = r_s[k];
        st.alpha = -(rk + d2) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) {
            const R yo = st.y[i];
            st.y[i] = fma(st.alpha, st.yh[phys(i, PH)], yo);
            st.yh[phys(i, PH)] = fma(st.alpha, yo, st.yh[phys(i, PH)]);
        }
        if (lane == lb) st.y[SC * LB + J] = st.alpha;
        shift_up<R, SC + 1, PH>(st.yh, st.alpha, lane);
        shift_up<R, SC + 1, PH>(st.rh, rk, lane);
    }
}

template <typename R, int S, int SC>
__device__ __forceinline__ void ytab_superrows(LevState<R, S>& st, int n, int lane, const R* r_s, R* ytab, R* ibeta_tab) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < 32; ++lb) {
            const int k0 = SC * LSPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 + 0 >= 1 && k0 + 0 < n) ytab_step<R, S, SC, 0>(st, k0 + 0, n, lane, lb, r_s, ytab, ibeta_tab);
            if (k0 + 1 < n) ytab_step<R, S, SC, 1>(st, k0 + 1, n, lane, lb, r_s, ytab, ibeta_tab);
            if (k0 + 2 < n) ytab_step<R, S, SC, 2>(st, k0 + 2, n, lane, lb, r_s, ytab, ibeta_tab);
            if (k0 + 3 < n) ytab_step<R, S, SC, 3>(st, k0 + 3, n, lane, lb, r_s, ytab, ibeta_tab);
        }
        ytab_superrows<R, S, SC + 1>(st, n, lane, r_s, ytab, ibeta_tab);
    }
}

// one warp: the shared y-recursion of T = SymToeplitz(vc) (n x n); rnorm receives r = vc[1:] / vc[0]
template <typename R, int S>
__global__ void __launch_bounds__(32)
k_levinson_ytab(int n, const R* __restrict__ vc, R* __restrict__ rnorm, R* __restrict__ ytab, R* __restrict__ ibeta_tab) {
    __shared__ R r_s[S * LSPAN];
    const int lane = threadIdx.x;
    const R d = vc[0];
    for (int i = lane; i < n - 1; i += 32) { const R v = (d != (R)1) ? vc[i + 1] / d : vc[i + 1]; r_s[i] = v; rnorm[i] = v; }
    __syncwarp();
    LevState<R, S> st;
#pragma unroll
    for (int i = 0; i < S * LB; ++i) { st.x[i] = 0; st.y[i] = 0; st.yh[i] = 0; st.rh[i] = 0; }
    const R r0 = r_s[0];
    if (lane == 0) { st.y[0] = -r0; st.yh[0] = -r0; st.rh[0] = r0; }
    st.alpha = -r0;
    st.beta = (R)1;
    ytab_superrows<R, S, 0>(st, n, lane, r_s, ytab, ibeta_tab);
}

// two right-hand sides per warp: the tabulated reverse(y_k) and the shifting r-window are loaded / kept once and
// used twice, and the two dot products share one fused two-value reduction
template <typename R, int S, int SC, int J>
__device__ __forceinline__ void rhs_step(R* xa, R* xb, R* rh, int k, int n, int lane, int lb, const R* r_s, const R* ba_s,
                                         const R* bb_s, const R* __restrict__ ytab, const R* __restrict__ ibeta_tab) {
    constexpr int ACT = (SC + 1) * LB;
    constexpr int PH = (J + LB - 1) & (LB - 1);
    // reverse(y_k) for this lane's slots: independent of the reduction, so the loads are in flight during it
    const R* row = ytab + (long long)k * (S * LSPAN) + lane * LB;
    R yh[ACT];
#pragma unroll
    for (int sr = 0; sr <= SC; ++sr) {
        if constexpr (sizeof(R) == 8) {
            const double2 v0 = __ldg(reinterpret_cast<const double2*>(row + sr * LSPAN));
            const double2 v1 = __ldg(reinterpret_cast<const double2*>(row + sr * LSPAN) + 1);
            yh[sr * LB + 0] = v0.x; yh[sr * LB + 1] = v0.y; yh[sr * LB + 2] = v1.x; yh[sr * LB + 3] = v1.y;
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + sr * LSPAN));
            yh[sr * LB + 0] = v.x; yh[sr * LB + 1] = v.y; yh[sr * LB + 2] = v.z; yh[sr * LB + 3] = v.w;
        }
    }
    const R ibeta = __ldg(ibeta_tab + k);
    R pa[LB], pb[LB];
#pragma unroll
    for (int q = 0; q < LB; ++q) { pa[q] = 0; pb[q] = 0; }
#pragma unroll
    for (int i = 0; i < ACT; ++i) {
        pa[i % LB] = fma(rh[phys(i, PH)], xa[i], pa[i % LB]);
        pb[i % LB] = fma(rh[phys(i, PH)], xb[i], pb[i % LB]);
    }
    R da = (pa[0] + pa[1]) + (pa[2] + pa[3]);
    R db = (pb[0] + pb[1]) + (pb[2] + pb[3]);
    warp_sum2(da, db, lane);
    const R mua = (ba_s[k] - da) * ibeta;
    const R mub = (bb_s[k] - db) * ibeta;
#pragma unroll
    for (int i = 0; i < ACT; ++i) { xa[i] = fma(mua, yh[i], xa[i]); xb[i] = fma(mub, yh[i], xb[i]); }
    if (lane == lb) { xa[SC * LB + J] = mua; xb[SC * LB + J] = mub; }
    if (k < n - 1) shift_up<R, SC + 1, PH>(rh, r_s[k], lane);
}

template <typename R, int S, int SC>
__device__ __forceinline__ void rhs_superrows(R* xa, R* xb, R* rh, int n, int lane, const R* r_s, const R* ba_s, const R* bb_s,
                                              const R* ytab, const R* ibeta_tab) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < 32; ++lb) {
            const int k0 = SC * LSPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 + 0 >= 1 && k0 + 0 < n) rhs_step<R, S, SC, 0>(xa, xb, rh, k0 + 0, n, lane, lb, r_s, ba_s, bb_s, ytab, ibeta_tab);
            if (k0 + 1 < n) rhs_step<R, S, SC, 1>(xa, xb, rh, k0 + 1, n, lane, lb, r_s, ba_s, bb_s, ytab, ibeta_tab);
            if (k0 + 2 < n) rhs_step<R, S, SC, 2>(xa, xb, rh, k0 + 2, n, lane, lb, r_s, ba_s, bb_s, ytab, ibeta_tab);
            if (k0 + 3 < n) rhs_step<R, S, SC, 3>(xa, xb, rh, k0 + 3, n, lane, lb, r_s, ba_s, bb_s, ytab, ibeta_tab);
        }
        rhs_superrows<R, S, SC + 1>(xa, xb, rh, n, lane, r_s, ba_s, bb_s, ytab, ibeta_tab);
    }
}

// one warp per PAIR of right-hand side columns; X = T \ B with the tabulated y-recursion; diag = vc[0]
template <typename R, int S>
__global__ void __launch_bounds__(128, 3)
k_levinson_rhs(int n, int64_t nrhs, const R* __restrict__ rnorm, const R* __restrict__ vc, const R* __restrict__ Bm, int64_t ldb,
               R* __restrict__ Xm, int64_t ldx, const R* __restrict__ ytab, const R* __restrict__ ibeta_tab) {
    extern __shared__ __align__(16) unsigned char smraw[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int64_t cola = 2 * ((int64_t)blockIdx.x * (blockDim.x >> 5) + wic);
    if (cola >= nrhs) return;
    const bool two = cola + 1 < nrhs;
    const int64_t colb = two ? cola + 1 : cola;
    R* r_s = reinterpret_cast<R*>(smraw) + (size_t)wic * (3 * S * LSPAN);
    R* ba_s = r_s + S * LSPAN;
    R* bb_s = ba_s + S * LSPAN;
    for (int i = lane; i < n - 1; i += 32) r_s[i] = rnorm[i];
    for (int i = lane; i < n; i += 32) { ba_s[i] = Bm[cola * ldb + i]; bb_s[i] = Bm[colb * ldb + i]; }
    __syncwarp();
    R xa[S * LB], xb[S * LB], rh[S * LB];
#pragma unroll
    for (int i = 0; i < S * LB; ++i) { xa[i] = 0; xb[i] = 0; rh[i] = 0; }
    if (lane == 0) { xa[0] = ba_s[0]; xb[0] = bb_s[0]; rh[0] = r_s[0]; }
    rhs_superrows<R, S, 0>(xa, xb, rh, n, lane, r_s, ba_s, bb_s, ytab, ibeta_tab);
    const R d = vc[0];
    R* outa = Xm + cola * ldx;
    R* outb = Xm + colb * ldx;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = s * LSPAN + lane * LB + j;
            if (i < n) {
                outa[i] = (d != (R)1) ? xa[s * LB + j] / d : xa[s * LB + j];
                if (two) outb[i] = (d != (R)1) ? xb[s * LB + j] / d : xb[s * LB + j];
            }
        }
}

// ---------------------------------------------------------------------------------------------------------
// The same shape in generator form (levinson_gen.cu): the shared part is the Durbin recursion on (A, B); row k of
// the table holds the WHOLE shifting vector B_k (reverse(y_k) in slots < k, the generator v_k in slots >= k) and
// 1/beta_k.  A right-hand side then needs no dot product at all:  mu = -X[k]/beta_k ;  X += mu*B_k ;  X[k] = mu
// with X = x_k in slots < k and -(b - T x_k) in slots >= k -- n FMAs per step and column, one pivot broadcast, no
// reduction; NC columns share every table row a warp loads.
// ---------------------------------------------------------------------------------------------------------
template <typename R, int S, int SC, int J>
__device__ __forceinline__ void btab_step(R* a, R* b, R& alpha, R& beta, int k, int n, int lane, int lb, R* __restrict__ btab,
                                          R* __restrict__ ibeta_tab) {
    constexpr int N = S * LB;
    constexpr int PH = J;
    beta *= ((R)1 - alpha * alpha);
    const R ibeta = fast_rcp2(beta);
    if (lane == 0) ibeta_tab[k] = ibeta;
    R* row = btab + (long long)k * (S * LSPAN);
#pragma unroll
    for (int i = 0; i < N; ++i) row[(i / LB) * LSPAN + lane * LB + (i % LB)] = b[phys(i, PH)];
    if (k < n - 1) {
        const R ak = __shfl_sync(0xffffffffu, a[SC * LB + J], lb);
        alpha = -ak * ibeta;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const R ao = a[i], bo = b[phys(i, PH)];
            a[i] = fma(alpha, bo, ao);
            b[phys(i, PH)] = fma(alpha, ao, bo);
        }
        if (lane == lb) a[SC * LB + J] = alpha;
        shift_up<R, S, PH>(b, alpha, lane);
    }
}
template <typename R, int S, int SC>
__device__ __forceinline__ void btab_superrows(R* a, R* b, R& alpha, R& beta, int n, int lane, R* btab, R* ibeta_tab) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < 32; ++lb) {
            const int k0 = SC * LSPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 + 0 < n) btab_step<R, S, SC, 0>(a, b, alpha, beta, k0 + 0, n, lane, lb, btab, ibeta_tab);
            if (k0 + 1 < n) btab_step<R, S, SC, 1>(a, b, alpha, beta, k0 + 1, n, lane, lb, btab, ibeta_tab);
            if (k0 + 2 < n) btab_step<R, S, SC, 2>(a, b, alpha, beta, k0 + 2, n, lane, lb, btab, ibeta_tab);
            if (k0 + 3 < n) btab_step<R, S, SC, 3>(a, b, alpha, beta, k0 + 3, n, lane, lb, btab, ibeta_tab);
        }
        btab_superrows<R, S, SC + 1>(a, b, alpha, beta, n, lane, btab, ibeta_tab);
    }
}
// one warp: the shared recursion of T = SymToeplitz(vc) (n x n), rows B_k and 1/beta_k for k = 0 .. n-1
template <typename R, int S>
__global__ void __launch_bounds__(32)
k_levinson_btab(int n, const R* __restrict__ vc, R* __restrict__ btab, R* __restrict__ ibeta_tab) {
    __shared__ R r_s[S * LSPAN + LB];
    const int lane = threadIdx.x;
    const R d = vc[0];
    if (lane == 0) r_s[0] = (R)1;
    for (int i = lane; i < S * LSPAN; i += 32) r_s[1 + i] = (i < n - 1) ? ((d != (R)1) ? vc[i + 1] / d : vc[i + 1]) : (R)0;
    __syncwarp();
    R a[S * LB], b[S * LB];
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int sl = sr * LSPAN + lane * LB + j;
            a[sr * LB + j] = r_s[1 + sl];
            b[sr * LB + j] = (sl < n) ? r_s[sl] : (R)0;
        }
    R alpha = 0, beta = 1;
    btab_superrows<R, S, 0>(a, b, alpha, beta, n, lane, btab, ibeta_tab);
}

template <typename R, int S, int NC, int SC, int J>
__device__ __forceinline__ void rhsg_step(R (*x)[S * LB], int k, int lane, int lb, const R* __restrict__ btab,
                                          const R* __restrict__ ibeta_tab) {
    constexpr int N = S * LB;
    const R* row = btab + (long long)k * (S * LSPAN) + lane * LB;
    R bk[N];
#pragma unroll
    for (int sr = 0; sr < S; ++sr) {
        if constexpr (sizeof(R) == 8) {
            const double2 v0 = __ldg(reinterpret_cast<const double2*>(row + sr * LSPAN));
            const double2 v1 = __ldg(reinterpret_cast<const double2*>(row + sr * LSPAN) + 1);
            bk[sr * LB + 0] = v0.x; bk[sr * LB + 1] = v0.y; bk[sr * LB + 2] = v1.x; bk[sr * LB + 3] = v1.y;
        } else {
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + sr * LSPAN));
            bk[sr * LB + 0] = v.x; bk[sr * LB + 1] = v.y; bk[sr * LB + 2] = v.z; bk[sr * LB + 3] = v.w;
        }
    }
    const R ibeta = __ldg(ibeta_tab + k);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const R xk = __shfl_sync(0xffffffffu, x[c][SC * LB + J], lb);
        const R mu = -xk * ibeta;
#pragma unroll
        for (int i = 0; i < N; ++i) x[c][i] = fma(mu, bk[i], x[c][i]);
        if (lane == lb) x[c][SC * LB + J] = mu;
    }
}
template <typename R, int S, int NC, int SC>
__device__ __forceinline__ void rhsg_superrows(R (*x)[S * LB], int n, int lane, const R* btab, const R* ibeta_tab) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < 32; ++lb) {
            const int k0 = SC * LSPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 + LB <= n) {
                rhsg_step<R, S, NC, SC, 0>(x, k0 + 0, lane, lb, btab, ibeta_tab);
                rhsg_step<R, S, NC, SC, 1>(x, k0 + 1, lane, lb, btab, ibeta_tab);
                rhsg_step<R, S, NC, SC, 2>(x, k0 + 2, lane, lb, btab, ibeta_tab);
                rhsg_step<R, S, NC, SC, 3>(x, k0 + 3, lane, lb, btab, ibeta_tab);
                continue;
            }
            if (k0 + 0 < n) rhsg_step<R, S, NC, SC, 0>(x, k0 + 0, lane, lb, btab, ibeta_tab);
            if (k0 + 1 < n) rhsg_step<R, S, NC, SC, 1>(x, k0 + 1, lane, lb, btab, ibeta_tab);
            if (k0 + 2 < n) rhsg_step<R, S, NC, SC, 2>(x, k0 + 2, lane, lb, btab, ibeta_tab);
            if (k0 + 3 < n) rhsg_step<R, S, NC, SC, 3>(x, k0 + 3, lane, lb, btab, ibeta_tab);
        }
        rhsg_superrows<R, S, NC, SC + 1>(x, n, lane, btab, ibeta_tab);
    }
}
// one warp per NC right-hand-side columns; X = T \ B; diag = vc[0]
template <typename R, int S, int NC>
__global__ void __launch_bounds__(128, (sizeof(R) == 8 && NC * S > 8) ? 2 : 3)
k_levinson_rhs_gen(int n, int64_t nrhs, const R* __restrict__ vc, const R* __restrict__ Bm, int64_t ldb, R* __restrict__ Xm, int64_t ldx,
                   const R* __restrict__ btab, const R* __restrict__ ibeta_tab) {
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int64_t col0 = NC * ((int64_t)blockIdx.x * (blockDim.x >> 5) + wic);
    if (col0 >= nrhs) return;
    R x[NC][S * LB];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int64_t col = (col0 + c < nrhs) ? col0 + c : nrhs - 1;        // surplus columns shadow the last one, never stored
        const R* bc = Bm + col * ldb;
#pragma unroll
        for (int sr = 0; sr < S; ++sr)
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int sl = sr * LSPAN + lane * LB + j;
                x[c][sr * LB + j] = (sl < n) ? -bc[sl] : (R)0;
            }
    }
    rhsg_superrows<R, S, NC, 0>(x, n, lane, btab, ibeta_tab);
    const R d = vc[0];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (col0 + c >= nrhs) break;
        R* out = Xm + (col0 + c) * ldx;
#pragma unroll
        for (int sr = 0; sr < S; ++sr)
#pragma unroll
            for (int j = 0; j < LB; ++j) {
                const int i = sr * LSPAN + lane * LB + j;
                if (i < n) out[i] = (d != (R)1) ? x[c][sr * LB + j] / d : x[c][sr * LB + j];
            }
    }
}

template <typename R>
static int launch_levinson_multirhs_gen(int64_t n, int64_t nrhs, const R* vc, const R* B, int64_t ldb, R* X, int64_t ldx,
                                        R* work /* n*S*128 + n */, cudaStream_t s) {
    const int S = (int)((n + LSPAN - 1) / LSPAN);
    R* btab = work;                                   // first: keeps the table rows 16-byte aligned
    R* ibeta_tab = btab + n * (int64_t)(S * LSPAN);
#ifndef TB_MRHS_NC
#define TB_MRHS_NC 4
#endif
    constexpr int NC = TB_MRHS_NC;                    // right-hand sides per warp (they share every table row the warp loads)
    const int wpc = 4;
    const int64_t ngroups = (nrhs + NC - 1) / NC;
    const unsigned grid = (unsigned)((ngroups + wpc - 1) / wpc);
#define TB_MRG_CASE(SV)                                                                                               \
    case SV:                                                                                                          \
        k_levinson_btab<R, SV><<<1, 32, 0, s>>>((int)n, vc, btab, ibeta_tab);                                        \
        k_levinson_rhs_gen<R, SV, NC><<<grid, wpc * 32, 0, s>>>((int)n, nrhs, vc, B, ldb, X, ldx, btab, ibeta_tab);  \
        break;
    switch (S) { TB_MRG_CASE(1) TB_MRG_CASE(2) TB_MRG_CASE(3) TB_MRG_CASE(4) }
#undef TB_MRG_CASE
    count_launch(2);
    return cuda_status(cudaGetLastError(), "multi-RHS levinson kernels (generator form)");
}

template <typename R>
static int launch_levinson_multirhs(int64_t n, int64_t nrhs, const R* vc, const R* B, int64_t ldb, R* X, int64_t ldx,
                                    R* work /* n*S*128 + n + n */, cudaStream_t s) {
    const int S = (int)((n + LSPAN - 1) / LSPAN);
    R* ytab = work;                                   // first: keeps the table rows 16-byte aligned
    R* ibeta_tab = ytab + n * (int64_t)(S * LSPAN);
    R* rnorm = ibeta_tab + n;
    const int wpc = 4;
    const int64_t npairs = (nrhs + 1) / 2;
    const unsigned grid = (unsigned)((npairs + wpc - 1) / wpc);
    const size_t smem = (size_t)wpc * 3 * S * LSPAN * sizeof(R);
#define TB_MR_CASE(SV)                                                                                                   \
    case SV:                                                                                                             \
        k_levinson_ytab<R, SV><<<1, 32, 0, s>>>((int)n, vc, rnorm, ytab, ibeta_tab);                                    \
        k_levinson_rhs<R, SV><<<grid, wpc * 32, smem, s>>>((int)n, nrhs, rnorm, vc, B, ldb, X, ldx, ytab, ibeta_tab);  \
        break;
    switch (S) { TB_MR_CASE(1) TB_MR_CASE(2) TB_MR_CASE(3) TB_MR_CASE(4) }
#undef TB_MR_CASE
    count_launch(2);
    return cuda_status(cudaGetLastError(), "multi-RHS levinson kernels");
}

int levinson_multirhs_dispatch(int dtype, int64_t n, int64_t nrhs, const void* vc, const void* B, int64_t ldb, void* X,
                               int64_t ldx, void* work, cudaStream_t s) {
    if (n < 2 || n > 4 * LSPAN) return set_error(TB200_UNSUPPORTED, "multi-RHS levinson: 2 <= n <= 512 (use tb200_levinson_batched)");
    if (g_lev_gen) {
        if (dtype == TB200_F64) return launch_levinson_multirhs_gen<double>(n, nrhs, (const double*)vc, (const double*)B, ldb, (double*)X, ldx, (double*)work, s);
        if (dtype == TB200_F32) return launch_levinson_multirhs_gen<float>(n, nrhs, (const float*)vc, (const float*)B, ldb, (float*)X, ldx, (float*)work, s);
    }
    if (dtype == TB200_F64) return launch_levinson_multirhs<double>(n, nrhs, (const double*)vc, (const double*)B, ldb, (double*)X, ldx, (double*)work, s);
    if (dtype == TB200_F32) return launch_levinson_multirhs<float>(n, nrhs, (const float*)vc, (const float*)B, ldb, (float*)X, ldx, (float*)work, s);
    return set_error(TB200_UNSUPPORTED, "levinson: only real Float32/Float64 systems");
}

int levinson_dispatch(int dtype, bool durbin, int64_t n, int64_t nsys, const void* r, int64_t ldr, const void* b,
                      int64_t ldb, void* x, int64_t ldx, const void* diag, cudaStream_t s) {
    if (dtype == TB200_F64) {
        return durbin ? launch_levinson<double, true>(n, nsys, (const double*)r, ldr, nullptr, 0, (double*)x, ldx, nullptr, s)
                      : launch_levinson<double, false>(n, nsys, (const double*)r, ldr, (const double*)b, ldb, (double*)x, ldx, (const double*)diag, s);
    }
    if (dtype == TB200_F32) {
        return durbin ? launch_levinson<float, true>(n, nsys, (const float*)r, ldr, nullptr, 0, (float*)x, ldx, nullptr, s)
                      : launch_levinson<float, false>(n, nsys, (const float*)r, ldr, (const float*)b, ldb, (float*)x, ldx, (const float*)diag, s);
    }
    return set_error(TB200_UNSUPPORTED, "levinson/durbin: only real Float32/Float64 systems");
}

}  // namespace tb
