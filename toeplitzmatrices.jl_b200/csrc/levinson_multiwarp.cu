// Levinson / Durbin for 512 < n <= 2048: 2 or 4 warps per system (levinson_kernels.cuh for the shared pieces).
#include "levinson_kernels.cuh"

namespace tb {


// ---------------------------------------------------------------------------------------------------------
// 512 < n <= 2048: W = 2 or 4 warps per system, same register layout stretched over 32*W lanes
// (element i: lane (i/4) % (32 W), register (i / (128 W))*4 + i%4).  Per step the warps meet ONCE, at a named
// barrier, to exchange (a) their partial dot products and (b) the element of the reversed-y vector that crosses
// each warp boundary: the shift of yh is deferred to the next step, after that step's reduction barrier, where
// it is still ahead of yh's next use; the shift of rh needs no exchange because rh is the reversed r sequence,
// whose boundary elements lane 0 reads straight from shared memory.  Exchange slots are double-buffered by step
// parity (a warp can run at most one step ahead of the slowest warp of its system).
// ---------------------------------------------------------------------------------------------------------
template <typename R, int S, int W, int SC, int J, bool DURBIN, bool NEWROW>
__device__ __forceinline__ void mw_step(LevState<R, S>& st, int k, int n, int lane, int wis, int lb, const R* r_s, const R* b_s,
                                        R* xch, int bar_id) {
    constexpr int SPAN = LSPAN * W;
    constexpr int ACT = (SC + 1) * LB;
    constexpr int PH = (J + LB - 1) & (LB - 1);            // phase of rh (and of yh once its pending shift is applied)
    constexpr int PHP = (PH + LB - 1) & (LB - 1);          // phase of yh before the pending shift
    constexpr int WRAPP = (LB - 1 - PHP) & (LB - 1);
    constexpr int SROWS_PREV = NEWROW ? SC : SC + 1;       // super-rows the previous step's shift covers
    constexpr int XS = 2 * W + W * S;                      // doubles per parity buffer
    const int glane = wis * 32 + lane;
    st.beta *= ((R)1 - st.alpha * st.alpha);
    const R ibeta = fast_rcp(st.beta);
    R p1[LB], p2[LB];
#pragma unroll
    for (int q = 0; q < LB; ++q) { p1[q] = 0; p2[q] = 0; }
#pragma unroll
    for (int i = 0; i < ACT; ++i) {
        if (!DURBIN) p1[i % LB] = fma(st.rh[phys(i, PH)], st.x[i], p1[i % LB]);
        p2[i % LB] = fma(st.rh[phys(i, PH)], st.y[i], p2[i % LB]);
    }
    R d1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
    R d2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
    if (!DURBIN) warp_sum2(d1, d2, lane);
    else d2 = warp_sum(d2);
    R* slot = xch + (k & 1) * XS;
    if (lane == 0) { slot[wis * 2 + 0] = d1; slot[wis * 2 + 1] = d2; }
    const bool pending = k > 1;
    if (pending && lane == 31) {
#pragma unroll
        for (int sr = 0; sr < SROWS_PREV; ++sr) slot[2 * W + wis * S + sr] = st.yh[sr * LB + WRAPP];
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * W) : "memory");
    d1 = 0; d2 = 0;
#pragma unroll
    for (int w = 0; w < W; ++w) { d1 += slot[w * 2 + 0]; d2 += slot[w * 2 + 1]; }
    if (pending) {                                         // deferred shift of yh (head = the previous step's alpha)
#pragma unroll
        for (int sr = 0; sr < SROWS_PREV; ++sr) {
            const R rot = __shfl_sync(0xffffffffu, st.yh[sr * LB + WRAPP], (lane + 31) & 31);
            R in = rot;
            if (lane == 0) {
                if (wis > 0) in = slot[2 * W + (wis - 1) * S + sr];
                else in = (sr == 0) ? st.alpha : slot[2 * W + (W - 1) * S + (sr - 1)];
            }
            st.yh[sr * LB + WRAPP] = in;
        }
    }
    if (!DURBIN) {
        const R mu = (b_s[k] - d1) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) st.x[i] = fma(mu, st.yh[phys(i, PH)], st.x[i]);
        if (glane == lb) st.x[SC * LB + J] = mu;
    }
    if (DURBIN || k < n - 1) {
        const R rk = r_s[k];
        st.alpha = -(rk + d2) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) {
            const R yo = st.y[i];
            st.y[i] = fma(st.alpha, st.yh[phys(i, PH)], yo);
            st.yh[phys(i, PH)] = fma(st.alpha, yo, st.yh[phys(i, PH)]);
        }
        if (glane == lb) st.y[SC * LB + J] = st.alpha;
        // rh <- shift_up(rh): the element entering lane 0 of a warp is r[k - i0], i0 = first index the lane holds
        constexpr int WRAP = (LB - 1 - PH) & (LB - 1);
#pragma unroll
        for (int sr = 0; sr <= SC; ++sr) {
            const R rot = __shfl_sync(0xffffffffu, st.rh[sr * LB + WRAP], (lane + 31) & 31);
            const int src = k - (sr * SPAN + wis * LSPAN);
            st.rh[sr * LB + WRAP] = (lane == 0) ? (src >= 0 ? r_s[src] : (R)0) : rot;
        }
    }
}

template <typename R, int S, int W, int SC, bool DURBIN>
__device__ __forceinline__ void mw_superrows(LevState<R, S>& st, int n, int lane, int wis, const R* r_s, const R* b_s, R* xch,
                                             int bar_id) {
    if constexpr (SC < S) {
        constexpr int SPAN = LSPAN * W;
        for (int lb = 0; lb < 32 * W; ++lb) {
            const int k0 = SC * SPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 >= 1) {
                if (lb == 0) mw_step<R, S, W, SC, 0, DURBIN, (SC > 0)>(st, k0, n, lane, wis, lb, r_s, b_s, xch, bar_id);
                else mw_step<R, S, W, SC, 0, DURBIN, false>(st, k0, n, lane, wis, lb, r_s, b_s, xch, bar_id);
            }
            if (k0 + 1 < n) mw_step<R, S, W, SC, 1, DURBIN, false>(st, k0 + 1, n, lane, wis, lb, r_s, b_s, xch, bar_id);
            if (k0 + 2 < n) mw_step<R, S, W, SC, 2, DURBIN, false>(st, k0 + 2, n, lane, wis, lb, r_s, b_s, xch, bar_id);
            if (k0 + 3 < n) mw_step<R, S, W, SC, 3, DURBIN, false>(st, k0 + 3, n, lane, wis, lb, r_s, b_s, xch, bar_id);
        }
        mw_superrows<R, S, W, SC + 1, DURBIN>(st, n, lane, wis, r_s, b_s, xch, bar_id);
    }
}

// 128 threads: 4/W systems per CTA
template <typename R, int S, int W, bool DURBIN>
__global__ void __launch_bounds__(128, 3)
k_levinson_multiwarp(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm, int64_t ldb,
                     R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
    constexpr int SPAN = LSPAN * W;
    constexpr int NREG = S * LB;
    constexpr int XS = 2 * W + W * S;
    constexpr int PER_SYS = 2 * S * SPAN + 2 * XS;         // r, b, two exchange buffers
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = (int)n64;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sic = warp / W, wis = warp % W;              // system in CTA, warp in system
    const int64_t sys = (int64_t)blockIdx.x * (4 / W) + sic;
    if (sys >= nsys) return;                               // whole systems leave together: the named barrier stays consistent
    const int glane = wis * 32 + lane;
    const int bar_id = 1 + sic;
    const int nr = DURBIN ? n : n - 1;
    R* r_s = reinterpret_cast<R*>(smraw) + (size_t)sic * PER_SYS;
    R* b_s = r_s + S * SPAN;
    R* xch = b_s + S * SPAN;
    const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
    const bool scale = (dg != (R)1);
    for (int i = glane; i < nr; i += 32 * W) {
        const R v = Rm[sys * ldr + i];
        r_s[i] = scale ? v / dg : v;
    }
    if (!DURBIN)
        for (int i = glane; i < n; i += 32 * W) b_s[i] = Bm[sys * ldb + i];
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * W) : "memory");

    LevState<R, S> st;
#pragma unroll
    for (int i = 0; i < NREG; ++i) { st.x[i] = 0; st.y[i] = 0; st.yh[i] = 0; st.rh[i] = 0; }
    const R r0 = r_s[0];
    if (glane == 0) { st.y[0] = -r0; st.yh[0] = -r0; st.rh[0] = r0; if (!DURBIN) st.x[0] = b_s[0]; }
    st.alpha = -r0;
    st.beta = (R)1;
    mw_superrows<R, S, W, 0, DURBIN>(st, n, lane, wis, r_s, b_s, xch, bar_id);

    R* out = Xm + sys * ldx;
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = sr * SPAN + glane * LB + j;
            if (i < n) {
                R v = DURBIN ? st.y[sr * LB + j] : st.x[sr * LB + j];
                if (scale) v /= dg;
                out[i] = v;
            }
        }
}

template <typename R, int S, int W, bool DURBIN>
static cudaError_t launch_multiwarp(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                    const R* diag, cudaStream_t s) {
    constexpr int SPAN = LSPAN * W;
    constexpr int XS = 2 * W + W * S;
    constexpr size_t smem = (size_t)(4 / W) * (2 * S * SPAN + 2 * XS) * sizeof(R);
    auto kern = k_levinson_multiwarp<R, S, W, DURBIN>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int per_cta = 4 / W;
    const unsigned grid = (unsigned)((nsys + per_cta - 1) / per_cta);
    kern<<<grid, 128, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);
    return cudaGetLastError();
}

template <typename R, bool DURBIN>
cudaError_t launch_levinson_mw(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                               const R* diag, cudaStream_t s) {
    if (n <= 6 * LSPAN) return launch_multiwarp<R, 3, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    if (n <= 8 * LSPAN) return launch_multiwarp<R, 4, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    if (n <= 12 * LSPAN) return launch_multiwarp<R, 3, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    return launch_multiwarp<R, 4, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
}

#define TB_INST(R, D) template cudaError_t launch_levinson_mw<R, D>(int64_t, int64_t, const R*, int64_t, const R*, int64_t, R*, int64_t, const R*, cudaStream_t);
TB_INST(float, false) TB_INST(float, true) TB_INST(double, false) TB_INST(double, true)
#undef TB_INST

}  // namespace tb
