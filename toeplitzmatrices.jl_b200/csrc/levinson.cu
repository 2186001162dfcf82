// Batched Levinson / Durbin recursions (K5): one warp per independent system, the
// iterates held in registers, reversed operands kept as *shifted* companions so every
// step is lane-aligned FMAs + one rotate-shuffle per 128 elements + one fused
// two-value warp reduction.
//
// Reference recurrences: /root/reference/src/directLinearSolvers.jl
//   durbin!  :15-28, levinson! :100-121, reverse_dot :137-145, reverse_increment! :148-168,
//   levinson(T::SymmetricToeplitz, b) :123-134 (diag normalisation).
//
// Layout ("blocked-cyclic", B = 4): element i lives in lane (i/4)%32, register
// (i/128)*4 + i%4.  With yh[i] = y[k-1-i] and rh[i] = r[k-1-i] (k = current order) the
// reference's reversed dot / reversed increment become elementwise:
//   reverse_dot(r_k, x_k) = sum rh[i]*x[i];  x_k += mu*reverse(y_k)  <=>  x[i] += mu*yh[i]
//   y_k += alpha*reverse(y_k)  <=>  y[i] += alpha*yh[i],  yh <- shift_up(yh + alpha*y_old)
// and growing the order by one is a shift by one element (register rename inside a block
// of 4, one rotate shuffle across lanes per 128 elements).
#include "capi_internal.h"

namespace tb {

constexpr int LB = 4;              // block size of the blocked-cyclic layout
constexpr int LSPAN = 32 * LB;     // elements per super-row

// sums over groups of P consecutive lanes (P = 32: the whole warp; 16 / 8: two / four systems per warp)
template <typename R, int P = 32>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
    for (int o = P / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fused reduction of two values: after the first exchange lanes 0-15 carry d1 partials and lanes 16-31 d2
// partials, so the remaining four levels reduce both at once (5 double shuffles instead of 10).
template <typename R, int P = 32>
__device__ __forceinline__ void warp_sum2(R& d1, R& d2, int lane) {
    const bool hi = (lane & (P / 2)) != 0;
    R keep = hi ? d2 : d1, give = hi ? d1 : d2;
    keep += __shfl_xor_sync(0xffffffffu, give, P / 2);
#pragma unroll
    for (int o = P / 4; o >= 1; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    const int base = lane & ~(P - 1);
    d1 = __shfl_sync(0xffffffffu, keep, base);
    d2 = __shfl_sync(0xffffffffu, keep, base + P / 2);
}

// 1/x without the IEEE slow path: hardware seed (MUFU.RCP64H) + three Newton steps (error <~ 1 ulp; beta is a
// product of (1 - alpha^2) factors of an SPD recursion, i.e. finite, positive and far from the denormal range).
__device__ __forceinline__ double fast_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ float fast_rcp(float x) { return 1.0f / x; }

// Shifting vectors (yh, rh) are never moved between registers: a compile-time phase ph (= number of shifts so far
// mod 4) maps the logical slot q of a lane's block of four to the physical register (q - ph) mod 4, so growing the
// order by one is ONE in-place rotate-shuffle per super-row (the register that held slot 3 becomes slot 0 of the next
// lane) plus a phase increment -- no register copies.  Stationary vectors (x, y) use physical == logical.
__device__ __forceinline__ constexpr int phys(int i, int ph) { return (i & ~(LB - 1)) | (((i & (LB - 1)) - ph) & (LB - 1)); }

// shift the first SROWS super-rows of v (currently at phase PH) up by one element, inserting `head` at logical index 0;
// afterwards the vector is at phase (PH + 1) mod 4
template <typename R, int SROWS, int PH, int P = 32>
__device__ __forceinline__ void shift_up(R* v, R head, int lane) {
    R carry = head;                // value entering lane 0 of the current super-row
    const int sub = lane & (P - 1);
    const int src = (lane & ~(P - 1)) | ((sub + P - 1) & (P - 1));     // rotate inside the group of P lanes
#pragma unroll
    for (int s = 0; s < SROWS; ++s) {
        constexpr int wrap = (LB - 1 - PH) & (LB - 1);       // physical register holding logical slot 3
        const R rot = __shfl_sync(0xffffffffu, v[s * LB + wrap], src);
        v[s * LB + wrap] = (sub == 0) ? carry : rot;
        carry = rot;               // lane 0 sees the last lane's last element: head of the next super-row
    }
}

template <typename R, int S>
struct LevState {
    R x[S * LB], y[S * LB], yh[S * LB], rh[S * LB];
    R alpha, beta;
};

// one recursion step k = SC*128 + lb*4 + J at compile-time block position J (phase of yh/rh before the step: (J+3)%4)
template <typename R, int S, int SC, int J, bool DURBIN, bool INTERIOR = false, int P = 32>
__device__ __forceinline__ void lev_step(LevState<R, S>& st, int k, int n, int lane, int lb, const R* r_s, const R* b_s) {
    const int sub = lane & (P - 1);
    constexpr int ACT = (SC + 1) * LB;
    constexpr int PH = (J + LB - 1) & (LB - 1);
    st.beta *= ((R)1 - st.alpha * st.alpha);
    // one reciprocal per step instead of the reference's two divisions by beta (<= 1 ulp apart);
    // it depends only on the previous step, so it overlaps the dot products and the reduction
    const R ibeta = fast_rcp(st.beta);
    // four partial sums per dot product keep the FMA chains short
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
    if (!DURBIN) warp_sum2<R, P>(d1, d2, lane);
    else d2 = warp_sum<R, P>(d2);
    if (!DURBIN) {
        const R mu = (b_s[k] - d1) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) st.x[i] = fma(mu, st.yh[phys(i, PH)], st.x[i]);
        if (sub == lb) st.x[SC * LB + J] = mu;
    }
    if (DURBIN || INTERIOR || k < n - 1) {
        const R rk = r_s[k];
        st.alpha = -(rk + d2) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) {
            const R yo = st.y[i];
            st.y[i] = fma(st.alpha, st.yh[phys(i, PH)], yo);
            st.yh[phys(i, PH)] = fma(st.alpha, yo, st.yh[phys(i, PH)]);
        }
        if (sub == lb) st.y[SC * LB + J] = st.alpha;
        shift_up<R, SC + 1, PH, P>(st.yh, st.alpha, lane);
        shift_up<R, SC + 1, PH, P>(st.rh, rk, lane);
    }
}

// all steps k in super-row SC (k = SC*128 .. SC*128+127), then recurse to SC+1
template <typename R, int S, int SC, bool DURBIN, int P = 32>
__device__ __forceinline__ void run_superrows(LevState<R, S>& st, int n, int lane, const R* r_s, const R* b_s) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < P; ++lb) {
            const int k0 = SC * (P * LB) + lb * LB;
            if (k0 >= n) break;
            if (k0 >= 1 && k0 + LB < n) {      // interior block: four unconditional steps (no control-flow merges,
                lev_step<R, S, SC, 0, DURBIN, true, P>(st, k0 + 0, n, lane, lb, r_s, b_s);   // so no register copies)
                lev_step<R, S, SC, 1, DURBIN, true, P>(st, k0 + 1, n, lane, lb, r_s, b_s);
                lev_step<R, S, SC, 2, DURBIN, true, P>(st, k0 + 2, n, lane, lb, r_s, b_s);
                lev_step<R, S, SC, 3, DURBIN, true, P>(st, k0 + 3, n, lane, lb, r_s, b_s);
                continue;
            }
            if (k0 + 0 >= 1 && k0 + 0 < n) lev_step<R, S, SC, 0, DURBIN, false, P>(st, k0 + 0, n, lane, lb, r_s, b_s);
            if (k0 + 1 < n) lev_step<R, S, SC, 1, DURBIN, false, P>(st, k0 + 1, n, lane, lb, r_s, b_s);
            if (k0 + 2 < n) lev_step<R, S, SC, 2, DURBIN, false, P>(st, k0 + 2, n, lane, lb, r_s, b_s);
            if (k0 + 3 < n) lev_step<R, S, SC, 3, DURBIN, false, P>(st, k0 + 3, n, lane, lb, r_s, b_s);
        }
        run_superrows<R, S, SC + 1, DURBIN, P>(st, n, lane, r_s, b_s);
    }
}

template <typename R, int S, bool DURBIN, int P = 32>
__global__ void __launch_bounds__(128, DURBIN ? 4 : 3)
k_levinson_warp(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm,
                int64_t ldb, R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
    // P lanes per system: 32 / P systems share a warp and run the same instruction stream (P < 32 for small n,
    // where a step is a handful of FMAs and the reduction + rotate latency is what one pays)
    constexpr int NREG = S * LB;
    constexpr int SPAN = P * LB;
    constexpr int SPW = 32 / P;                            // systems per warp
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = (int)n64;
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int sub = lane & (P - 1), grp = lane / P;
    const int64_t sys0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + wic) * SPW + grp;
    if (sys0 - grp >= nsys) return;                        // the whole warp is past the end
    const bool live = sys0 < nsys;
    const int64_t sys = live ? sys0 : nsys - 1;            // idle groups shadow the last system, never store
    const int nr = DURBIN ? n : n - 1;                     // entries of r
    R* r_s = reinterpret_cast<R*>(smraw) + ((size_t)wic * SPW + grp) * (2 * S * SPAN);
    R* b_s = r_s + S * SPAN;
    const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
    const bool scale = (dg != (R)1);
    for (int i = sub; i < nr; i += P) {
        const R v = Rm[sys * ldr + i];
        r_s[i] = scale ? v / dg : v;
    }
    if (!DURBIN)
        for (int i = sub; i < n; i += P) b_s[i] = Bm[sys * ldb + i];
    __syncwarp();

    LevState<R, S> st;
#pragma unroll
    for (int i = 0; i < NREG; ++i) { st.x[i] = 0; st.y[i] = 0; st.yh[i] = 0; st.rh[i] = 0; }
    if (n == 1 && !DURBIN) {          // 1x1 system: x = b (levinson! with empty r is not reachable in the reference)
        if (sub == 0 && live) Xm[sys * ldx] = scale ? b_s[0] / dg : b_s[0];
        return;
    }
    const R r0 = r_s[0];
    if (sub == 0) { st.y[0] = -r0; st.yh[0] = -r0; st.rh[0] = r0; if (!DURBIN) st.x[0] = b_s[0]; }
    st.alpha = -r0;
    st.beta = (R)1;
    run_superrows<R, S, 0, DURBIN, P>(st, n, lane, r_s, b_s);

    if (!live) return;
    R* out = Xm + sys * ldx;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = s * SPAN + sub * LB + j;
            if (i < n) {
                R v = DURBIN ? st.y[s * LB + j] : st.x[s * LB + j];
                if (scale) v /= dg;
                out[i] = v;
            }
        }
}

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
static int launch_levinson(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x,
                           int64_t ldx, const R* diag, cudaStream_t s) {
    if (nsys <= 0 || n <= 0) return TB200_OK;
    if (n <= 4 * 16 * LB) {                   // small orders: 8 (n <= 64), 4 (n <= 128) or 2 (n <= 256) systems per warp
        const int wpc = 4;
#define TB_LEV_SMALL(SV, PV)                                                                                        \
    {                                                                                                               \
        const unsigned grid = (unsigned)((nsys + wpc * (32 / PV) - 1) / (wpc * (32 / PV)));                        \
        const size_t smem = (size_t)wpc * (32 / PV) * 2 * SV * (PV * LB) * sizeof(R);                               \
        k_levinson_warp<R, SV, DURBIN, PV><<<grid, wpc * 32, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);    \
    }
        // smallest group of lanes that still holds a system in S <= 4 register blocks per lane
        if (n <= 16) TB_LEV_SMALL(1, 4)
        else if (n <= 32) TB_LEV_SMALL(2, 4)
        else if (n <= 48) TB_LEV_SMALL(3, 4)
        else if (n <= 64) TB_LEV_SMALL(4, 4)
        else if (n <= 128 && !DURBIN) TB_LEV_SMALL(2, 16)      // measured: Levinson 102 M (16 lanes) vs 93 M (8 lanes) at n = 128
        else if (n <= 96) TB_LEV_SMALL(3, 8)
        else if (n <= 128) TB_LEV_SMALL(4, 8)
        else if (n <= 192) TB_LEV_SMALL(3, 16)
        else TB_LEV_SMALL(4, 16)
#undef TB_LEV_SMALL
        count_launch();
        return cuda_status(cudaGetLastError(), "levinson warp kernel (several systems per warp)");
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
        cudaError_t e;
        if (n <= 6 * LSPAN) e = launch_multiwarp<R, 3, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        else if (n <= 8 * LSPAN) e = launch_multiwarp<R, 4, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        else if (n <= 12 * LSPAN) e = launch_multiwarp<R, 3, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        else e = launch_multiwarp<R, 4, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
        count_launch();
        return cuda_status(e, "levinson multi-warp kernel");
    }
    const size_t smem = (size_t)5 * n * sizeof(R);
    if (smem > 200 * 1024) return set_error(TB200_UNSUPPORTED, "levinson/durbin: n too large for the shared-memory kernel");
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
