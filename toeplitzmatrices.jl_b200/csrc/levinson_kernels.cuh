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
#pragma once
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
// same with two Newton steps: the seed carries >= 20 bits, so two steps reach the rounding floor (<= 2 ulp); the third
// step of fast_rcp only polishes the last bit and sits on the critical alpha -> beta -> 1/beta -> alpha chain
__device__ __forceinline__ double fast_rcp2(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ float fast_rcp2(float x) { return 1.0f / x; }
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

// other translation units (built in parallel): several systems per warp (n <= 256) and several warps per system
// (512 < n <= 2048)
template <typename R, bool DURBIN>
cudaError_t launch_levinson_small(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                  const R* diag, cudaStream_t s);
template <typename R, bool DURBIN>
cudaError_t launch_levinson_mw(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                               const R* diag, cudaStream_t s);

// generator (Schur) form for n <= 512: no reductions, 3 FMAs per slot and step (levinson_gen.cu)
template <typename R, bool DURBIN>
cudaError_t launch_levinson_gen(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                const R* diag, cudaStream_t s);
// generator form, 2 or 4 warps per system, 512 < n <= 2048 (levinson_gen_mw.cu)
template <typename R, bool DURBIN>
cudaError_t launch_levinson_gen_mw(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                   const R* diag, cudaStream_t s);
extern int g_lev_gen;     // 1 (default): n <= 512 takes the generator-form kernels; 0: the dot-product kernels (tb200_set_option("lev_gen"))

}  // namespace tb
