// Levinson / Durbin for 256 < n <= 512, register-lean variant ("sx"): one warp per system as in
// levinson_kernels.cuh, but only the Durbin pair (y and its reversed companion yh) lives in registers.
//   * reverse(r_k) is not kept as a shifting register vector: rh[i] = r[k-1-i] is read from the staged copy of r in
//     shared memory (behind a 128-entry zero pad, so indices just past the current order read zeros) -- 32 registers
//     and one rotate-shuffle chain per step less;
//   * x lives in shared memory: a lane only ever touches its own elements (blocked-cyclic layout), so the dot
//     reverse_dot(r_k, x_k) and the update x_k += mu * reverse(y_k) are private load / FMA / store sequences.
// The thread state drops from 64 to 32 doubles, which lets four (Levinson) to six (Durbin) CTAs of four warps share
// an SM instead of three / four: the recursion is bound by the dependent reduce -> scalar -> update -> shift chain of
// each warp, so more resident warps is what raises FP64 pipe utilisation.
//
// Reference recurrences: /root/reference/src/directLinearSolvers.jl:15-28 (durbin!), :100-121 (levinson!),
// :137-168 (reverse_dot, reverse_increment!), :123-134 (diag normalisation).
#include "levinson_kernels.cuh"

namespace tb {

constexpr int SX_PAD = LSPAN;          // zeros in front of r: k-1-i >= -128 inside the active super-rows

template <typename R, int S>
struct DurState {
    R y[S * LB], yh[S * LB];
    R alpha, beta;
};

// one recursion step k = SC*128 + lb*4 + J
template <typename R, int S, int SC, int J, bool DURBIN, bool INTERIOR>
__device__ __forceinline__ void sx_step(DurState<R, S>& st, int k, int n, int lane, int lb, const R* rz, R* xs, const R* __restrict__ bg) {
    constexpr int ACT = (SC + 1) * LB;
    constexpr int PH = (J + LB - 1) & (LB - 1);
    st.beta *= ((R)1 - st.alpha * st.alpha);
    const R ibeta = fast_rcp(st.beta);
    R bk = 0;
    if (!DURBIN) bk = __ldg(bg + k);
    const R* rr = rz + SX_PAD + (k - 1) - lane * LB;          // rh[s*128 + lane*4 + j] = rr[-(s*128) - j]
    R p1[LB], p2[LB];
#pragma unroll
    for (int q = 0; q < LB; ++q) { p1[q] = 0; p2[q] = 0; }
#pragma unroll
    for (int i = 0; i < ACT; ++i) {
        const R rh = rr[-((i / LB) * LSPAN) - (i % LB)];
        if (!DURBIN) p1[i % LB] = fma(rh, xs[(i / LB) * LSPAN + lane * LB + (i % LB)], p1[i % LB]);
        p2[i % LB] = fma(rh, st.y[i], p2[i % LB]);
    }
    R d1 = (p1[0] + p1[1]) + (p1[2] + p1[3]);
    R d2 = (p2[0] + p2[1]) + (p2[2] + p2[3]);
    if (!DURBIN) warp_sum2<R, 32>(d1, d2, lane);
    else d2 = warp_sum<R, 32>(d2);
    if (!DURBIN) {
        const R mu = (bk - d1) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) {
            R* px = xs + (i / LB) * LSPAN + lane * LB + (i % LB);
            *px = fma(mu, st.yh[phys(i, PH)], *px);
        }
        if (lane == lb) xs[SC * LSPAN + lb * LB + J] = mu;
    }
    if (DURBIN || INTERIOR || k < n - 1) {
        const R rk = rz[SX_PAD + k];
        st.alpha = -(rk + d2) * ibeta;
#pragma unroll
        for (int i = 0; i < ACT; ++i) {
            const R yo = st.y[i];
            st.y[i] = fma(st.alpha, st.yh[phys(i, PH)], yo);
            st.yh[phys(i, PH)] = fma(st.alpha, yo, st.yh[phys(i, PH)]);
        }
        if (lane == lb) st.y[SC * LB + J] = st.alpha;
        shift_up<R, SC + 1, PH, 32>(st.yh, st.alpha, lane);
    }
}

template <typename R, int S, int SC, bool DURBIN>
__device__ __forceinline__ void sx_superrows(DurState<R, S>& st, int n, int lane, const R* rz, R* xs, const R* __restrict__ bg) {
    if constexpr (SC < S) {
        for (int lb = 0; lb < 32; ++lb) {
            const int k0 = SC * LSPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 >= 1 && k0 + LB < n) {
                sx_step<R, S, SC, 0, DURBIN, true>(st, k0 + 0, n, lane, lb, rz, xs, bg);
                sx_step<R, S, SC, 1, DURBIN, true>(st, k0 + 1, n, lane, lb, rz, xs, bg);
                sx_step<R, S, SC, 2, DURBIN, true>(st, k0 + 2, n, lane, lb, rz, xs, bg);
                sx_step<R, S, SC, 3, DURBIN, true>(st, k0 + 3, n, lane, lb, rz, xs, bg);
                continue;
            }
            if (k0 + 0 >= 1 && k0 + 0 < n) sx_step<R, S, SC, 0, DURBIN, false>(st, k0 + 0, n, lane, lb, rz, xs, bg);
            if (k0 + 1 < n) sx_step<R, S, SC, 1, DURBIN, false>(st, k0 + 1, n, lane, lb, rz, xs, bg);
            if (k0 + 2 < n) sx_step<R, S, SC, 2, DURBIN, false>(st, k0 + 2, n, lane, lb, rz, xs, bg);
            if (k0 + 3 < n) sx_step<R, S, SC, 3, DURBIN, false>(st, k0 + 3, n, lane, lb, rz, xs, bg);
        }
        sx_superrows<R, S, SC + 1, DURBIN>(st, n, lane, rz, xs, bg);
    }
}

// shared memory per warp: [SX_PAD zeros][S*128 entries of r (zero filled past its end)][S*128 entries of x]
template <typename R, int S, bool DURBIN>
__global__ void __launch_bounds__(128, DURBIN ? 5 : 4)
k_levinson_sx(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm, int64_t ldb,
              R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
    constexpr int SPANS = S * LSPAN;
    constexpr int PERW = SX_PAD + SPANS + (DURBIN ? 0 : SPANS);
    extern __shared__ __align__(16) unsigned char smraw[];
    const int n = (int)n64;
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    const int64_t sys = (int64_t)blockIdx.x * (blockDim.x >> 5) + wic;
    if (sys >= nsys) return;
    R* rz = reinterpret_cast<R*>(smraw) + (size_t)wic * PERW;
    R* xs = rz + SX_PAD + SPANS;
    const int nr = DURBIN ? n : n - 1;
    const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
    const bool scale = (dg != (R)1);
    for (int i = lane; i < SX_PAD; i += 32) rz[i] = 0;
    for (int i = lane; i < SPANS; i += 32) {
        R v = 0;
        if (i < nr) { v = Rm[sys * ldr + i]; if (scale) v = v / dg; }
        rz[SX_PAD + i] = v;
        if (!DURBIN) xs[i] = 0;
    }
    const R* bg = DURBIN ? nullptr : Bm + sys * ldb;
    __syncwarp();

    DurState<R, S> st;
#pragma unroll
    for (int i = 0; i < S * LB; ++i) { st.y[i] = 0; st.yh[i] = 0; }
    const R r0 = rz[SX_PAD];
    if (lane == 0) { st.y[0] = -r0; st.yh[0] = -r0; if (!DURBIN) xs[0] = __ldg(bg); }
    st.alpha = -r0;
    st.beta = (R)1;
    __syncwarp();
    sx_superrows<R, S, 0, DURBIN>(st, n, lane, rz, xs, bg);

    R* out = Xm + sys * ldx;
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = s * LSPAN + lane * LB + j;
            if (i < n) {
                R v = DURBIN ? st.y[s * LB + j] : xs[i];
                if (scale) v /= dg;
                out[i] = v;
            }
        }
}

template <typename R, bool DURBIN>
cudaError_t launch_levinson_sx(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                               const R* diag, cudaStream_t s) {
    const int S = (int)((n + LSPAN - 1) / LSPAN);
    const int wpc = 4;
    const unsigned grid = (unsigned)((nsys + wpc - 1) / wpc);
    const size_t smem = (size_t)wpc * (SX_PAD + (size_t)S * LSPAN * (DURBIN ? 1 : 2)) * sizeof(R);
#define TB_SX_CASE(SV)                                                                                            \
    case SV: {                                                                                                    \
        auto kern = k_levinson_sx<R, SV, DURBIN>;                                                                 \
        static bool configured[64] = {};                                                                          \
        int dev = 0;                                                                                              \
        cudaGetDevice(&dev);                                                                                      \
        if (!configured[dev & 63]) {                                                                              \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
            if (e != cudaSuccess) return e;                                                                       \
            cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);                     \
            configured[dev & 63] = true;                                                                          \
        }                                                                                                         \
        kern<<<grid, wpc * 32, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);                                 \
        break;                                                                                                    \
    }
    switch (S) {
        TB_SX_CASE(3) TB_SX_CASE(4)
    default: return cudaErrorInvalidValue;
    }
#undef TB_SX_CASE
    return cudaGetLastError();
}

template cudaError_t launch_levinson_sx<double, false>(int64_t, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, const double*, cudaStream_t);
template cudaError_t launch_levinson_sx<double, true>(int64_t, int64_t, const double*, int64_t, const double*, int64_t, double*, int64_t, const double*, cudaStream_t);
template cudaError_t launch_levinson_sx<float, false>(int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, const float*, cudaStream_t);
template cudaError_t launch_levinson_sx<float, true>(int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, const float*, cudaStream_t);

}  // namespace tb
