// Levinson / Durbin in generator form for 512 < n <= 2048: W = 2 or 4 warps per system.
//
// Same recursion as levinson_gen.cu (three n-slot vectors A, B, X; per step  alpha = -A[k]/beta, mu = -X[k]/beta,
// X += mu*B, (A, B) <- (A + alpha*B, shift(B + alpha*A))), with the register layout stretched over 32*W lanes
// (slot s: lane (s/4) % (32 W), register (s / (128 W))*4 + s%4).  Nothing is reduced, so the warps of a system meet
// ONCE per step, at a named barrier, to pick up
//   (a) the two pivots A[k], X[k], published by the lane that owns slot k at the end of step k-1, and
//   (b) the element of B that crosses each warp boundary in the shift: the shift of step k-1 is deferred to the start
//       of step k, after the barrier, where it is still ahead of B's next use.
// Exchange slots are double-buffered by step parity (a warp runs at most one step ahead of the slowest warp of its
// system).  The dot-product kernels of levinson_multiwarp.cu remain selectable (option lev_gen = 0).
//
// Reference recurrences: /root/reference/src/directLinearSolvers.jl:15-28 (durbin!), :100-121 (levinson!), :123-134.
#include "levinson_kernels.cuh"

namespace tb {

template <typename R, int S, bool DURBIN>
struct GenStateMW {
    R a[S * LB], b[S * LB], x[DURBIN ? 1 : S * LB];
    R alpha, beta;
};

// one step k = SC*SPAN + lb*4 + J; during the step B carries k shifts (phase J); LAST: final step of levinson! (mu only)
template <typename R, int S, int W, int SC, int J, bool DURBIN, bool LAST>
__device__ __forceinline__ void gmw_step(GenStateMW<R, S, DURBIN>& st, int k, int lane, int wis, int lb, R* xch, int bar_id) {
    constexpr int N = S * LB;
    constexpr int PH = J;
    constexpr int PHP = (J + LB - 1) & (LB - 1);           // phase of B before the pending shift
    constexpr int WRAPP = (LB - 1 - PHP) & (LB - 1);       // register that rotates out in the pending shift
    constexpr int WRAPN = (LB - 1 - PH) & (LB - 1);        // ... in the shift this step leaves pending
    constexpr int XS = 2 + W * S;                          // values per parity buffer
    const int glane = wis * 32 + lane;
    R* slot = xch + (k & 1) * XS;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * W) : "memory");
    const R ak = slot[0], xk = slot[1];
    if (k > 0) {                                           // deferred shift of B, head = the previous step's alpha
#pragma unroll
        for (int sr = 0; sr < S; ++sr) {
            const R rot = __shfl_sync(0xffffffffu, st.b[sr * LB + WRAPP], (lane + 31) & 31);
            R in = rot;
            if (lane == 0) {
                if (wis > 0) in = slot[2 + (wis - 1) * S + sr];
                else in = (sr == 0) ? st.alpha : slot[2 + (W - 1) * S + (sr - 1)];
            }
            st.b[sr * LB + WRAPP] = in;
        }
    }
    st.beta *= ((R)1 - st.alpha * st.alpha);               // alpha = 0 before the first step: beta_0 = 1
    const R ibeta = fast_rcp2(st.beta);
    if constexpr (!DURBIN) {
        const R mu = -xk * ibeta;
#pragma unroll
        for (int i = 0; i < N; ++i) st.x[i] = fma(mu, st.b[phys(i, PH)], st.x[i]);
        if (glane == lb) st.x[SC * LB + J] = mu;
    }
    if constexpr (!LAST) {
        const R alpha = -ak * ibeta;
        st.alpha = alpha;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const R ao = st.a[i], bo = st.b[phys(i, PH)];
            st.a[i] = fma(alpha, bo, ao);
            st.b[phys(i, PH)] = fma(alpha, ao, bo);
        }
        if (glane == lb) st.a[SC * LB + J] = alpha;
        // publish for step k+1: boundary elements of the shift left pending, and the pivots of slot k+1
        R* nslot = xch + ((k + 1) & 1) * XS;
        if (lane == 31) {
#pragma unroll
            for (int sr = 0; sr < S; ++sr) nslot[2 + wis * S + sr] = st.b[sr * LB + WRAPN];
        }
        constexpr int RN = SC * LB + ((J + 1) & (LB - 1));             // same lane (J < 3) / first register of the next lane's block
        constexpr int RW = (SC + 1 < S) ? (SC + 1) * LB : SC * LB;     // ... of the next super-row, lane 0
        const bool wrap = (J == LB - 1) && (lb == 32 * W - 1);
        const int owner_next = (J == LB - 1) ? (wrap ? 0 : lb + 1) : lb;
        if (glane == owner_next) {
            nslot[0] = wrap ? st.a[RW] : st.a[RN];
            if constexpr (!DURBIN) nslot[1] = wrap ? st.x[RW] : st.x[RN];
        }
    }
}

template <typename R, int S, int W, int SC, bool DURBIN>
__device__ __forceinline__ void gmw_superrows(GenStateMW<R, S, DURBIN>& st, int n, int lane, int wis, R* xch, int bar_id) {
    if constexpr (SC < S) {
        constexpr int SPAN = LSPAN * W;
        const int nalpha = DURBIN ? n : n - 1;             // steps that need alpha
        for (int lb = 0; lb < 32 * W; ++lb) {
            const int k0 = SC * SPAN + lb * LB;
            if (k0 >= n) break;
            if (k0 + LB <= nalpha) {
                gmw_step<R, S, W, SC, 0, DURBIN, false>(st, k0 + 0, lane, wis, lb, xch, bar_id);
                gmw_step<R, S, W, SC, 1, DURBIN, false>(st, k0 + 1, lane, wis, lb, xch, bar_id);
                gmw_step<R, S, W, SC, 2, DURBIN, false>(st, k0 + 2, lane, wis, lb, xch, bar_id);
                gmw_step<R, S, W, SC, 3, DURBIN, false>(st, k0 + 3, lane, wis, lb, xch, bar_id);
                continue;
            }
            if (k0 + 0 < nalpha) gmw_step<R, S, W, SC, 0, DURBIN, false>(st, k0 + 0, lane, wis, lb, xch, bar_id);
            else if (!DURBIN && k0 + 0 < n) gmw_step<R, S, W, SC, 0, DURBIN, true>(st, k0 + 0, lane, wis, lb, xch, bar_id);
            if (k0 + 1 < nalpha) gmw_step<R, S, W, SC, 1, DURBIN, false>(st, k0 + 1, lane, wis, lb, xch, bar_id);
            else if (!DURBIN && k0 + 1 < n) gmw_step<R, S, W, SC, 1, DURBIN, true>(st, k0 + 1, lane, wis, lb, xch, bar_id);
            if (k0 + 2 < nalpha) gmw_step<R, S, W, SC, 2, DURBIN, false>(st, k0 + 2, lane, wis, lb, xch, bar_id);
            else if (!DURBIN && k0 + 2 < n) gmw_step<R, S, W, SC, 2, DURBIN, true>(st, k0 + 2, lane, wis, lb, xch, bar_id);
            if (k0 + 3 < nalpha) gmw_step<R, S, W, SC, 3, DURBIN, false>(st, k0 + 3, lane, wis, lb, xch, bar_id);
            else if (!DURBIN && k0 + 3 < n) gmw_step<R, S, W, SC, 3, DURBIN, true>(st, k0 + 3, lane, wis, lb, xch, bar_id);
        }
        gmw_superrows<R, S, W, SC + 1, DURBIN>(st, n, lane, wis, xch, bar_id);
    }
}

// 128 threads: 4/W systems per CTA
template <typename R, int S, int W, bool DURBIN>
__global__ void __launch_bounds__(128, sizeof(R) == 8 ? (DURBIN ? (S >= 4 ? 3 : 4) : (S >= 4 ? 2 : 3)) : 4)
k_levinson_gen_mw(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm, int64_t ldb,
                  R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
    constexpr int SPAN = LSPAN * W;
    constexpr int XS = 2 + W * S;
    constexpr int PER_SYS = S * SPAN + LB + 2 * XS;        // staged r behind the unit diagonal, two exchange buffers
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
    R* xch = r_s + S * SPAN + LB;
    const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
    const bool scale = (dg != (R)1);
    if (glane == 0) r_s[0] = (R)1;
    for (int i = glane; i < S * SPAN; i += 32 * W) {
        R v = 0;
        if (i < nr) { v = Rm[sys * ldr + i]; if (scale) v /= dg; }
        r_s[1 + i] = v;
    }
    const R* bcol = DURBIN ? nullptr : Bm + sys * ldb;
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(32 * W) : "memory");

    GenStateMW<R, S, DURBIN> st;
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int s = sr * SPAN + glane * LB + j;
            st.a[sr * LB + j] = r_s[1 + s];
            st.b[sr * LB + j] = (s < n) ? r_s[s] : (R)0;
            if constexpr (!DURBIN) st.x[sr * LB + j] = (s < n) ? -bcol[s] : (R)0;
        }
    if (glane == 0) {                                      // pivots of step 0 (parity 0): A[0] = r[1], X[0] = -b[1]
        xch[0] = r_s[1];
        xch[1] = DURBIN ? (R)0 : -bcol[0];
    }
    st.alpha = 0;
    st.beta = (R)1;
    gmw_superrows<R, S, W, 0, DURBIN>(st, n, lane, wis, xch, bar_id);

    R* out = Xm + sys * ldx;
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = sr * SPAN + glane * LB + j;
            if (i < n) {
                R v;
                if constexpr (DURBIN) v = st.a[sr * LB + j]; else v = st.x[sr * LB + j];
                if (scale) v /= dg;
                out[i] = v;
            }
        }
}

template <typename R, int S, int W, bool DURBIN>
static cudaError_t launch_gen_mw(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                 const R* diag, cudaStream_t s) {
    constexpr int SPAN = LSPAN * W;
    constexpr int XS = 2 + W * S;
    constexpr size_t smem = (size_t)(4 / W) * (S * SPAN + LB + 2 * XS) * sizeof(R);
    auto kern = k_levinson_gen_mw<R, S, W, DURBIN>;
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
cudaError_t launch_levinson_gen_mw(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                   const R* diag, cudaStream_t s) {
    if (n <= 6 * LSPAN) return launch_gen_mw<R, 3, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    if (n <= 8 * LSPAN) return launch_gen_mw<R, 4, 2, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    if (n <= 12 * LSPAN) return launch_gen_mw<R, 3, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
    return launch_gen_mw<R, 4, 4, DURBIN>(n, nsys, r, ldr, b, ldb, x, ldx, diag, s);
}

#define TB_INST(R, D) template cudaError_t launch_levinson_gen_mw<R, D>(int64_t, int64_t, const R*, int64_t, const R*, int64_t, R*, int64_t, const R*, cudaStream_t);
TB_INST(float, false) TB_INST(float, true) TB_INST(double, false) TB_INST(double, true)
#undef TB_INST

}  // namespace tb
