// Levinson / Durbin in generator form, n <= 512: one warp (or one group of P lanes) per system, NO reductions.
//
// The reference recursion (/root/reference/src/directLinearSolvers.jl:15-28 durbin!, :100-121 levinson!) spends every
// step in two reversed dot products
//     mu_k    = (b[k+1] - reverse_dot(r_k, x_k)) / beta_k
//     alpha_k = -(r[k+1] + reverse_dot(r_k, y_k)) / beta_k
// whose warp reductions are a dependent chain of shuffles that no amount of FMA throughput hides.  Both numerators are
// entries of vectors that obey the SAME two-term recurrence as y itself (the Schur / generator form of the
// Levinson-Durbin recursion):  with e_k = (1, y_k, 0...), its reversal ehat_k, u_k = T e_k, v_k = T ehat_k and
// w_k = b - T (x_k, 0...),
//     r[k+1] + reverse_dot(r_k, y_k) = u_k[k+1],   b[k+1] - reverse_dot(r_k, x_k) = w_k[k],   beta_k = v_k[k]
//     u_{k+1}[j] = u_k[j] + alpha_k v_k[j-1],  v_{k+1}[j] = v_k[j-1] + alpha_k u_k[j],  w_{k+1}[j] = w_k[j] - mu_k v_k[j]
// which is exactly the shape of  y += alpha*reverse(y)  and  x += mu*reverse(y).  The entries of y / reverse(y) / x that
// exist at order k (indices < k) and the entries of u / v / w that are still needed (indices > k) are complementary,
// so one array of n slots holds both halves:
//     A[s] = y_k[s]            (s < k),   u_k[s+1]  (s >= k)        stationary
//     B[s] = reverse(y_k)[s]   (s < k),   v_k[s]    (s >= k)        shifts up by one slot per step
//     X[s] = x_k[s]            (s < k),  -w_k[s]    (s >= k)        stationary
// and a step is, for ALL slots alike,
//     alpha = -A[k]/beta, mu = -X[k]/beta;   X += mu*B;  X[k] = mu;   (A, B) <- (A + alpha*B, shift(B + alpha*A));  A[k] = alpha
// i.e. 3 FMAs per slot (2 for Durbin), two one-element broadcasts, one rotate-shuffle per 128 slots -- no reduction,
// no reversed-r vector, 96 instead of 128 state registers.  beta follows the reference's own product recurrence
// beta *= 1 - alpha^2 (:106), so the divisions are the reference's; alpha and mu are mathematically identical to the
// dot-product forms and differ by rounding only (the generator recursion is the backward-stable Schur algorithm).
// Start:  A = r (r[1..]), B = (1, r[1], r[2], ...), X = -b.
//
// Register layout as in levinson_kernels.cuh (blocked-cyclic, phase-indexed shifting vector).
// (A software-pipelined step -- slot k+1 updated and broadcast first, 1/beta_{k+1} started beside the FMAs -- measured
// the same 12.5 M solves/s: the kernel is throttled by the FP64 pipe, not by its dependent chain; profiles/experiments_r02.md.)
#include "levinson_kernels.cuh"

namespace tb {

template <typename R, int S, bool DURBIN>
struct GenState {
    R a[S * LB], b[S * LB], x[DURBIN ? 1 : S * LB];
    R alpha, beta;
};

// one step k = SC*SPAN + lb*4 + J (J, SC compile time).  B has been shifted k times: phase J.
// LAST: the final step of levinson! (k = n-1) needs mu only (:114 `if k < n-1`).
template <typename R, int S, int SC, int J, bool DURBIN, bool LAST, int P>
__device__ __forceinline__ void gen_step(GenState<R, S, DURBIN>& st, int lane, int lb) {
    constexpr int N = S * LB;
    constexpr int PH = J;
    const int sub = lane & (P - 1);
    const int owner = (lane & ~(P - 1)) | lb;
    st.beta *= ((R)1 - st.alpha * st.alpha);          // alpha = 0 before the first step: beta_0 = 1
    const R ibeta = fast_rcp2(st.beta);
    const R ak = __shfl_sync(0xffffffffu, st.a[SC * LB + J], owner);
    if constexpr (!DURBIN) {
        const R xk = __shfl_sync(0xffffffffu, st.x[SC * LB + J], owner);
        const R mu = -xk * ibeta;
#pragma unroll
        for (int i = 0; i < N; ++i) st.x[i] = fma(mu, st.b[phys(i, PH)], st.x[i]);
        if (sub == lb) st.x[SC * LB + J] = mu;
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
        if (sub == lb) st.a[SC * LB + J] = alpha;
        shift_up<R, S, PH, P>(st.b, alpha, lane);
    }
}

template <typename R, int S, int SC, bool DURBIN, int P>
__device__ __forceinline__ void gen_superrows(GenState<R, S, DURBIN>& st, int n, int lane) {
    if constexpr (SC < S) {
        // steps that need alpha: all n of durbin!, the first n-1 of levinson!
        const int nalpha = DURBIN ? n : n - 1;
        for (int lb = 0; lb < P; ++lb) {
            const int k0 = SC * (P * LB) + lb * LB;
            if (k0 >= n) break;
            if (k0 + LB <= nalpha) {
                gen_step<R, S, SC, 0, DURBIN, false, P>(st, lane, lb);
                gen_step<R, S, SC, 1, DURBIN, false, P>(st, lane, lb);
                gen_step<R, S, SC, 2, DURBIN, false, P>(st, lane, lb);
                gen_step<R, S, SC, 3, DURBIN, false, P>(st, lane, lb);
                continue;
            }
            // the block that contains the end of the recursion
            if (k0 + 0 < nalpha) gen_step<R, S, SC, 0, DURBIN, false, P>(st, lane, lb);
            else if (!DURBIN && k0 + 0 < n) gen_step<R, S, SC, 0, DURBIN, true, P>(st, lane, lb);
            if (k0 + 1 < nalpha) gen_step<R, S, SC, 1, DURBIN, false, P>(st, lane, lb);
            else if (!DURBIN && k0 + 1 < n) gen_step<R, S, SC, 1, DURBIN, true, P>(st, lane, lb);
            if (k0 + 2 < nalpha) gen_step<R, S, SC, 2, DURBIN, false, P>(st, lane, lb);
            else if (!DURBIN && k0 + 2 < n) gen_step<R, S, SC, 2, DURBIN, true, P>(st, lane, lb);
            if (k0 + 3 < nalpha) gen_step<R, S, SC, 3, DURBIN, false, P>(st, lane, lb);
            else if (!DURBIN && k0 + 3 < n) gen_step<R, S, SC, 3, DURBIN, true, P>(st, lane, lb);
        }
        gen_superrows<R, S, SC + 1, DURBIN, P>(st, n, lane);
    }
}

// CTAs of 4 warps per SM the register budget is cut for: the state is (2 or 3) * 4S values per lane plus ~35 registers
// of addressing / scalars; the budgets below are the largest occupancy that compiles without spills
#ifndef TB_GEN_LEV_CTAS
#define TB_GEN_LEV_CTAS 4
#endif
#ifndef TB_GEN_DUR_CTAS
#define TB_GEN_DUR_CTAS 4
#endif
template <typename R, bool DURBIN, int S> constexpr int gen_min_ctas() {
    if (sizeof(R) == 8) return DURBIN ? (S == 4 ? TB_GEN_DUR_CTAS : 5) : (S >= 3 ? TB_GEN_LEV_CTAS : 4);
    return DURBIN ? (S == 4 ? 5 : 6) : (S == 4 ? 4 : 6);
}

template <typename R, int S, bool DURBIN, int P>
__global__ void __launch_bounds__(128, gen_min_ctas<R, DURBIN, S>())
k_levinson_gen(int64_t n64, int64_t nsys, const R* __restrict__ Rm, int64_t ldr, const R* __restrict__ Bm, int64_t ldb,
               R* __restrict__ Xm, int64_t ldx, const R* __restrict__ diag) {
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
    // staged copy of r behind one slot holding the unit diagonal: B = (1, r[1], r[2], ...) is the same array one slot earlier
    R* r_s = reinterpret_cast<R*>(smraw) + ((size_t)wic * SPW + grp) * (S * SPAN + LB);
    const R dg = (!DURBIN && diag) ? diag[sys] : (R)1;
    const bool scale = (dg != (R)1);
    if (sub == 0) r_s[0] = (R)1;
    for (int i = sub; i < S * SPAN; i += P) {
        R v = 0;
        if (i < nr) { v = Rm[sys * ldr + i]; if (scale) v /= dg; }
        r_s[1 + i] = v;
    }
    __syncwarp();

    GenState<R, S, DURBIN> st;
    const R* bcol = DURBIN ? nullptr : Bm + sys * ldb;
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int s = sr * SPAN + sub * LB + j;
            st.a[sr * LB + j] = r_s[1 + s];
            st.b[sr * LB + j] = (s < n) ? r_s[s] : (R)0;
            if constexpr (!DURBIN) st.x[sr * LB + j] = (s < n) ? -bcol[s] : (R)0;
        }
    st.alpha = 0;
    st.beta = (R)1;
    gen_superrows<R, S, 0, DURBIN, P>(st, n, lane);

    if (!live) return;
    R* out = Xm + sys * ldx;
#pragma unroll
    for (int sr = 0; sr < S; ++sr)
#pragma unroll
        for (int j = 0; j < LB; ++j) {
            const int i = sr * SPAN + sub * LB + j;
            if (i < n) {
                R v;
                if constexpr (DURBIN) v = st.a[sr * LB + j]; else v = st.x[sr * LB + j];
                if (scale) v /= dg;
                out[i] = v;
            }
        }
}

template <typename R, bool DURBIN>
cudaError_t launch_levinson_gen(int64_t n, int64_t nsys, const R* r, int64_t ldr, const R* b, int64_t ldb, R* x, int64_t ldx,
                                const R* diag, cudaStream_t s) {
    const int wpc = 4;
#define TB_LEV_GEN(SV, PV)                                                                                          \
    {                                                                                                               \
        const unsigned grid = (unsigned)((nsys + wpc * (32 / PV) - 1) / (wpc * (32 / PV)));                        \
        const size_t smem = (size_t)wpc * (32 / PV) * (SV * (PV * LB) + LB) * sizeof(R);                            \
        k_levinson_gen<R, SV, DURBIN, PV><<<grid, wpc * 32, smem, s>>>(n, nsys, r, ldr, b, ldb, x, ldx, diag);     \
    }
    // smallest group of lanes that holds a system in S <= 4 register blocks per lane
    if (n <= 16) TB_LEV_GEN(1, 4)
    else if (n <= 32) TB_LEV_GEN(2, 4)
    else if (n <= 48) TB_LEV_GEN(3, 4)
    else if (n <= 64) TB_LEV_GEN(4, 4)
    else if (n <= 96) TB_LEV_GEN(3, 8)
    else if (n <= 128) TB_LEV_GEN(4, 8)
    else if (n <= 192) TB_LEV_GEN(3, 16)
    else if (n <= 256) TB_LEV_GEN(4, 16)
    else if (n <= 384) TB_LEV_GEN(3, 32)
    else TB_LEV_GEN(4, 32)
#undef TB_LEV_GEN
    return cudaGetLastError();
}

#define TB_INST(R, D) template cudaError_t launch_levinson_gen<R, D>(int64_t, int64_t, const R*, int64_t, const R*, int64_t, R*, int64_t, const R*, cudaStream_t);
TB_INST(float, false) TB_INST(float, true) TB_INST(double, false) TB_INST(double, true)
#undef TB_INST

}  // namespace tb
