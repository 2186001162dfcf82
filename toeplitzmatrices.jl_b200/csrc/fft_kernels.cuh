// Fused FFT-convolution kernels (sm_100a).  See DESIGN.md "Kernels".
//
//  k_rows  : contiguous L-point transforms, one FFT per NT threads.  Used (a) as the
//            whole matvec for embeddings that fit shared memory (load x with implicit
//            zero pad / Hankel reversal / two real columns packed as re+im -> DIF FFT ->
//            x spectrum in registers -> DIT inverse FFT -> alpha/beta/real-part store),
//            and (b) as the fused middle pass of the two-level transform.
//  k_colsA : first pass of the two-level transform for large N' = N1*N2: TA adjacent
//            strided columns per CTA, embedding fused into the load, inter-level
//            twiddle fused into the store.
//  k_colsC : last pass: conj twiddle fused into the load, inverse column FFTs,
//            epilogue y = alpha*maybereal(tmp)/N' + beta*y fused into the store.
//
// Replaces /root/reference/src/linearalgebra.jl:42-80 (mul! on a factorization),
// :225-242 (Circulant ldiv!) and the FFT inside factorize (:122-131 etc.).
#pragma once
#include "fft_core.cuh"

namespace tb {

enum IoMode : int { IO_REAL = 0, IO_CPLX = 1, IO_PAIR = 2, IO_SPEC = 3 };

template <typename R>
struct ConvArgs {
    // input side
    const void* x; long long ldx; int x_mode; long long n_in; int reverse_x;
    // output side
    void* y; long long ldy; int y_mode; long long m_out;
    long long ncols;          // number of user columns (PAIR mode: odd tail)
    long long nfft;           // number of FFT slots to run
    // spectrum (layout order); spec_rows rows of L; row = slot % spec_rows
    const cx<R>* spec; int spec_rows;
    int do_fwd, do_inv;
    R scale; cx<R> alpha, beta; int has_beta;
    const cx<R>* tw;
    // two-level only
    cx<R>* scratch;           // nfft * Np complex
    int log2n2;               // N2 = contiguous (row) length
    const cx<R>* tw_lo; const cx<R>* tw_hi; int tw_hb;   // w_Np^m = lo[m & mask] * hi[m >> hb]
    const cx<R>* tw_g;        // [E][N2]: w_Np^(delta_i * j2), delta_i = frequency offset of register i (thread-independent);
                              // the kernels read rows i = 2^b only and build the rest as subset products
    // column indirection (batched Krylov solvers: only the columns still iterating are transformed): user column of
    // logical column j is colmap[j]; NULL = identity.  x and y use the same map.
    const int* colmap;
    // per-column device scalars (batched Krylov: r -= alpha_j * A * uh fused into the epilogue): the effective alpha of
    // user column j is alpha * (col_alpha[j*stride], col_alpha[j*stride + 1]); NULL = none
    const double* col_alpha; int col_alpha_stride;
    // host side only: launch this pass with programmatic stream serialization (its CTAs may become resident while the
    // previous pass of the same product drains; they block in griddepcontrol.wait until that pass has completed)
    int pdl;
};

// Programmatic dependent launch (PDL).  Both are no-ops in a kernel that was launched normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Inter-level twiddles of one thread (fixed column j2): w^(k1(i) j2) = w^(kbase(t) j2) * w^(delta_i j2).
// One split-table lookup for the base, LOGE coalesced loads from the [E][N2] table, E-1 complex multiplies.
template <typename R, class F, bool GTAB = true>
__device__ __forceinline__ void level_twiddles(const ConvArgs<R>& p, int t, int j2, cx<R>* w);

// ---- per-slot input / output views: the embedding (zero pad, Hankel reversal), the packing of
// ---- two real columns as re + i*im, and the alpha/beta/real-part epilogue are fused here ----
__device__ __forceinline__ void cp_async_elem(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_elem(float* dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <typename R>
struct InView {
    const R* re; const R* im; const cx<R>* c; int mode; int n_in; int rev;
    __device__ __forceinline__ InView(const ConvArgs<R>& p, long long f) {
        mode = p.x_mode; n_in = (int)p.n_in; rev = p.reverse_x;
        re = nullptr; im = nullptr; c = nullptr;
        const int* cm = p.colmap;
        if (mode == IO_CPLX) c = reinterpret_cast<const cx<R>*>(p.x) + (cm ? (long long)cm[f] : f) * p.ldx;
        else if (mode == IO_REAL) re = reinterpret_cast<const R*>(p.x) + (cm ? (long long)cm[f] : f) * p.ldx;
        else if (mode == IO_PAIR) {
            const R* base = reinterpret_cast<const R*>(p.x);
            re = base + (cm ? (long long)cm[2 * f] : 2 * f) * p.ldx;
            im = (2 * f + 1 < p.ncols) ? base + (cm ? (long long)cm[2 * f + 1] : 2 * f + 1) * p.ldx : nullptr;
        }
    }
    // real modes only: start the asynchronous copy (LDGSTS) of logical element idx into the caller's own staging slots
    __device__ __forceinline__ void stage_real(int idx, R* dre, R* dim_) const {
        if (idx >= n_in) { *dre = 0; *dim_ = 0; return; }
        const int src = rev ? (n_in - 1 - idx) : idx;
        cp_async_elem(dre, re + src);
        if (im) cp_async_elem(dim_, im + src);
        else *dim_ = 0;
    }
    __device__ __forceinline__ cx<R> load(int idx) const {
        if (idx >= n_in) return mk<R>(0, 0);
        const int src = rev ? (n_in - 1 - idx) : idx;
        if (mode == IO_CPLX) return ld_stream(c + src);
        const R a = ld_stream(re + src);
        const R b = im ? ld_stream(im + src) : (R)0;
        return mk<R>(a, b);
    }
};

template <typename R>
struct OutView {
    R* re; R* im; cx<R>* c; int mode; int m_out; int has_beta; cx<R> as, beta; R as_im; bool unit;
    __device__ __forceinline__ OutView(const ConvArgs<R>& p, long long f) {
        mode = p.y_mode; m_out = (int)p.m_out; has_beta = p.has_beta; beta = p.beta;
        as = mk<R>(p.alpha.x * p.scale, p.alpha.y * p.scale);       // alpha/N' folded into one factor
        re = nullptr; im = nullptr; c = nullptr;
        const int* cm = p.colmap;
        long long j0 = 0, j1 = 0;
        if (mode == IO_CPLX) { j0 = cm ? (long long)cm[f] : f; c = reinterpret_cast<cx<R>*>(p.y) + j0 * p.ldy; }
        else if (mode == IO_REAL) { j0 = cm ? (long long)cm[f] : f; re = reinterpret_cast<R*>(p.y) + j0 * p.ldy; }
        else if (mode == IO_PAIR) {
            R* base = reinterpret_cast<R*>(p.y);
            j0 = cm ? (long long)cm[2 * f] : 2 * f;
            re = base + j0 * p.ldy;
            if (2 * f + 1 < p.ncols) { j1 = cm ? (long long)cm[2 * f + 1] : 2 * f + 1; im = base + j1 * p.ldy; }
        }
        as_im = as.x;
        if (p.col_alpha) {                                          // per-column device scalar times alpha/N'
            const double* a0 = p.col_alpha + j0 * p.col_alpha_stride;
            const cx<R> s0 = mk<R>((R)a0[0], (R)a0[1]);
            if (mode == IO_CPLX) as = cmul(as, s0);
            else {
                if (im) as_im = as.x * (R)p.col_alpha[j1 * p.col_alpha_stride];
                as.x *= s0.x;
            }
        }
        unit = !has_beta && as.x == (R)1 && as.y == (R)0 && !p.col_alpha;   // pass B of the two-level transform: plain stores
    }
    __device__ __forceinline__ void store(int idx, cx<R> v) const {
        if (idx >= m_out) return;
        if (mode == IO_CPLX) {
            if (unit) { c[idx] = v; return; }
            cx<R> o = cmul(as, v);
            if (has_beta) { const cx<R> b = cmul(beta, c[idx]); o.x += b.x; o.y += b.y; }
            c[idx] = o;
        } else {
            R o0 = as.x * v.x;
            if (has_beta) o0 = fma(beta.x, re[idx], o0);
            re[idx] = o0;
            if (im) {
                R o1 = as_im * v.y;
                if (has_beta) o1 = fma(beta.x, im[idx], o1);
                im[idx] = o1;
            }
        }
    }
};

// occupancy target: 128 registers/thread for Float64, 64 for Float32, limited by shared memory
template <typename R>
constexpr int min_ctas(int threads, int smem_bytes) {
    int by_regs = 65536 / (threads * 2 * ((sizeof(R) == 8 ? 4 : 2) << le_of<R>()));   // budget: 2x the data registers
    int by_smem = (220 * 1024) / (smem_bytes > 0 ? smem_bytes : 1);
    int m = by_regs < by_smem ? by_regs : by_smem;
    return m < 1 ? 1 : m;
}

// ---- contiguous transforms -------------------------------------------------------------
template <typename R, int LOG2L, int TS>
__global__ void __launch_bounds__((1 << (LOG2L - le_of<R>())) * TS, min_ctas<R>((1 << (LOG2L - le_of<R>())) * TS, (int)sizeof(cx<R>) * (1 << LOG2L) * TS))
k_rows(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L>;
    constexpr int L = F::L, NT = F::NT, E = F::E;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int slot = threadIdx.x / NT, t = threadIdx.x % NT;
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw) + slot * L;
    RowMap<R> map;
    pdl_wait();                    // pass B of a PDL chain: the workspace is complete from here on
    pdl_launch_dependents();

    for (long long f0 = (long long)blockIdx.x * TS; f0 < p.nfft; f0 += (long long)gridDim.x * TS) {
        const long long f = f0 + slot;
        const bool live = f < p.nfft;
        const long long fc = live ? f : 0;          // dead slots run on slot 0's data, never store
        cx<R> a[E];
        const cx<R>* srow = p.spec ? p.spec + (fc % p.spec_rows) * (long long)L + t : nullptr;
        if (p.x_mode == IO_SPEC) {
            const cx<R>* xin = reinterpret_cast<const cx<R>*>(p.x) + fc * (long long)L + t;
#pragma unroll
            for (int i = 0; i < E; ++i) a[i] = ld_stream(xin + i * NT);
        } else {
            const InView<R> in(p, fc);
#pragma unroll
            for (int c = 0; c < E; ++c) a[c] = in.load(t + c * NT);
        }
        if (p.do_fwd) F::forward(a, t, sm, map, p.tw);
        if (srow) {
#pragma unroll
            for (int i = 0; i < E; ++i) a[i] = cmul(a[i], ld_stream(srow + i * NT));
        }
        if (p.do_inv) F::inverse(a, t, sm, map, p.tw);
        if (live) {
            if (p.y_mode == IO_SPEC) {
                cx<R>* o = reinterpret_cast<cx<R>*>(p.y) + f * (long long)L + t;
#pragma unroll
                for (int i = 0; i < E; ++i) o[i * NT] = a[i];
            } else {
                const OutView<R> out(p, f);
#pragma unroll
                for (int j = 0; j < E; ++j) out.store(t + j * NT, a[j]);
            }
        }
        __syncthreads();
    }
}

// ---- two-level, strided passes -----------------------------------------------------------
template <typename R>
__device__ __forceinline__ cx<R> level_twiddle(const ConvArgs<R>& p, unsigned m) {
    const cx<R> lo = __ldg(p.tw_lo + (m & ((1u << p.tw_hb) - 1u)));
    const cx<R> hi = __ldg(p.tw_hi + (m >> p.tw_hb));
    return cmul(lo, hi);
}

// GTAB = false: outer pass of a three-level plan, where N2 is far too long for the [E][N2] table -- every factor comes
// from the split tables.  A compile-time choice: as a run-time branch it cost the two-level kernels 56 bytes of spills
// and 6 % of the headline rate.
template <typename R, class F, bool GTAB>
__device__ __forceinline__ void level_twiddles(const ConvArgs<R>& p, int t, int j2, cx<R>* w) {
    if constexpr (!GTAB) {
#pragma unroll
        for (int i = 0; i < F::E; ++i) w[i] = level_twiddle<R>(p, (unsigned)F::freq_of_pos(F::pos_of_reg(i, t)) * (unsigned)j2);
        return;
    }
    const unsigned kbase = (unsigned)F::freq_of_pos(F::pos_of_reg(0, t));
    const cx<R> base = level_twiddle<R>(p, kbase * (unsigned)j2);
    // delta_i is linear in the bits of i (pos_of_reg and freq_of_pos move disjoint bit fields), so the E factors
    // w^(delta_i j2) are the subset products of LOGE generators: LOGE table loads and E-1 multiplies instead of E loads
    // and E multiplies -- the table rows were as much L2 traffic as the tile itself
    const cx<R>* g = p.tw_g + j2;
    w[0] = base;
#pragma unroll
    for (int b = 0; b < F::LOGE; ++b) {
        const cx<R> gb = __ldg(g + ((long long)(1 << b) << p.log2n2));
#pragma unroll
        for (int i = (1 << b); i < (2 << b); ++i) w[i] = cmul(w[i - (1 << b)], gb);
    }
}

// grid.x = (N2/TA) * nfft ; block = TA * NT1.  HALF: the input occupies at most the first half of the embedding
// (n <= N'/2, every Toeplitz/Hankel product), so rows >= N1/2 are padding: not loaded, first butterfly level pruned.
template <typename R, int LOG2L1, int TA, bool HALF, bool GTAB = true>
__global__ void __launch_bounds__((1 << (LOG2L1 - le_of<R>())) * TA, min_ctas<R>((1 << (LOG2L1 - le_of<R>())) * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsA(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L1>;
    constexpr int NT = F::NT, E = F::E;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int l2 = p.log2n2;
    const unsigned tiles = (1u << l2) / TA;
    const long long f = blockIdx.x / tiles;
    const int j2 = (int)(blockIdx.x % tiles) * TA + q;
    ColMap<TA> map{q};
    pdl_launch_dependents();
    cx<R> a[E];
    {
        const InView<R> in(p, f);
#pragma unroll
        for (int c = 0; c < (HALF ? E / 2 : E); ++c) a[c] = in.load(((t + c * NT) << l2) + j2);
    }
    F::template forward<ColMap<TA>, false, HALF>(a, t, sm, map, p.tw);
    cx<R>* out = p.scratch + (f << (l2 + LOG2L1)) + j2;
    cx<R> w[E];
    level_twiddles<R, F, GTAB>(p, t, j2, w);
#pragma unroll
    for (int i = 0; i < E; ++i) out[(long long)F::pos_of_reg(i, t) << l2] = cmul(a[i], w[i]);
}

// HALF: only the first half of the embedding is wanted (m <= N'/2): outputs j >= E/2 are never stored and the
// compiler drops the second half of the last butterfly level.
template <typename R, int LOG2L1, int TA, bool HALF, bool GTAB = true>
__global__ void __launch_bounds__((1 << (LOG2L1 - le_of<R>())) * TA, min_ctas<R>((1 << (LOG2L1 - le_of<R>())) * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsC(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L1>;
    constexpr int NT = F::NT, E = F::E;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int l2 = p.log2n2;
    const unsigned tiles = (1u << l2) / TA;
    const long long f = blockIdx.x / tiles;
    const int j2 = (int)(blockIdx.x % tiles) * TA + q;
    ColMap<TA> map{q};
    const cx<R>* in = p.scratch + (f << (l2 + LOG2L1)) + j2;
    cx<R> a[E];
    {
        cx<R> w[E];
        level_twiddles<R, F, GTAB>(p, t, j2, w);      // tables only: may run before the previous pass has finished
        pdl_wait();
#pragma unroll
        for (int i = 0; i < E; ++i) a[i] = cmulc(ld_stream(in + ((long long)F::pos_of_reg(i, t) << l2)), w[i]);
    }
    F::inverse(a, t, sm, map, p.tw);
    const OutView<R> out(p, f);
#pragma unroll
    for (int j = 0; j < (HALF ? E / 2 : E); ++j) out.store(((t + j * NT) << l2) + j2, a[j]);
}

}  // namespace tb
