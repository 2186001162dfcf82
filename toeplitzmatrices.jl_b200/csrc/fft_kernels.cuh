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
};

// ---- element load / store with the embedding, packing and epilogue fused ---------------
template <typename R>
__device__ __forceinline__ cx<R> load_elem(const void* x, long long ldx, int mode, long long f, long long idx,
                                           long long n_in, int rev, long long ncols) {
    if (idx >= n_in) return mk<R>(0, 0);
    const long long src = rev ? (n_in - 1 - idx) : idx;
    if (mode == IO_CPLX) return __ldg(reinterpret_cast<const cx<R>*>(x) + f * ldx + src);
    if (mode == IO_REAL) return mk<R>(__ldg(reinterpret_cast<const R*>(x) + f * ldx + src), 0);
    // IO_PAIR: columns 2f, 2f+1
    const R* p = reinterpret_cast<const R*>(x) + (2 * f) * ldx + src;
    R re = __ldg(p);
    R im = (2 * f + 1 < ncols) ? __ldg(p + ldx) : (R)0;
    return mk<R>(re, im);
}

template <typename R>
__device__ __forceinline__ void store_elem(const ConvArgs<R>& p, long long f, long long idx, cx<R> v) {
    if (idx >= p.m_out) return;
    v.x *= p.scale; v.y *= p.scale;
    if (p.y_mode == IO_CPLX) {
        cx<R>* q = reinterpret_cast<cx<R>*>(p.y) + f * p.ldy + idx;
        cx<R> o = cmul(p.alpha, v);
        if (p.has_beta) { cx<R> b = cmul(p.beta, *q); o.x += b.x; o.y += b.y; }
        *q = o;
    } else if (p.y_mode == IO_REAL) {
        R* q = reinterpret_cast<R*>(p.y) + f * p.ldy + idx;
        R o = p.alpha.x * v.x;
        if (p.has_beta) o = fma(p.beta.x, *q, o);
        *q = o;
    } else {  // IO_PAIR
        R* q = reinterpret_cast<R*>(p.y) + (2 * f) * p.ldy + idx;
        R o0 = p.alpha.x * v.x;
        if (p.has_beta) o0 = fma(p.beta.x, *q, o0);
        *q = o0;
        if (2 * f + 1 < p.ncols) {
            R o1 = p.alpha.x * v.y;
            if (p.has_beta) o1 = fma(p.beta.x, q[p.ldy], o1);
            q[p.ldy] = o1;
        }
    }
}

// ---- contiguous transforms -------------------------------------------------------------
// occupancy target: 128 registers/thread for Float64, 64 for Float32, limited by shared memory
template <typename R>
constexpr int min_ctas(int threads, int smem_bytes) {
    int by_regs = 65536 / (threads * (sizeof(R) == 8 ? 128 : 64));
    int by_smem = (220 * 1024) / (smem_bytes > 0 ? smem_bytes : 1);
    int m = by_regs < by_smem ? by_regs : by_smem;
    return m < 1 ? 1 : m;
}

template <typename R, int LOG2L, int TS>
__global__ void __launch_bounds__((1 << LOG2L) / E * TS, min_ctas<R>((1 << LOG2L) / E * TS, (int)sizeof(cx<R>) * (1 << LOG2L) * TS))
k_rows(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L>;
    constexpr int L = F::L, NT = F::NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int slot = threadIdx.x / NT, t = threadIdx.x % NT;
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw) + slot * L;
    RowMap<R> map;

    for (long long f0 = (long long)blockIdx.x * TS; f0 < p.nfft; f0 += (long long)gridDim.x * TS) {
        const long long f = f0 + slot;
        const bool live = f < p.nfft;
        const long long fc = live ? f : 0;          // dead slots run on slot 0's data, never store
        cx<R> a[E];
        const cx<R>* srow = p.spec ? p.spec + (fc % p.spec_rows) * (long long)L : nullptr;
        if (p.x_mode == IO_SPEC) {
#pragma unroll
            for (int i = 0; i < E; ++i) a[i] = __ldg(reinterpret_cast<const cx<R>*>(p.x) + fc * (long long)L + i * NT + t);
        } else {
#pragma unroll
            for (int c = 0; c < E; ++c)
                a[c] = load_elem<R>(p.x, p.ldx, p.x_mode, fc, t + c * NT, p.n_in, p.reverse_x, p.ncols);
        }
        if (p.do_fwd) F::forward(a, t, sm, map, p.tw);
        if (srow) {
#pragma unroll
            for (int i = 0; i < E; ++i) a[i] = cmul(a[i], __ldg(srow + i * NT + t));
        }
        if (p.do_inv) F::inverse(a, t, sm, map, p.tw);
        if (live) {
            if (p.y_mode == IO_SPEC) {
                cx<R>* o = reinterpret_cast<cx<R>*>(p.y) + f * (long long)L;
#pragma unroll
                for (int i = 0; i < E; ++i) o[i * NT + t] = a[i];
            } else {
#pragma unroll
                for (int j = 0; j < E; ++j) store_elem<R>(p, f, t + j * NT, a[j]);
            }
        }
        __syncthreads();
    }
}

// ---- two-level, strided passes -----------------------------------------------------------
template <typename R>
__device__ __forceinline__ cx<R> level_twiddle(const ConvArgs<R>& p, long long m) {
    const cx<R> lo = __ldg(p.tw_lo + (m & ((1LL << p.tw_hb) - 1)));
    const cx<R> hi = __ldg(p.tw_hi + (m >> p.tw_hb));
    return cmul(lo, hi);
}

// grid.x = (N2/TA) * nfft ; block = TA * NT1
template <typename R, int LOG2L1, int TA>
__global__ void __launch_bounds__((1 << LOG2L1) / E * TA, min_ctas<R>((1 << LOG2L1) / E * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsA(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L1>;
    constexpr int NT = F::NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const long long N2 = 1LL << p.log2n2;
    const long long tiles = N2 / TA;
    const long long f = blockIdx.x / tiles;
    const long long j2 = (blockIdx.x % tiles) * TA + q;
    ColMap<TA> map{q};
    cx<R> a[E];
#pragma unroll
    for (int c = 0; c < E; ++c) {
        const long long idx = (long long)(t + c * NT) * N2 + j2;
        a[c] = load_elem<R>(p.x, p.ldx, p.x_mode, f, idx, p.n_in, p.reverse_x, p.ncols);
    }
    F::forward(a, t, sm, map, p.tw);
    cx<R>* out = p.scratch + f * (N2 << LOG2L1);
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int P = F::pos_of_reg(i, t);
        const long long k1 = F::freq_of_pos(P);
        out[(long long)P * N2 + j2] = cmul(a[i], level_twiddle<R>(p, k1 * j2));
    }
}

template <typename R, int LOG2L1, int TA>
__global__ void __launch_bounds__((1 << LOG2L1) / E * TA, min_ctas<R>((1 << LOG2L1) / E * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsC(const ConvArgs<R> p) {
    using F = SlotFFT<R, LOG2L1>;
    constexpr int NT = F::NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const long long N2 = 1LL << p.log2n2;
    const long long tiles = N2 / TA;
    const long long f = blockIdx.x / tiles;
    const long long j2 = (blockIdx.x % tiles) * TA + q;
    ColMap<TA> map{q};
    const cx<R>* in = p.scratch + f * (N2 << LOG2L1);
    cx<R> a[E];
#pragma unroll
    for (int i = 0; i < E; ++i) {
        const int P = F::pos_of_reg(i, t);
        const long long k1 = F::freq_of_pos(P);
        a[i] = cmulc(in[(long long)P * N2 + j2], level_twiddle<R>(p, k1 * j2));
    }
    F::inverse(a, t, sm, map, p.tw);
#pragma unroll
    for (int j = 0; j < E; ++j) store_elem<R>(p, f, (long long)(t + j * NT) * N2 + j2, a[j]);
}

}  // namespace tb
