// Shared-memory radix-16 FFT building blocks for sm_100a (hand-written; no cuFFT).
//
// Design: an L-point complex FFT (L = 2^LOG2L) is done by NT = L/16 threads, 16
// points per thread in registers.  Forward transforms are decimation-in-frequency
// (natural order in -> digit-reversed order out), inverse transforms are the exact
// mirror (decimation-in-time, digit-reversed in -> natural out), so a
// forward -> pointwise spectrum multiply -> inverse chain needs no re-ordering pass
// at all and the multiply happens in registers between the two halves.  This is
// what replaces the reference's `mul!(tmp, dft, tmp); tmp .*= vcvr_dft; dft \ tmp`
// (/root/reference/src/linearalgebra.jl:63-67).
//
// Shared memory is used in place (a thread reads and writes the same 16 slots in a
// stage), one __syncthreads per stage, with an XOR swizzle that makes every stage
// bank-conflict free for 16-byte (complex<double>) and 8-byte (complex<float>)
// elements (verified by brute force over all stages).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tb {

// Points per thread (log2).  Radix-16 threads for both precisions: fewer shared-memory round trips
// (the LSU is the tightest resource); radix-8 Float64 threads (64 registers, 32 warps/SM) were measured
// 4 % slower once the stage twiddles are generated in registers (profiles/experiments_r01.md).
#ifndef TB_LE_F64
#define TB_LE_F64 4
#endif
#ifndef TB_LE_F32
#define TB_LE_F32 4
#endif
template <typename R> __host__ __device__ constexpr int le_of() { return sizeof(R) == 8 ? TB_LE_F64 : TB_LE_F32; }
constexpr int MAX_LOG2L = 14;    // longest single-pass FFT (16384 points, Float32 only)

template <typename R> struct CxT;
template <> struct CxT<float> { using type = float2; };
template <> struct CxT<double> { using type = double2; };
template <typename R> using cx = typename CxT<R>::type;

template <typename R> __host__ __device__ __forceinline__ cx<R> mk(R a, R b) { cx<R> r; r.x = a; r.y = b; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { C r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { C r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
// a*b
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r;
    r.x = fma(a.x, b.x, -(a.y * b.y));
    r.y = fma(a.x, b.y, a.y * b.x);
    return r;
}
// a*conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r;
    r.x = fma(a.x, b.x, a.y * b.y);
    r.y = fma(a.y, b.x, -(a.x * b.y));
    return r;
}

__host__ __device__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}
__host__ __device__ constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }

// Per-length stage twiddle table (E = 2^le points per thread): stage k (stride
// s = L/E^(k+1) > 1) owns the block
//   T[off_k + (c-1)*s + r] = exp(-2 pi i r c / (E s)),  c = 1..E-1, r = 0..s-1
// so a warp reads consecutive entries (thread order) and late stages touch only a few
// hundred bytes that stay in L1.
__host__ __device__ constexpr int stage_tw_off(int log2l, int k, int le) {
    int off = 0;
    for (int kk = 0; kk < k; ++kk) {
        int ls = log2l - le * (kk + 1);
        if (ls > 0) off += ((1 << le) - 1) << ls;
    }
    return off;
}
__host__ __device__ constexpr int stage_tw_total(int log2l, int le) { return stage_tw_off(log2l, log2l / le, le); }

// d * exp(-/+ 2 pi i k/16); k is a compile-time constant after unrolling.
template <typename R, bool INV>
__device__ __forceinline__ cx<R> mul_w16(cx<R> d, int k) {
    const R C1 = (R)0.92387953251128675613, S1 = (R)0.38268343236508977173;
    const R C2 = (R)0.70710678118654752440;
    switch (k) {
    case 0: return d;
    case 4: return INV ? mk<R>(-d.y, d.x) : mk<R>(d.y, -d.x);
    case 2: return INV ? mk<R>((d.x - d.y) * C2, (d.x + d.y) * C2) : mk<R>((d.x + d.y) * C2, (d.y - d.x) * C2);
    case 6: return INV ? mk<R>(-(d.x + d.y) * C2, (d.x - d.y) * C2) : mk<R>((d.y - d.x) * C2, -(d.x + d.y) * C2);
    default: {
        R wr = (k == 1) ? C1 : (k == 3) ? S1 : (k == 5) ? -S1 : -C1;
        R ws = (k == 1 || k == 7) ? S1 : C1;     // sin(2 pi k/16) > 0 for k in 1..7
        R wi = INV ? ws : -ws;
        return mk<R>(fma(d.x, wr, -(d.y * wi)), fma(d.x, wi, d.y * wr));
    }
    }
}

// ---- fused-multiply-add butterflies (TB_FMA_BFLY) -------------------------------------------------------------
// u' = u + w v, v' = u - w v with w = exp(-/+ 2 pi i k/16): the twiddle rides the additions.  General w: the sum in four
// FMAs and the difference as 2u - (u + w v) in two more (six instructions instead of 4 + 4); w = (+-1 +- i)/sqrt(2): two
// additions and four FMAs (instead of 2 + 2 + 4).  Both register transforms use the decimation-in-time order (bit-
// reversed in, natural out) so that every twiddle sits in front of a butterfly; the forward transform's natural-in /
// bit-reversed-out contract is kept by renaming registers, which costs nothing once the loops are unrolled.
#ifndef TB_FMA_BFLY
#define TB_FMA_BFLY 1
#endif
template <typename R, bool INV>
__device__ __forceinline__ void bfly_w16(cx<R>& u, cx<R>& v, int k) {
    const R C1 = (R)0.92387953251128675613, S1 = (R)0.38268343236508977173;
    const R C2 = (R)0.70710678118654752440;
    switch (k) {
    case 0: { const cx<R> s = cadd(u, v); v = csub(u, v); u = s; return; }
    case 4: {
        const cx<R> t = INV ? mk<R>(-v.y, v.x) : mk<R>(v.y, -v.x);
        const cx<R> s = cadd(u, t); v = csub(u, t); u = s; return;
    }
    case 2: case 6: {
        // w v = C2 * (p, q):  k = 2: fwd (vx + vy, vy - vx), inv (vx - vy, vx + vy);  k = 6: fwd (vy - vx, -(vx + vy)), inv (-(vx + vy), vx - vy)
        const R sum = v.x + v.y, dif = v.x - v.y;
        R pp, qq;
        if (k == 2) { pp = INV ? dif : sum; qq = INV ? sum : -dif; }
        else        { pp = INV ? -sum : -dif; qq = INV ? dif : -sum; }
        const cx<R> s = mk<R>(fma(C2, pp, u.x), fma(C2, qq, u.y));
        v = mk<R>(fma(-C2, pp, u.x), fma(-C2, qq, u.y));
        u = s;
        return;
    }
    default: {
        const R wr = (k == 1) ? C1 : (k == 3) ? S1 : (k == 5) ? -S1 : -C1;
        const R ws = (k == 1 || k == 7) ? S1 : C1;
        const R wi = INV ? ws : -ws;
        const cx<R> s = mk<R>(fma(v.x, wr, fma(-v.y, wi, u.x)), fma(v.x, wi, fma(v.y, wr, u.y)));
        v = mk<R>(fma((R)2, u.x, -s.x), fma((R)2, u.y, -s.y));
        u = s;
        return;
    }
    }
}
// a[brev(c)] = x[c] in -> natural out; INV selects the conjugate twiddles
template <typename R, int RAD, bool INV, int H0 = 1>
__device__ __forceinline__ void dit_fused(cx<R>* a) {
#pragma unroll
    for (int h = H0; h <= RAD / 2; h <<= 1) {
#pragma unroll
        for (int blk = 0; blk < RAD; blk += 2 * h) {
#pragma unroll
            for (int i = 0; i < h; ++i) bfly_w16<R, INV>(a[blk + i], a[blk + i + h], i * (8 / h));
        }
    }
}

// Inverse stage with its twiddles: element c (held in a[brev(c)]) is multiplied by conj(w[c]), c >= 1, and the first
// decimation-in-time level pairs elements c and c + RAD/2 -- one product, then the sum in four FMAs and the difference
// as 2p - s (ten instructions per pair instead of twelve).
template <typename R, int RAD>
__device__ __forceinline__ void dit_regs_tw(cx<R>* a, const cx<R>* w) {
    constexpr int LG = ilog2c(RAD);
#pragma unroll
    for (int m = 0; m < RAD / 2; ++m) {
        const int c0 = brev(2 * m, LG), c1 = brev(2 * m + 1, LG);
        const cx<R> p = (c0 == 0) ? a[2 * m] : cmulc(a[2 * m], w[c0]);
        const cx<R> v = a[2 * m + 1], ww = w[c1];
        const cx<R> sum = mk<R>(fma(v.x, ww.x, fma(v.y, ww.y, p.x)), fma(v.y, ww.x, fma(-v.x, ww.y, p.y)));
        a[2 * m + 1] = mk<R>(fma((R)2, p.x, -sum.x), fma((R)2, p.y, -sum.y));
        a[2 * m] = sum;
    }
    dit_fused<R, RAD, true, 2>(a);
}

// In-register radix-RAD DIF butterfly: natural in -> a[brev(c)] = X[c].
template <typename R, int RAD>
__device__ __forceinline__ void dif_regs(cx<R>* a) {
#if TB_FMA_BFLY
    {
        cx<R> b[RAD];
#pragma unroll
        for (int i = 0; i < RAD; ++i) b[brev(i, ilog2c(RAD))] = a[i];
        dit_fused<R, RAD, false>(b);
#pragma unroll
        for (int c = 0; c < RAD; ++c) a[brev(c, ilog2c(RAD))] = b[c];
        return;
    }
#endif
#pragma unroll
    for (int h = RAD / 2; h >= 1; h >>= 1) {
#pragma unroll
        for (int blk = 0; blk < RAD; blk += 2 * h) {
#pragma unroll
            for (int i = 0; i < h; ++i) {
                cx<R> u = a[blk + i], v = a[blk + i + h];
                a[blk + i] = cadd(u, v);
                a[blk + i + h] = mul_w16<R, false>(csub(u, v), i * (8 / h));
            }
        }
    }
}
// Same with a[RAD/2 .. RAD) == 0 on entry (the zero padding of a circulant embedding, src/linearalgebra.jl:57-62):
// the first butterfly level degenerates to a[i + RAD/2] = a[i] * w^i and costs no additions.
template <typename R, int RAD>
__device__ __forceinline__ void dif_regs_halfzero(cx<R>* a) {
    static_assert(RAD == 16, "half-zero variant is written for radix-16 threads");
#if TB_FMA_BFLY
    {
        // bit-reversed order puts the zero half on the odd positions: the first decimation-in-time level is a copy
        cx<R> b[RAD];
#pragma unroll
        for (int i = 0; i < RAD / 2; ++i) { b[brev(i, 4)] = a[i]; b[brev(i, 4) + 1] = a[i]; }
        dit_fused<R, RAD, false, 2>(b);
#pragma unroll
        for (int c = 0; c < RAD; ++c) a[brev(c, 4)] = b[c];
        return;
    }
#endif
#pragma unroll
    for (int i = 0; i < RAD / 2; ++i) a[i + RAD / 2] = mul_w16<R, false>(a[i], i);
#pragma unroll
    for (int h = RAD / 4; h >= 1; h >>= 1) {
#pragma unroll
        for (int blk = 0; blk < RAD; blk += 2 * h) {
#pragma unroll
            for (int i = 0; i < h; ++i) {
                cx<R> u = a[blk + i], v = a[blk + i + h];
                a[blk + i] = cadd(u, v);
                a[blk + i + h] = mul_w16<R, false>(csub(u, v), i * (8 / h));
            }
        }
    }
}
// Mirror: a[brev(c)] = X[c] in -> natural out, conjugate twiddles (unnormalised inverse).
template <typename R, int RAD>
__device__ __forceinline__ void dit_regs(cx<R>* a) {
#if TB_FMA_BFLY
    dit_fused<R, RAD, true>(a);
    return;
#endif
#pragma unroll
    for (int h = 1; h <= RAD / 2; h <<= 1) {
#pragma unroll
        for (int blk = 0; blk < RAD; blk += 2 * h) {
#pragma unroll
            for (int i = 0; i < h; ++i) {
                cx<R> u = a[blk + i], v = mul_w16<R, true>(a[blk + i + h], i * (8 / h));
                a[blk + i] = cadd(u, v);
                a[blk + i + h] = csub(u, v);
            }
        }
    }
}

// Bank-conflict-free XOR swizzles for the contiguous ("row") layout: address = i ^ f(i).
// f is linear over XOR, and inside a stage an element index is base | (c << ls) with
// disjoint bits, so address = (base ^ f(base)) ^ ((c << ls) ^ f(c << ls)): one per-thread
// term plus a compile-time constant per element (a plain immediate offset when f(c << ls) == 0).
template <typename R> __host__ __device__ constexpr int swz_f(int i);
template <> __host__ __device__ constexpr int swz_f<double>(int i) { return ((i >> 3) ^ (i >> 4)) & 7; }
template <> __host__ __device__ constexpr int swz_f<float>(int i) { return (i >> 4) & 15; }

template <typename R> struct RowMap {       // one FFT contiguous in shared memory
    static constexpr int TSHIFT = 0;        // FFT thread t sits at bit 0 of the CTA thread index
    __device__ __forceinline__ int pre(int base) const { return base ^ swz_f<R>(base); }
    __device__ __forceinline__ int at(int pre_, int off) const {
        constexpr int FIELD = sizeof(R) == 8 ? 7 : 15;     // bits the swizzle may flip
        return (swz_f<R>(off) == 0 && (off & FIELD) == 0) ? pre_ + off : pre_ ^ (off ^ swz_f<R>(off));
    }
};
template <int TA> struct ColMap {           // TA FFTs interleaved element-wise
    static constexpr int TSHIFT = ilog2c(TA);    // thread index = t * TA + q
    int q;
    __device__ __forceinline__ int pre(int base) const { return base * TA + q; }
    __device__ __forceinline__ int at(int pre_, int off) const { return pre_ + off * TA; }
};

template <typename R> __device__ __forceinline__ cx<R> ldg_cx(const cx<R>* p) { return __ldg(p); }

// Stage twiddles of one thread: w[c] = exp(-2 pi i r c / (E s)), c = 1..E-1.  The load/store unit is the
// tightest resource of the Float64 transforms (profiles/experiments_r01.md), so for Float64 only w[1] is
// loaded and the other powers are built by a depth-<=4 product tree in the (less loaded) FMA pipe
// (about 4 extra roundings per twiddle: ~1e-15 relative in Float64, ~3e-7 in Float32 -- inside the
// 1e-12 / 1e-5 parity budgets, measured by the GPU parity tests).  TB_TWGEN_*=0 restores full tables.
#ifndef TB_TWGEN_F64
#define TB_TWGEN_F64 1
#endif
#ifndef TB_TWGEN_F32
#define TB_TWGEN_F32 1
#endif
template <typename R, int E>
__device__ __forceinline__ void stage_twiddles_of_thread(const cx<R>* __restrict__ blk, int s, int r, cx<R>* w) {
    if constexpr ((sizeof(R) == 8 && TB_TWGEN_F64) || (sizeof(R) == 4 && TB_TWGEN_F32)) {
        w[1] = ldg_cx<R>(blk + r);
#pragma unroll
        for (int c = 2; c < E; ++c) w[c] = cmul(w[c / 2], w[c - c / 2]);
    } else {
#pragma unroll
        for (int c = 1; c < E; ++c) w[c] = ldg_cx<R>(blk + (c - 1) * s + r);
    }
}

// Streaming loads that do not allocate in L1, so the only L1-resident global data are the
// stage twiddle tables (re-read by every row a CTA processes).
__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double r;
    asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float r;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

#define TB_SM_LOAD(dst, src) dst = src
#define TB_SM_STORE(dst, src) dst = src

// Barrier between the stores of one stage and the loads of the next.  An element's owner is its index with the
// stage's digit removed, so the threads that exchange data differ only in bits [lo, lo + LOGE) of t, lo = the smaller
// of the two strides' log2 (forward: the next stage's, inverse: the storing stage's).  When that field (shifted by the map's thread
// layout) lies inside the lane index, the exchange never leaves a warp and __syncwarp is enough -- two of the four
// exchanges of the 4096-point row pass.
template <int LOGE, int TSHIFT>
__device__ __forceinline__ void exchange_barrier(int lo) {
    if (lo + LOGE + TSHIFT <= 5) __syncwarp();
    else __syncthreads();
}

// One L-point FFT held by NT threads (t = 0..NT-1), E = 2^LE complex registers each.
template <typename R, int LOG2L, int LE = le_of<R>()>
struct SlotFFT {
    static constexpr int LOGE = LE;
    static constexpr int E = 1 << LE;
    static_assert(LOG2L >= LOGE, "FFT shorter than the per-thread radix");
    static constexpr int L = 1 << LOG2L;
    static constexpr int NT = L / E;
    static constexpr int NFULL = LOG2L / LOGE;
    static constexpr int REM = LOG2L % LOGE;
    static constexpr int RR = 1 << REM;

    // A one-bit remainder (L = E^k * 2) is not given a shared-memory exchange of its own: the last radix-2 level pairs
    // positions P and P^1, which after the last full stage sit in threads t and t^1 -- the same warp -- so it is done
    // with one shuffle per register (half the shared-memory-pipe wavefronts of a store + load, and no barrier).
    static constexpr bool REM_SHFL = (REM == 1);
    // position (in the digit-reversed output order) of register i of thread t
    __host__ __device__ static constexpr int pos_of_reg(int i, int t) {
        return REM == 0 ? t * E + brev(i, LOGE)
             : REM_SHFL ? ((t >> 1) << (1 + LOGE)) + (t & 1) + (brev(i, LOGE) << 1)
                        : (t + (i / RR) * NT) * RR + brev(i % RR, REM);
    }
    // the cross-lane radix-2 level (its own inverse up to the factor 2 the caller's 1/N carries)
    template <class Map>
    __device__ __forceinline__ static void rem_shuffle(cx<R>* a, int t) {
        constexpr int LM = 1 << Map::TSHIFT;             // lane distance of thread t ^ 1
        const bool odd = (t & 1) != 0;
#pragma unroll
        for (int i = 0; i < E; ++i) {
            cx<R> v;
            v.x = __shfl_xor_sync(0xffffffffu, a[i].x, LM);
            v.y = __shfl_xor_sync(0xffffffffu, a[i].y, LM);
            a[i] = odd ? csub(v, a[i]) : cadd(a[i], v);
        }
    }
    // frequency index held at position P
    __host__ __device__ static constexpr int freq_of_pos(int P) {
        int k = 0, mult = 1;
        for (int s = 0; s < NFULL; ++s) {
            k += ((P >> (LOG2L - LOGE * (s + 1))) & (E - 1)) * mult;
            mult <<= LOGE;
        }
        if (REM > 0) k += (P & (RR - 1)) * mult;
        return k;
    }

    // On entry a[c] = x[t + c*NT] (or, with FROM_SMEM, the natural-order input already sits in shared
    // memory at sm[map(idx)]).  On exit register i holds X[freq_of_pos(pos_of_reg(i,t))].
    // HALF_IN: a[E/2 .. E) are known zeros (upper half of the input is padding) and need not be initialised.
    template <class Map, bool FROM_SMEM = false, bool HALF_IN = false>
    __device__ __forceinline__ static void forward(cx<R>* a, int t, cx<R>* sm, Map map,
                                                   const cx<R>* __restrict__ tw) {
#pragma unroll
        for (int k = 0; k < NFULL; ++k) {
            const int ls = LOG2L - LOGE * (k + 1);       // log2 of the stride
            const int s = 1 << ls;
            const int g = t >> ls, r = t & (s - 1);
            const int pre = map.pre((g << (ls + LOGE)) + r);
            if (k > 0 || FROM_SMEM) {
#pragma unroll
                for (int c = 0; c < E; ++c) TB_SM_LOAD(a[c], sm[map.at(pre, c << ls)]);
            }
            if constexpr (HALF_IN && !FROM_SMEM && E == 16) {
                if (k == 0) dif_regs_halfzero<R, E>(a);
                else dif_regs<R, E>(a);
            } else {
                dif_regs<R, E>(a);
            }
            if (ls > 0) {
                cx<R> w[E];
                stage_twiddles_of_thread<R, E>(tw + stage_tw_off(LOG2L, k, LOGE), s, r, w);
#pragma unroll
                for (int c = 1; c < E; ++c) a[brev(c, LOGE)] = cmul(a[brev(c, LOGE)], w[c]);
            }
            const bool last = (k == NFULL - 1) && (REM == 0 || REM_SHFL);
            if (!last) {
#pragma unroll
                for (int c = 0; c < E; ++c) TB_SM_STORE(sm[map.at(pre, c << ls)], a[brev(c, LOGE)]);
                if (k + 1 < NFULL) exchange_barrier<LOGE, Map::TSHIFT>(ls - LOGE);
                else __syncthreads();
            }
        }
        if constexpr (REM_SHFL) {
            rem_shuffle<Map>(a, t);
        } else if (REM > 0) {
#pragma unroll
            for (int u = 0; u < E / RR; ++u) {
                const int pre = map.pre((t + u * NT) * RR);
#pragma unroll
                for (int c = 0; c < RR; ++c) TB_SM_LOAD(a[u * RR + c], sm[map.at(pre, c)]);
            }
#pragma unroll
            for (int u = 0; u < E / RR; ++u) dif_regs<R, RR>(a + u * RR);
        }
    }

    // Mirror of forward().  On exit a[j] = y[t + j*NT] (unnormalised).
    template <class Map>
    __device__ __forceinline__ static void inverse(cx<R>* a, int t, cx<R>* sm, Map map,
                                                   const cx<R>* __restrict__ tw) {
        if constexpr (REM_SHFL) {
            rem_shuffle<Map>(a, t);
        } else if (REM > 0) {
#pragma unroll
            for (int u = 0; u < E / RR; ++u) dit_regs<R, RR>(a + u * RR);
#pragma unroll
            for (int u = 0; u < E / RR; ++u) {
                const int pre = map.pre((t + u * NT) * RR);
#pragma unroll
                for (int c = 0; c < RR; ++c) TB_SM_STORE(sm[map.at(pre, c)], a[u * RR + c]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = NFULL - 1; k >= 0; --k) {
            const int ls = LOG2L - LOGE * (k + 1);
            const int s = 1 << ls;
            const int g = t >> ls, r = t & (s - 1);
            const int pre = map.pre((g << (ls + LOGE)) + r);
            const bool first = (k == NFULL - 1) && (REM == 0 || REM_SHFL);
            if (!first) {
#pragma unroll
                for (int c = 0; c < E; ++c) TB_SM_LOAD(a[brev(c, LOGE)], sm[map.at(pre, c << ls)]);
            }
            bool butterflies_done = false;
            if (ls > 0) {
                cx<R> w[E];
                stage_twiddles_of_thread<R, E>(tw + stage_tw_off(LOG2L, k, LOGE), s, r, w);
#if TB_FMA_BFLY
                dit_regs_tw<R, E>(a, w);
                butterflies_done = true;
#else
#pragma unroll
                for (int c = 1; c < E; ++c) a[brev(c, LOGE)] = cmulc(a[brev(c, LOGE)], w[c]);
#endif
            }
            if (!butterflies_done) dit_regs<R, E>(a);
            if (k > 0) {
#pragma unroll
                for (int c = 0; c < E; ++c) TB_SM_STORE(sm[map.at(pre, c << ls)], a[c]);
                exchange_barrier<LOGE, Map::TSHIFT>(ls);     // writer and reader of an element differ in bits [ls, ls+LOGE) of t
            }
        }
    }
};

// Runtime versions of the layout maps (for the un-scramble kernel and the host).
struct PlanDims {
    int log2l;   // FFT length of this level
    int LOGE;    // log2 of the points per thread
    __host__ __device__ int L() const { return 1 << log2l; }
    __host__ __device__ int NT() const { return (1 << log2l) >> LOGE; }
    __host__ __device__ int nfull() const { return log2l / LOGE; }
    __host__ __device__ int rem() const { return log2l % LOGE; }
    __host__ __device__ int pos_of_reg(int i, int t) const {
        const int E = 1 << LOGE;
        int rm = rem(), rr = 1 << rm;
        if (rm == 1) return ((t >> 1) << (1 + LOGE)) + (t & 1) + (brev(i, LOGE) << 1);      // cross-lane radix-2 (SlotFFT::REM_SHFL)
        return rm == 0 ? t * E + brev(i, LOGE) : (t + (i / rr) * NT()) * rr + brev(i % rr, rm);
    }
    __host__ __device__ int freq_of_pos(int P) const {
        const int E = 1 << LOGE;
        int k = 0, mult = 1;
        for (int s = 0; s < nfull(); ++s) {
            k += ((P >> (log2l - LOGE * (s + 1))) & (E - 1)) * mult;
            mult <<= LOGE;
        }
        if (rem() > 0) k += (P & ((1 << rem()) - 1)) * mult;
        return k;
    }
    // inverse maps
    __host__ __device__ int pos_of_freq(int k) const {
        const int E = 1 << LOGE;
        int P = 0;
        for (int s = 0; s < nfull(); ++s) {
            P |= (k & (E - 1)) << (log2l - LOGE * (s + 1));
            k >>= LOGE;
        }
        if (rem() > 0) P |= k & ((1 << rem()) - 1);
        return P;
    }
    // spectrum-layout offset (i*NT + t) of position P
    __host__ __device__ int layout_of_pos(int P) const {
        const int E = 1 << LOGE;
        int rm = rem(), rr = 1 << rm, nt = NT();
        if (rm == 0) {
            int t = P >> LOGE, c = P & (E - 1);
            return brev(c, LOGE) * nt + t;
        }
        if (rm == 1) {
            const int g = P >> (1 + LOGE), c = (P >> 1) & (E - 1), r = P & 1;
            return brev(c, LOGE) * nt + ((g << 1) | r);
        }
        int b = P >> rm, c = P & (rr - 1);
        int t = b % nt, u = b / nt;
        return (u * rr + brev(c, rm)) * nt + t;
    }
};

}  // namespace tb
