// TMA-staged variants of the strided passes of the two-level transform (north_star: "TMA staging of multi-pass
// tiles for n >= 2^20").
//
//  k_colsA_tma : the x tile of a CTA -- TA adjacent columns of the N1 x N2 view of every user column of the slot -- is
//                fetched by ONE thread as a 3-D tensor box (cp.async.bulk.tensor, SASS UTMALDG) into shared memory and
//                completes on an mbarrier; zero padding (rows past n/N2, the missing partner of an odd column count)
//                is the tensor map's out-of-bounds fill.  The threads then pick their 16 points from the staged tile
//                (Hankel: with reversed indices) and run the same forward column FFT as k_colsA.
//  k_colsC_tma : after the inverse column FFT the scaled outputs are staged in shared memory and written back by one
//                bulk tensor store (UTMASTG); rows past m/N2 and the missing odd column are clipped by the tensor map.
//
// Eligibility (checked on the host, else the LDG/STG kernels run): n (resp. m) a multiple of N2, 16-byte aligned
// base and column stride, no column indirection, beta == 0, real pairs (Float64) or complex elements.
#pragma once
#include <cuda.h>
#include "fft_kernels.cuh"

namespace tb {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TB_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TB_MBAR_DONE;\n"
        "bra TB_MBAR_WAIT;\n"
        "TB_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// Geometry of a staged tile: 8-byte units; U units per element, NC user columns per FFT slot, rows in chunks of <= 256
// (the box-dimension limit of a tensor map).  Layout in shared memory: [chunk][column][row in chunk][TA * U units].
template <typename R, int LOG2L1, int TA, bool HALF, bool PAIR>
struct TmaTile {
    static constexpr int U = PAIR ? 1 : (int)(sizeof(cx<R>) / 8);
    static constexpr int NC = PAIR ? 2 : 1;
    static constexpr int ROWS = HALF ? (1 << (LOG2L1 - 1)) : (1 << LOG2L1);
    static constexpr int CH = ROWS < 256 ? ROWS : 256;
    static constexpr int NCHUNK = ROWS / CH;
    static constexpr int W = TA * U;                              // units per staged row
    static constexpr unsigned BYTES = (unsigned)NC * ROWS * W * 8;
    __device__ __forceinline__ static int unit(int col, int row, int q) {
        return (((row / CH) * NC + col) * CH + (row % CH)) * W + q * U;
    }
};

// grid.x = (N2/TA) * nfft ; block = TA * NT1 (same as k_colsA)
template <typename R, int LOG2L1, int TA, bool HALF, bool PAIR>
__global__ void __launch_bounds__((1 << (LOG2L1 - le_of<R>())) * TA, min_ctas<R>((1 << (LOG2L1 - le_of<R>())) * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsA_tma(const ConvArgs<R> p, const __grid_constant__ CUtensorMap xmap, int rows_in, int col0) {
    using F = SlotFFT<R, LOG2L1>;
    using TT = TmaTile<R, LOG2L1, TA, HALF, PAIR>;
    constexpr int NT = F::NT, E = F::E;
    extern __shared__ __align__(128) unsigned char smraw_tma[];
    unsigned char* smraw = smraw_tma;
    __shared__ __align__(8) unsigned long long bar;
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int l2 = p.log2n2;
    const unsigned tiles = (1u << l2) / TA;
    const long long f = blockIdx.x / tiles;
    const int j20 = (int)(blockIdx.x % tiles) * TA;
    const int j2 = j20 + q;
    const int rev = p.reverse_x;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_expect_tx(&bar, TT::BYTES);
        const int c0 = (rev ? ((1 << l2) - j20 - TA) : j20) * TT::U;
#pragma unroll
        for (int ch = 0; ch < TT::NCHUNK; ++ch)
            tma_load_3d(smraw + (size_t)ch * TT::NC * TT::CH * TT::W * 8, &xmap, c0, ch * TT::CH, col0 + (int)f * TT::NC, &bar);
    }
    __syncthreads();                       // the barrier is initialised before anyone polls it
    mbar_wait(&bar, 0);
    ColMap<TA> map{q};
    cx<R> a[E];
    {
        const double* st = reinterpret_cast<const double*>(smraw);
        const int qq = rev ? (TA - 1 - q) : q;
#pragma unroll
        for (int c = 0; c < (HALF ? E / 2 : E); ++c) {
            const int row = t + c * NT;
            // Hankel: logical element idx maps to n-1-idx = (rows_in-1-row, N2-1-j2); rows past the input are padding
            const int srow = rev ? (rows_in - 1 - row) : row;
            if (rev && srow < 0) { a[c] = mk<R>(0, 0); continue; }
            if constexpr (PAIR) {
                a[c] = mk<R>((R)st[TT::unit(0, srow, qq)], (R)st[TT::unit(1, srow, qq)]);
            } else {
                a[c] = *reinterpret_cast<const cx<R>*>(st + TT::unit(0, srow, qq));
            }
        }
    }
    __syncthreads();                       // staging aliases the FFT region: everyone has read before stage 0 stores
    F::template forward<ColMap<TA>, false, HALF>(a, t, sm, map, p.tw);
    cx<R>* out = p.scratch + (f << (l2 + LOG2L1)) + j2;
    cx<R> w[E];
    level_twiddles<R, F>(p, t, j2, w);
#pragma unroll
    for (int i = 0; i < E; ++i) out[(long long)F::pos_of_reg(i, t) << l2] = cmul(a[i], w[i]);
}

template <typename R, int LOG2L1, int TA, bool HALF, bool PAIR>
__global__ void __launch_bounds__((1 << (LOG2L1 - le_of<R>())) * TA, min_ctas<R>((1 << (LOG2L1 - le_of<R>())) * TA, (int)sizeof(cx<R>) * (1 << LOG2L1) * TA))
k_colsC_tma(const ConvArgs<R> p, const __grid_constant__ CUtensorMap ymap, int col0) {
    using F = SlotFFT<R, LOG2L1>;
    using TT = TmaTile<R, LOG2L1, TA, HALF, PAIR>;
    constexpr int NT = F::NT, E = F::E;
    extern __shared__ __align__(128) unsigned char smraw_tma[];
    unsigned char* smraw = smraw_tma;
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int l2 = p.log2n2;
    const unsigned tiles = (1u << l2) / TA;
    const long long f = blockIdx.x / tiles;
    const int j20 = (int)(blockIdx.x % tiles) * TA;
    const int j2 = j20 + q;
    ColMap<TA> map{q};
    const cx<R>* in = p.scratch + (f << (l2 + LOG2L1)) + j2;
    cx<R> a[E];
    {
        cx<R> w[E];
        level_twiddles<R, F>(p, t, j2, w);
#pragma unroll
        for (int i = 0; i < E; ++i) a[i] = cmulc(ld_stream(in + ((long long)F::pos_of_reg(i, t) << l2)), w[i]);
    }
    F::inverse(a, t, sm, map, p.tw);
    __syncthreads();                       // the last stage's shared-memory reads are done: the region becomes the staging tile
    {
        double* st = reinterpret_cast<double*>(smraw);
        const cx<R> as = mk<R>(p.alpha.x * p.scale, p.alpha.y * p.scale);
#pragma unroll
        for (int j = 0; j < (HALF ? E / 2 : E); ++j) {
            const int row = t + j * NT;
            if constexpr (PAIR) {
                st[TT::unit(0, row, q)] = (double)(as.x * a[j].x);
                st[TT::unit(1, row, q)] = (double)(as.x * a[j].y);
            } else {
                *reinterpret_cast<cx<R>*>(st + TT::unit(0, row, q)) = cmul(as, a[j]);
            }
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int ch = 0; ch < TT::NCHUNK; ++ch)
            tma_store_3d(&ymap, j20 * TT::U, ch * TT::CH, col0 + (int)f * TT::NC, smraw + (size_t)ch * TT::NC * TT::CH * TT::W * 8);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory must outlive the reads of the store
    }
}

}  // namespace tb
