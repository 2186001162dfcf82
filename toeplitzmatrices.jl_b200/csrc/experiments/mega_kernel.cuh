// Persistent "megakernel" for the batched two-level convolution (the headline path,
// /root/reference/src/linearalgebra.jl:88-99 column loop): ONE cooperative launch per batch call
// instead of three launches per column pair.  CTAs (2 per SM, all co-resident) pull tiles from a
// global ticket counter in a fixed software-pipelined order
//     A(0) B(0) | A(1) C(0) B(1) | A(2) C(1) B(2) | ... | C(P-1)
// (A = strided forward columns, B = contiguous rows fwd * spectrum * inv, C = strided inverse
// columns + epilogue).  A pass only ever waits on a pass that was handed out one full pass (>= one
// wave of tiles) earlier, through per-pair completion counters (release: __threadfence + atomicAdd,
// acquire: ld.acquire.gpu spin by one thread + __syncthreads).  Two workspaces (pair parity) keep
// the intermediate data L2-resident; the stage twiddle tables stay hot in L1 because CTAs persist.
// Deadlock freedom: tickets are claimed in order, a tile waits only on tiles with smaller tickets,
// and every claimed tile is held by a resident CTA (cooperative launch guarantees co-residency).
#pragma once
#include "../fft_kernels.cuh"

namespace tb {

template <typename R>
struct MegaArgs {
    ConvArgs<R> c;                 // x / y views, epilogue, spectrum, level twiddles, log2n2
    const cx<R>* tw_cols;          // stage table of the strided (N1) transforms
    const cx<R>* tw_rows;          // stage table of the contiguous (N2) transforms
    cx<R>* scratch[2];
    unsigned* ticket;              // [0] ticket, [1] abort flag
    unsigned* done;                // 3 * npairs counters: A, B, C per pair
    long long npairs;              // FFT slots (column pairs) in this call
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_cg(const double2* p) {
    double2 r;
    asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float2 ld_cg(const float2* p) {
    float2 r;
    asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p) : "memory");
    return r;
}

// ---- tile bodies: separate (non-inlined) functions so each gets its own register allocation ----
template <typename R, int LOG2L1, int LOG2L2, int TA>
__device__ __noinline__ void mega_tile_A(const MegaArgs<R>& m, long long i, unsigned rem, cx<R>* sm) {
    using FA = SlotFFT<R, LOG2L1>;
    constexpr int E = FA::E, l2 = LOG2L2;
    const ConvArgs<R>& p = m.c;
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int j2 = (int)rem * TA + q;
    ColMap<TA> map{q};
    cx<R> a[E];
    {
        const InView<R> in(p, i);
#pragma unroll
        for (int c = 0; c < E; ++c) a[c] = in.load(((t + c * FA::NT) << l2) + j2);
    }
    FA::forward(a, t, sm, map, m.tw_cols);
    cx<R>* out = m.scratch[i & 1] + j2;
    cx<R> w[E];
    level_twiddles<R, FA>(p, t, j2, w);
#pragma unroll
    for (int r = 0; r < E; ++r) out[(long long)FA::pos_of_reg(r, t) << l2] = cmul(a[r], w[r]);
}

template <typename R, int LOG2L1, int LOG2L2, int TA>
__device__ __noinline__ void mega_tile_B(const MegaArgs<R>& m, long long i, unsigned rem, cx<R>* sm) {
    using FB = SlotFFT<R, LOG2L2>;
    constexpr int E = FB::E, l2 = LOG2L2;
    const int t = threadIdx.x;
    RowMap<R> map;
    cx<R>* row = m.scratch[i & 1] + ((long long)rem << l2) + t;
    const cx<R>* srow = m.c.spec + ((long long)rem << l2) + t;
    cx<R> a[E];
#pragma unroll
    for (int c = 0; c < E; ++c) a[c] = ld_cg(row + c * FB::NT);
    FB::forward(a, t, sm, map, m.tw_rows);
#pragma unroll
    for (int r = 0; r < E; ++r) a[r] = cmul(a[r], ld_stream(srow + r * FB::NT));
    FB::inverse(a, t, sm, map, m.tw_rows);
#pragma unroll
    for (int c = 0; c < E; ++c) row[c * FB::NT] = a[c];
}

template <typename R, int LOG2L1, int LOG2L2, int TA>
__device__ __noinline__ void mega_tile_C(const MegaArgs<R>& m, long long i, unsigned rem, cx<R>* sm) {
    using FA = SlotFFT<R, LOG2L1>;
    constexpr int E = FA::E, l2 = LOG2L2;
    const ConvArgs<R>& p = m.c;
    const int q = threadIdx.x % TA, t = threadIdx.x / TA;
    const int j2 = (int)rem * TA + q;
    ColMap<TA> map{q};
    const cx<R>* in = m.scratch[i & 1] + j2;
    cx<R> a[E];
    {
        cx<R> w[E];
        level_twiddles<R, FA>(p, t, j2, w);
#pragma unroll
        for (int r = 0; r < E; ++r) a[r] = cmulc(ld_cg(in + ((long long)FA::pos_of_reg(r, t) << l2)), w[r]);
    }
    FA::inverse(a, t, sm, map, m.tw_cols);
    const OutView<R> out(p, i);
#pragma unroll
    for (int j = 0; j < E; ++j) out.store(((t + j * FA::NT) << l2) + j2, a[j]);
}

template <typename R, int LOG2L1, int LOG2L2, int TA>
__global__ void __launch_bounds__((1 << (LOG2L2 - le_of<R>())), 2)
k_mega(const __grid_constant__ MegaArgs<R> m) {
    using FA = SlotFFT<R, LOG2L1>;
    using FB = SlotFFT<R, LOG2L2>;
    static_assert(FA::NT * TA == FB::NT, "column and row tiles must use the same CTA size");
    constexpr unsigned tA = (1u << LOG2L2) / TA, tB = 1u << LOG2L1, tC = tA, per = tA + tB + tC;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* sm = reinterpret_cast<cx<R>*>(smraw);
    __shared__ unsigned s_tile;
    __shared__ int s_ok;
    const long long P = m.npairs;
    const unsigned long long total = (unsigned long long)P * per;
    unsigned* abort_flag = m.ticket + 1;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(m.ticket, 1u);
        __syncthreads();
        const unsigned T = s_tile;
        if (T >= total) break;
        // ---- decode the ticket ----------------------------------------------------------------
        long long i; unsigned rem; int type;             // type 0 = A, 1 = B, 2 = C
        if (T < tA + tB) { i = 0; rem = T; if (rem < tA) type = 0; else { type = 1; rem -= tA; } }
        else {
            const unsigned T2 = T - (tA + tB);
            i = 1 + T2 / per; rem = T2 % per;
            if (i < P) {
                if (rem < tA) type = 0;
                else if (rem < tA + tC) { type = 2; rem -= tA; i -= 1; }
                else { type = 1; rem -= tA + tC; }
            } else { type = 2; i = P - 1; }
        }
        // ---- wait for the producer pass ---------------------------------------------------------
        if (threadIdx.x == 0) {
            const unsigned* ctr = nullptr; unsigned target = 0;
            if (type == 0) { if (i >= 2) { ctr = m.done + 3 * (i - 2) + 2; target = tC; } }
            else if (type == 1) { ctr = m.done + 3 * i + 0; target = tA; }
            else { ctr = m.done + 3 * i + 1; target = tB; }
            int ok = 1;
            if (ctr) {
                unsigned spins = 0;
                while (ld_acquire_u32(ctr) < target) {
                    __nanosleep(40);
                    if (++spins > (1u << 25) || ld_acquire_u32(abort_flag)) { atomicExch(abort_flag, 1u); ok = 0; break; }
                }
            }
            s_ok = ok;
        }
        __syncthreads();
        if (!s_ok) break;
        if (type == 0) mega_tile_A<R, LOG2L1, LOG2L2, TA>(m, i, rem, sm);
        else if (type == 1) mega_tile_B<R, LOG2L1, LOG2L2, TA>(m, i, rem, sm);
        else mega_tile_C<R, LOG2L1, LOG2L2, TA>(m, i, rem, sm);
        // ---- publish completion -------------------------------------------------------------------
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(m.done + 3 * i + type, 1u);
        }
    }
}

template <typename R> cudaError_t launch_mega(int log2l1, int log2l2, int ta, const MegaArgs<R>& a, int num_sms, cudaStream_t s);
template <typename R> bool mega_supported(int log2l1, int log2l2, int ta);

template <typename R, int LOG2L1, int LOG2L2, int TA>
cudaError_t launch_mega_one(const MegaArgs<R>& a, int num_sms, cudaStream_t s) {
    constexpr int threads = 1 << (LOG2L2 - le_of<R>());
    constexpr size_t smem = sizeof(cx<R>) * (size_t)(1 << LOG2L2);
    static_assert(sizeof(cx<R>) * (size_t)(1 << LOG2L1) * TA <= smem, "column tile must fit the row buffer");
    auto kern = k_mega<R, LOG2L1, LOG2L2, TA>;
    static int occ_of[64] = {};               // per device
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_of[dev & 63];
    if (!occ) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        if (occ > 2) occ = 2;
    }
    void* params[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)(num_sms * occ)), dim3(threads), params, smem, s);
}

}  // namespace tb
