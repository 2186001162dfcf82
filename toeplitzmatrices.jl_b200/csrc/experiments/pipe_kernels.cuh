// Pass B as a persistent, double-buffered kernel: one CTA per SM loops over rows; while it transforms
// row i in shared-memory buffer D[i&1], cp.async (LDGSTS) streams row i+1 into D[(i+1)&1] and the
// spectrum row into S, so no warp ever waits on global-memory latency (the three-kernel path spends
// ~45 % of pass B in exactly that wait, profiles/experiments_r01.md).  192 KiB shared memory per CTA.
#pragma once
#include "../fft_kernels.cuh"

namespace tb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
template <typename C> __device__ __forceinline__ void cp_async_cx(C* dst, const C* src) {
    if constexpr (sizeof(C) == 16) cp_async16(dst, src); else cp_async8(dst, src);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows: scratch rows [0, nrows) of length L (in place); spectrum row = row % spec_rows
template <typename R, int LOG2L>
__global__ void __launch_bounds__((1 << (LOG2L - le_of<R>())), 1)
k_rowsB_pipe(cx<R>* __restrict__ scratch, long long nrows, const cx<R>* __restrict__ spec, int spec_rows,
             const cx<R>* __restrict__ tw) {
    using F = SlotFFT<R, LOG2L>;
    constexpr int L = F::L, NT = F::NT, E = F::E;
    extern __shared__ __align__(16) unsigned char smraw[];
    cx<R>* D0 = reinterpret_cast<cx<R>*>(smraw);
    cx<R>* D1 = D0 + L;
    cx<R>* S = D1 + L;
    const int t = threadIdx.x;
    RowMap<R> map;
    // natural index t + c*NT lives at swizzled address map.at(map.pre(t), c*NT)
    const int pre0 = map.pre(t);

    auto issue_row = [&](cx<R>* dst, long long row) {
        const cx<R>* src = scratch + row * (long long)L + t;
#pragma unroll
        for (int c = 0; c < E; ++c) cp_async_cx(dst + map.at(pre0, c * NT), src + c * NT);
    };
    long long row = blockIdx.x;
    if (row >= nrows) return;
    issue_row(D0, row);
    cp_async_commit();
    int cur = 0;
    for (; row < nrows; row += gridDim.x, cur ^= 1) {
        cx<R>* D = cur ? D1 : D0;
        const long long next = row + gridDim.x;
        if (next < nrows) issue_row(cur ? D0 : D1, next);
        cp_async_commit();
        {   // spectrum row in layout order: register i of thread t <- S[i*NT + t]
            const cx<R>* srow = spec + (row % spec_rows) * (long long)L + t;
#pragma unroll
            for (int i = 0; i < E; ++i) cp_async_cx(S + i * NT + t, srow + i * NT);
        }
        cp_async_commit();
        cp_async_wait<2>();              // this row's data has landed (the two newest groups may be in flight)
        __syncthreads();
        cx<R> a[E];
        F::template forward<RowMap<R>, true>(a, t, D, map, tw);
        cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int i = 0; i < E; ++i) a[i] = cmul(a[i], S[i * NT + t]);
        F::inverse(a, t, D, map, tw);
        cx<R>* out = scratch + row * (long long)L + t;
#pragma unroll
        for (int c = 0; c < E; ++c) out[c * NT] = a[c];
        __syncthreads();                 // D and S are free for the next prefetches
    }
}

template <typename R> cudaError_t launch_rowsB_pipe(int log2l, cx<R>* scratch, long long nrows, const cx<R>* spec,
                                                    int spec_rows, const cx<R>* tw, int num_sms, cudaStream_t s);
template <typename R> bool rowsB_pipe_supported(int log2l);

template <typename R, int LOG2L>
cudaError_t launch_rowsB_pipe_one(cx<R>* scratch, long long nrows, const cx<R>* spec, int spec_rows, const cx<R>* tw,
                                  int num_sms, cudaStream_t s) {
    constexpr int threads = 1 << (LOG2L - le_of<R>());
    constexpr size_t smem = 3 * sizeof(cx<R>) * (size_t)(1 << LOG2L);
    auto kern = k_rowsB_pipe<R, LOG2L>;
    static bool configured[64] = {};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    long long grid = nrows < num_sms ? nrows : num_sms;
    kern<<<(unsigned)grid, threads, smem, s>>>(scratch, nrows, spec, spec_rows, tw);
    return cudaGetLastError();
}

}  // namespace tb
