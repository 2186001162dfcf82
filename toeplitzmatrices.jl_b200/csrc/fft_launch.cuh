// Host-side launchers for the templated FFT kernels (one translation unit per
// precision x kernel family so nvcc can build them in parallel).
#pragma once
#include "fft_kernels.cuh"

namespace tb {

// slots per CTA for the contiguous kernel: keep >= 128 threads per CTA
template <typename R> constexpr int rows_ts(int log2l) {
    return (log2l - le_of<R>()) >= 7 ? 1 : (1 << (7 - (log2l - le_of<R>())));
}

template <typename R> cudaError_t launch_rows(int log2l, const ConvArgs<R>& a, int num_sms, cudaStream_t s);
template <typename R> cudaError_t launch_colsA(int log2l1, int ta, const ConvArgs<R>& a, cudaStream_t s);
template <typename R> cudaError_t launch_colsC(int log2l1, int ta, const ConvArgs<R>& a, cudaStream_t s);

template <typename R, int LOG2L>
cudaError_t launch_rows_one(const ConvArgs<R>& a, int num_sms, cudaStream_t s) {
    constexpr int TS = rows_ts<R>(LOG2L);
    constexpr int threads = (1 << (LOG2L - le_of<R>())) * TS;
    constexpr size_t smem = sizeof(cx<R>) * (size_t)(1 << LOG2L) * TS;
    auto kern = k_rows<R, LOG2L, TS>;
    static bool configured[64] = {};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    long long ctas = (a.nfft + TS - 1) / TS;
    long long cap = (long long)num_sms * 16;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    kern<<<(unsigned)ctas, threads, smem, s>>>(a);
    return cudaGetLastError();
}

template <typename R, int LOG2L1, int TA, bool LAST>
cudaError_t launch_cols_one(const ConvArgs<R>& a, cudaStream_t s) {
    constexpr int threads = (1 << (LOG2L1 - le_of<R>())) * TA;
    constexpr size_t smem = sizeof(cx<R>) * (size_t)(1 << LOG2L1) * TA;
    auto kern = LAST ? k_colsC<R, LOG2L1, TA> : k_colsA<R, LOG2L1, TA>;
    static bool configured[64] = {};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    const long long tiles = (1LL << a.log2n2) / TA;
    const long long grid = tiles * a.nfft;
    kern<<<(unsigned)grid, threads, smem, s>>>(a);
    return cudaGetLastError();
}

}  // namespace tb
