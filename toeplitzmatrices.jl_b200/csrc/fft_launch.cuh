// Host-side launchers for the templated FFT kernels (one translation unit per
// precision x kernel family so nvcc can build them in parallel).
#pragma once
#include <string.h>
#include "fft_kernels.cuh"

namespace tb {

// slots per CTA for the contiguous kernel: keep >= 128 threads per CTA
template <typename R> constexpr int rows_ts(int log2l) {
    return (log2l - le_of<R>()) >= 7 ? 1 : (1 << (7 - (log2l - le_of<R>())));
}

template <typename R> cudaError_t launch_rows(int log2l, const ConvArgs<R>& a, int num_sms, cudaStream_t s);
template <typename R> cudaError_t launch_colsA(int log2l1, int ta, const ConvArgs<R>& a, cudaStream_t s);
template <typename R> cudaError_t launch_colsC(int log2l1, int ta, const ConvArgs<R>& a, cudaStream_t s);
// fused pass C (pc) + pass A of the next group on the same workspace (pa); experiments/cols_ca.cu, TB_EXP_CA builds only
template <typename R> cudaError_t launch_colsCA(int log2l1, int ta, const ConvArgs<R>& pc, const ConvArgs<R>& pa, bool half, bool stage, cudaStream_t s);
template <typename R> bool cols_ca_supported(int log2l1, int ta);

// One pass: a plain launch, or -- a.pdl -- a programmatic dependent launch behind the previous pass of the same product
template <typename R, typename K>
cudaError_t launch_pass(K kern, unsigned grid, int threads, size_t smem, cudaStream_t s, const ConvArgs<R>& a) {
    if (!a.pdl) {
        kern<<<grid, threads, smem, s>>>(a);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

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
    return launch_pass(kern, (unsigned)ctas, threads, smem, s, a);
}

// pruned variants (half of the embedding is padding) exist for radix-16 threads with at least one full stage
extern int g_prune_half;
template <typename R, int LOG2L1, int TA, bool LAST, bool HALF>
cudaError_t launch_cols_v(const ConvArgs<R>& a, cudaStream_t s);

template <typename R, int LOG2L1, int TA, bool LAST>
cudaError_t launch_cols_one(const ConvArgs<R>& a, cudaStream_t s) {
    const long long half = 1LL << (a.log2n2 + LOG2L1 - 1);
    const bool can_half = g_prune_half && le_of<R>() == 4 && LOG2L1 >= 4 &&
                          (LAST ? (a.m_out <= half) : (a.n_in <= half && a.x_mode != IO_SPEC));
    if (can_half) return launch_cols_v<R, LOG2L1, TA, LAST, true>(a, s);
    return launch_cols_v<R, LOG2L1, TA, LAST, false>(a, s);
}

// the outer passes of the three-level plan (no [E][N2] twiddle table) exist for its one column length, 2^9
template <typename R, int LOG2L1, int TA, bool LAST, bool HALF>
cudaError_t launch_cols_v(const ConvArgs<R>& a, cudaStream_t s) {
    constexpr int threads = (1 << (LOG2L1 - le_of<R>())) * TA;
    constexpr size_t smem = sizeof(cx<R>) * (size_t)(1 << LOG2L1) * TA;
    if constexpr (LOG2L1 == 9 && TA == 128 / (int)sizeof(cx<R>)) {
        if (!a.tw_g) {
            auto kern3 = LAST ? k_colsC<R, LOG2L1, TA, HALF, false> : k_colsA<R, LOG2L1, TA, HALF, false>;
            static bool configured3[64] = {};
            int dev3 = 0;
            cudaGetDevice(&dev3);
            if (!configured3[dev3 & 63]) {
                cudaError_t e = cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                configured3[dev3 & 63] = true;
            }
            const long long tiles3 = (1LL << a.log2n2) / TA;
            return launch_pass(kern3, (unsigned)(tiles3 * a.nfft), threads, smem, s, a);
        }
    } else {
        if (!a.tw_g) return cudaErrorInvalidValue;
    }
    auto kern = LAST ? k_colsC<R, LOG2L1, TA, HALF> : k_colsA<R, LOG2L1, TA, HALF>;
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
    return launch_pass(kern, (unsigned)grid, threads, smem, s, a);
}

// ---- TMA-staged strided passes (tma_cols.cuh), instantiated for the plans of the BASELINE configs ----------------
struct TmaLaunch { int rows; int col0; bool half; bool pair; };
template <typename R> bool cols_tma_supported(int log2l1, int ta, bool pair);
template <typename R> cudaError_t launch_colsA_tma(int log2l1, int ta, const ConvArgs<R>& a, const void* tensor_map, const TmaLaunch& t, cudaStream_t s);
template <typename R> cudaError_t launch_colsC_tma(int log2l1, int ta, const ConvArgs<R>& a, const void* tensor_map, const TmaLaunch& t, cudaStream_t s);

}  // namespace tb
