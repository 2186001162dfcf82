// Instantiations and launchers of the TMA-staged strided passes (tma_cols.cuh).
#include "fft_launch.cuh"
#include "tma_cols.cuh"

namespace tb {

template <typename R, int LOG2L1, int TA, bool LAST, bool HALF, bool PAIR>
static cudaError_t launch_tma_one(const ConvArgs<R>& a, const CUtensorMap& map, const TmaLaunch& t, cudaStream_t s) {
    constexpr int threads = (1 << (LOG2L1 - le_of<R>())) * TA;
    constexpr size_t smem = sizeof(cx<R>) * (size_t)(1 << LOG2L1) * TA;
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const long long tiles = (1LL << a.log2n2) / TA;
    const unsigned grid = (unsigned)(tiles * a.nfft);
    if constexpr (LAST) {
        auto kern = k_colsC_tma<R, LOG2L1, TA, HALF, PAIR>;
        if (!configured[dev & 63]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev & 63] = true;
        }
        kern<<<grid, threads, smem, s>>>(a, map, t.col0);
    } else {
        auto kern = k_colsA_tma<R, LOG2L1, TA, HALF, PAIR>;
        if (!configured[dev & 63]) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            configured[dev & 63] = true;
        }
        kern<<<grid, threads, smem, s>>>(a, map, t.rows, t.col0);
    }
    return cudaGetLastError();
}

template <typename R, int LOG2L1, int TA, bool LAST>
static cudaError_t launch_tma_hp(const ConvArgs<R>& a, const CUtensorMap& map, const TmaLaunch& t, cudaStream_t s) {
    if constexpr (sizeof(R) == 8) {
        if (t.pair) return t.half ? launch_tma_one<R, LOG2L1, TA, LAST, true, true>(a, map, t, s)
                                  : launch_tma_one<R, LOG2L1, TA, LAST, false, true>(a, map, t, s);
    }
    return t.half ? launch_tma_one<R, LOG2L1, TA, LAST, true, false>(a, map, t, s)
                  : launch_tma_one<R, LOG2L1, TA, LAST, false, false>(a, map, t, s);
}

template <> bool cols_tma_supported<double>(int l1, int ta, bool) { return ta == 8 && (l1 == 9 || l1 == 10); }
template <> bool cols_tma_supported<float>(int l1, int ta, bool pair) { return !pair && ta == 8 && l1 == 11; }

template <> cudaError_t launch_colsA_tma<double>(int l1, int ta, const ConvArgs<double>& a, const void* m, const TmaLaunch& t, cudaStream_t s) {
    const CUtensorMap& map = *reinterpret_cast<const CUtensorMap*>(m);
    if (ta == 8 && l1 == 9) return launch_tma_hp<double, 9, 8, false>(a, map, t, s);
    if (ta == 8 && l1 == 10) return launch_tma_hp<double, 10, 8, false>(a, map, t, s);
    return cudaErrorInvalidValue;
}
template <> cudaError_t launch_colsC_tma<double>(int l1, int ta, const ConvArgs<double>& a, const void* m, const TmaLaunch& t, cudaStream_t s) {
    const CUtensorMap& map = *reinterpret_cast<const CUtensorMap*>(m);
    if (ta == 8 && l1 == 9) return launch_tma_hp<double, 9, 8, true>(a, map, t, s);
    if (ta == 8 && l1 == 10) return launch_tma_hp<double, 10, 8, true>(a, map, t, s);
    return cudaErrorInvalidValue;
}
template <> cudaError_t launch_colsA_tma<float>(int l1, int ta, const ConvArgs<float>& a, const void* m, const TmaLaunch& t, cudaStream_t s) {
    const CUtensorMap& map = *reinterpret_cast<const CUtensorMap*>(m);
    if (ta == 8 && l1 == 11) return launch_tma_hp<float, 11, 8, false>(a, map, t, s);
    return cudaErrorInvalidValue;
}
template <> cudaError_t launch_colsC_tma<float>(int l1, int ta, const ConvArgs<float>& a, const void* m, const TmaLaunch& t, cudaStream_t s) {
    const CUtensorMap& map = *reinterpret_cast<const CUtensorMap*>(m);
    if (ta == 8 && l1 == 11) return launch_tma_hp<float, 11, 8, true>(a, map, t, s);
    return cudaErrorInvalidValue;
}

}  // namespace tb
