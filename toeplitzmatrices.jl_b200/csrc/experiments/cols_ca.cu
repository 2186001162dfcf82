// Launchers of the fused "pass C of one group + pass A of the next" kernel (k_colsCA, fft_kernels.cuh) for the same
// (column length, tile width) plans as cols_f64.cu / cols_f32.cu.
#include "../fft_launch.cuh"
#include "ca_kernels.cuh"
namespace tb {

template <typename R, int LOG2L1, int TA, bool HALF, bool STAGE>
static cudaError_t launch_ca_v(const ConvArgs<R>& pc, const ConvArgs<R>& pa, cudaStream_t s) {
    constexpr int threads = (1 << (LOG2L1 - le_of<R>())) * TA;
    constexpr size_t smem = (size_t)ca_smem_bytes<R, LOG2L1, TA, STAGE>();
    auto kern = k_colsCA<R, LOG2L1, TA, HALF, STAGE>;
    static bool configured[64] = {};          // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = true;
    }
    const long long tiles = (1LL << pc.log2n2) / TA;
    const long long grid = tiles * (pc.nfft > pa.nfft ? pc.nfft : pa.nfft);
    kern<<<(unsigned)grid, threads, smem, s>>>(pc, pa);
    return cudaGetLastError();
}

// half: both sides half-padded (decided by the caller with the rule of launch_cols_one); stage: asynchronous x tile
template <typename R, int LOG2L1, int TA>
static cudaError_t launch_ca_one(const ConvArgs<R>& pc, const ConvArgs<R>& pa, bool half, bool stage, cudaStream_t s) {
    if constexpr (sizeof(R) == 8 && (size_t)ca_smem_bytes<R, LOG2L1, TA, true>() <= 200 * 1024) {
        if (half && stage) return launch_ca_v<R, LOG2L1, TA, true, true>(pc, pa, s);
    }
    if (half) return launch_ca_v<R, LOG2L1, TA, true, false>(pc, pa, s);
    return launch_ca_v<R, LOG2L1, TA, false, false>(pc, pa, s);
}

template <> cudaError_t launch_colsCA<double>(int log2l1, int ta, const ConvArgs<double>& pc, const ConvArgs<double>& pa, bool half, bool stage, cudaStream_t s) {
    switch (log2l1 * 100 + ta) {
    case 608: return launch_ca_one<double, 6, 8>(pc, pa, half, stage, s);
    case 708: return launch_ca_one<double, 7, 8>(pc, pa, half, stage, s);
    case 808: return launch_ca_one<double, 8, 8>(pc, pa, half, stage, s);
    case 908: return launch_ca_one<double, 9, 8>(pc, pa, half, stage, s);
    case 1008: return launch_ca_one<double, 10, 8>(pc, pa, half, stage, s);
    case 1108: return launch_ca_one<double, 11, 8>(pc, pa, half, stage, s);
    case 1104: return launch_ca_one<double, 11, 4>(pc, pa, half, stage, s);
    case 1202: return launch_ca_one<double, 12, 2>(pc, pa, half, stage, s);
    default: return cudaErrorInvalidValue;
    }
}
template <> cudaError_t launch_colsCA<float>(int log2l1, int ta, const ConvArgs<float>& pc, const ConvArgs<float>& pa, bool half, bool stage, cudaStream_t s) {
    switch (log2l1 * 100 + ta) {
    case 616: return launch_ca_one<float, 6, 16>(pc, pa, half, stage, s);
    case 716: return launch_ca_one<float, 7, 16>(pc, pa, half, stage, s);
    case 816: return launch_ca_one<float, 8, 16>(pc, pa, half, stage, s);
    case 916: return launch_ca_one<float, 9, 16>(pc, pa, half, stage, s);
    case 1016: return launch_ca_one<float, 10, 16>(pc, pa, half, stage, s);
    case 1108: return launch_ca_one<float, 11, 8>(pc, pa, half, stage, s);
    case 1204: return launch_ca_one<float, 12, 4>(pc, pa, half, stage, s);
    default: return cudaErrorInvalidValue;
    }
}
template <typename R> bool cols_ca_supported(int log2l1, int ta) {
    const int key = log2l1 * 100 + ta;
    if (sizeof(R) == 8) return key == 608 || key == 708 || key == 808 || key == 908 || key == 1008 || key == 1108 || key == 1104 || key == 1202;
    return key == 616 || key == 716 || key == 816 || key == 916 || key == 1016 || key == 1108 || key == 1204;
}
template bool cols_ca_supported<double>(int, int);
template bool cols_ca_supported<float>(int, int);
}  // namespace tb
