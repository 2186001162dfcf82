#include "fft_launch.cuh"
namespace tb {
#define TB_COLS_CASES(LAST)                                                        \
    switch (log2l1 * 100 + ta) {                                                   \
    case 608: return launch_cols_one<double, 6, 8, LAST>(a, s);                    \
    case 708: return launch_cols_one<double, 7, 8, LAST>(a, s);                    \
    case 808: return launch_cols_one<double, 8, 8, LAST>(a, s);                    \
    case 908: return launch_cols_one<double, 9, 8, LAST>(a, s);                    \
    case 1008: return launch_cols_one<double, 10, 8, LAST>(a, s);                  \
    case 1004: return launch_cols_one<double, 10, 4, LAST>(a, s);                  \
    case 904: return launch_cols_one<double, 9, 4, LAST>(a, s);                    \
    case 1108: return launch_cols_one<double, 11, 8, LAST>(a, s);                  \
    case 1104: return launch_cols_one<double, 11, 4, LAST>(a, s);                  \
    case 1202: return launch_cols_one<double, 12, 2, LAST>(a, s);                  \
    default: return cudaErrorInvalidValue;                                         \
    }
template <> cudaError_t launch_colsA<double>(int log2l1, int ta, const ConvArgs<double>& a, cudaStream_t s) { TB_COLS_CASES(false) }
template <> cudaError_t launch_colsC<double>(int log2l1, int ta, const ConvArgs<double>& a, cudaStream_t s) { TB_COLS_CASES(true) }
}  // namespace tb
