#include "fft_launch.cuh"
namespace tb {
#define TB_COLS_CASES(LAST)                                                        \
    switch (log2l1 * 100 + ta) {                                                   \
    case 616: return launch_cols_one<float, 6, 16, LAST>(a, s);                    \
    case 716: return launch_cols_one<float, 7, 16, LAST>(a, s);                    \
    case 816: return launch_cols_one<float, 8, 16, LAST>(a, s);                    \
    case 916: return launch_cols_one<float, 9, 16, LAST>(a, s);                    \
    case 1016: return launch_cols_one<float, 10, 16, LAST>(a, s);                  \
    case 1108: return launch_cols_one<float, 11, 8, LAST>(a, s);                   \
    case 1104: return launch_cols_one<float, 11, 4, LAST>(a, s);                   \
    case 1204: return launch_cols_one<float, 12, 4, LAST>(a, s);                   \
    default: return cudaErrorInvalidValue;                                         \
    }
template <> cudaError_t launch_colsA<float>(int log2l1, int ta, const ConvArgs<float>& a, cudaStream_t s) { TB_COLS_CASES(false) }
template <> cudaError_t launch_colsC<float>(int log2l1, int ta, const ConvArgs<float>& a, cudaStream_t s) { TB_COLS_CASES(true) }
}  // namespace tb
