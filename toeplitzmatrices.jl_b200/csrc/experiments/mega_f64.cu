#include "mega_kernel.cuh"
namespace tb {
template <> bool mega_supported<double>(int l1, int l2, int ta) { return l1 == 9 && l2 == 12 && ta == 8; }
template <> cudaError_t launch_mega<double>(int l1, int l2, int ta, const MegaArgs<double>& a, int num_sms, cudaStream_t s) {
    if (l1 == 9 && l2 == 12 && ta == 8) return launch_mega_one<double, 9, 12, 8>(a, num_sms, s);
    return cudaErrorInvalidValue;
}
template <> bool mega_supported<float>(int, int, int) { return false; }
template <> cudaError_t launch_mega<float>(int, int, int, const MegaArgs<float>&, int, cudaStream_t) { return cudaErrorInvalidValue; }
}  // namespace tb
