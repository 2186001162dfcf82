#include "fft_launch.cuh"
namespace tb {
template <> cudaError_t launch_rows<double>(int log2l, const ConvArgs<double>& a, int num_sms, cudaStream_t s) {
    switch (log2l) {
    case 4: return launch_rows_one<double, 4>(a, num_sms, s);
    case 5: return launch_rows_one<double, 5>(a, num_sms, s);
    case 6: return launch_rows_one<double, 6>(a, num_sms, s);
    case 7: return launch_rows_one<double, 7>(a, num_sms, s);
    case 8: return launch_rows_one<double, 8>(a, num_sms, s);
    case 9: return launch_rows_one<double, 9>(a, num_sms, s);
    case 10: return launch_rows_one<double, 10>(a, num_sms, s);
    case 11: return launch_rows_one<double, 11>(a, num_sms, s);
    case 12: return launch_rows_one<double, 12>(a, num_sms, s);
    case 13: return launch_rows_one<double, 13>(a, num_sms, s);
    default: return cudaErrorInvalidValue;
    }
}
}  // namespace tb
