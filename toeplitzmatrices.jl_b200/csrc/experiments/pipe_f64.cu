#include "pipe_kernels.cuh"
namespace tb {
template <> bool rowsB_pipe_supported<double>(int l) { return l == 12 || l == 11; }
template <> cudaError_t launch_rowsB_pipe<double>(int l, cx<double>* sc, long long nrows, const cx<double>* spec, int sr,
                                                  const cx<double>* tw, int sms, cudaStream_t s) {
    if (l == 12) return launch_rowsB_pipe_one<double, 12>(sc, nrows, spec, sr, tw, sms, s);
    if (l == 11) return launch_rowsB_pipe_one<double, 11>(sc, nrows, spec, sr, tw, sms, s);
    return cudaErrorInvalidValue;
}
template <> bool rowsB_pipe_supported<float>(int l) { return l == 13; }
template <> cudaError_t launch_rowsB_pipe<float>(int l, cx<float>* sc, long long nrows, const cx<float>* spec, int sr,
                                                 const cx<float>* tw, int sms, cudaStream_t s) {
    if (l == 13) return launch_rowsB_pipe_one<float, 13>(sc, nrows, spec, sr, tw, sms, s);
    return cudaErrorInvalidValue;
}
}  // namespace tb
