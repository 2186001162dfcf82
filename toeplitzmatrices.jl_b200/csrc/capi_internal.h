// Internal glue shared by the translation units of libtoeplitz_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/toeplitz_b200.h"

namespace tb {

int set_error(int status, const char* fmt, ...);
int cuda_status(cudaError_t e, const char* what);
void count_launch(int n = 1);

int levinson_dispatch(int dtype, bool durbin, int64_t n, int64_t nsys, const void* r, int64_t ldr, const void* b,
                      int64_t ldb, void* x, int64_t ldx, const void* diag, cudaStream_t s);

int levinson_multirhs_dispatch(int dtype, int64_t n, int64_t nrhs, const void* vc, const void* B, int64_t ldb, void* X,
                               int64_t ldx, void* work, cudaStream_t s);

}  // namespace tb
