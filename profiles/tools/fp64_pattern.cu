// Microbenchmark: FP64 FMA rate of the generator-form Levinson update pattern
//   A[i] = fma(alpha, B[i], A[i]);  B[i] = fma(alpha, A_old[i], B[i]);  X[i] = fma(mu, B_old[i], X[i])     (i = 0..15)
// (three distinct 64-bit register operands per FMA pair, one shared scalar) against the plain acc = acc*a + b chain of
// fp64_peak.cu, at 3 and 4 warps per SM sub-partition.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pattern fp64_pattern.cu && ./fp64_pattern
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int MODE>
__global__ void __launch_bounds__(128) pat_kernel(double* out, int iters, double a0, double m0) {
    double A[N], B[N], X[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { A[i] = threadIdx.x + i; B[i] = 0.5 * i; X[i] = 1.0 / (i + 1); }
    double alpha = a0, mu = m0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {            // Levinson pattern
#pragma unroll
            for (int i = 0; i < N; ++i) X[i] = fma(mu, B[i], X[i]);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ao = A[i], bo = B[i];
                A[i] = fma(alpha, bo, ao);
                B[i] = fma(alpha, ao, bo);
            }
        } else if (MODE == 1) {     // same FMA count, independent chains with shared scalars only
#pragma unroll
            for (int i = 0; i < N; ++i) { X[i] = fma(X[i], mu, alpha); A[i] = fma(A[i], alpha, mu); B[i] = fma(B[i], alpha, mu); }
        } else {                    // Durbin pattern (2 FMAs per slot)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double ao = A[i], bo = B[i];
                A[i] = fma(alpha, bo, ao);
                B[i] = fma(alpha, ao, bo);
            }
        }
        alpha = -alpha * 0.999; mu = mu * 1.0001;      // scalars change every step, as in the recursion
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += A[i] + B[i] + X[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int N, int MODE>
double run(int sms, int ctas_per_sm, int iters) {
    double* out;
    cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * 128);
    pat_kernel<N, MODE><<<sms * ctas_per_sm, 128>>>(out, 10, 0.3, 0.7);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    pat_kernel<N, MODE><<<sms * ctas_per_sm, 128>>>(out, iters, 0.3, 0.7);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    const double fmas = (MODE == 2 ? 2.0 : 3.0) * N;
    return 2.0 * fmas * (double)iters * sms * ctas_per_sm * 128 / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int c : {1, 2, 3, 4, 6, 8}) {
        printf("%d warps/SMSP: levinson pattern %.2f  independent chains %.2f  durbin pattern %.2f TFLOP/s\n", c,
               run<16, 0>(sms, c, 20000), run<16, 1>(sms, c, 20000), run<16, 2>(sms, c, 20000));
    }
    return 0;
}
