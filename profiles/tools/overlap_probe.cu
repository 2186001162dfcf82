// Does a phase-structured kernel (shared-memory phase | barrier | FP64 phase | barrier) overlap its LSU and
// FP64 work across the CTAs resident on an SM?  Times LSU-only, FP64-only and both.
#include <cstdio>
#include <cuda_runtime.h>

template <bool LSU, bool FMA>
__global__ void __launch_bounds__(256, 2) probe(double2* out, int iters, double a, double b) {
    extern __shared__ double2 sm[];
    const int t = threadIdx.x;
    double2 r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = make_double2(t + i, t - i);
    for (int i = t; i < 4096; i += 256) sm[i] = make_double2(i, -i);
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        if (LSU) {
#pragma unroll
            for (int c = 0; c < 16; ++c) sm[(t + c * 256) ^ ((it & 7))] = r[c];
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 16; ++c) r[c] = sm[((t * 16 + c) ^ (t >> 1 & 7)) & 4095];
        }
        if (FMA) {
#pragma unroll
            for (int k = 0; k < 7; ++k)
#pragma unroll
                for (int c = 0; c < 16; ++c) { r[c].x = fma(r[c].x, a, b); r[c].y = fma(r[c].y, a, r[c].x); }
        }
        __syncthreads();
    }
    double2 s = make_double2(0, 0);
#pragma unroll
    for (int c = 0; c < 16; ++c) { s.x += r[c].x; s.y += r[c].y; }
    out[blockIdx.x * 256 + t] = s;
}

template <bool LSU, bool FMA>
float run(int sms, int iters, double2* out) {
    auto k = probe<LSU, FMA>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<sms * 2, 256, 65536>>>(out, 10, 1.0000001, 1e-9);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<<<sms * 2, 256, 65536>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double2* out; cudaMalloc(&out, sizeof(double2) * sms * 2 * 256);
    const int iters = 20000;
    float a = run<true, false>(sms, iters, out), b = run<false, true>(sms, iters, out), c = run<true, true>(sms, iters, out);
    printf("per iteration per SM (2 CTAs x 256 thr): LSU-only %.0f cyc, FP64-only %.0f cyc, both %.0f cyc (sum %.0f, max %.0f)\n",
           a * 1e-3 / iters * 1.965e9, b * 1e-3 / iters * 1.965e9, c * 1e-3 / iters * 1.965e9,
           (a + b) * 1e-3 / iters * 1.965e9, (a > b ? a : b) * 1e-3 / iters * 1.965e9);
    return 0;
}
