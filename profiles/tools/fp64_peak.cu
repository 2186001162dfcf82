// Microbenchmark: sustained FP64 / FP32 FMA throughput and shared-memory bandwidth of the GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu && ./fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void smem_kernel(double2* out, int iters) {
    extern __shared__ double2 sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double2 v = sm[(t + c * 512 + it) & 4095];
            acc.x += v.x; acc.y += v.y;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) sm[(t + c * 512 + it) & 4095] = acc;
    }
    out[blockIdx.x * blockDim.x + t] = acc;
}

template <typename T, int ILP>
double run(int sms, int ctas_per_sm, int threads, int iters) {
    T* out;
    cudaMalloc(&out, sizeof(T) * sms * ctas_per_sm * threads);
    fma_kernel<T, ILP><<<sms * ctas_per_sm, threads>>>(out, 10, (T)1.0001, (T)0.5);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_kernel<T, ILP><<<sms * ctas_per_sm, threads>>>(out, iters, (T)1.0001, (T)0.5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(out);
    return 2.0 * ILP * (double)iters * sms * ctas_per_sm * threads / (ms * 1e-3) / 1e12;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("SMs %d, clock %d kHz\n", sms, clk);
    for (int w : {128, 256, 512, 1024})
        printf("FP64 FMA ILP8  %4d thr x2 CTA/SM: %.2f TFLOP/s\n", w, run<double, 8>(sms, 2, w, 20000));
    // occupancy x ILP sweep: how many independent FP64 FMAs per SM sub-partition the pipe needs
    for (int w : {128, 256, 512, 1024}) {
        printf("FP64 FMA %4d thr/SM: ILP1 %.2f  ILP2 %.2f  ILP4 %.2f  ILP8 %.2f TFLOP/s\n", w, run<double, 1>(sms, 1, w, 40000),
               run<double, 2>(sms, 1, w, 40000), run<double, 4>(sms, 1, w, 20000), run<double, 8>(sms, 1, w, 20000));
    }
    printf("FP64 FMA ILP16 512 thr x2: %.2f TFLOP/s\n", run<double, 16>(sms, 2, 512, 10000));
    printf("FP32 FMA ILP8  1024 thr x2: %.2f TFLOP/s\n", run<float, 8>(sms, 2, 1024, 40000));
    // shared memory: 16 B per lane loads+stores
    double2* out;
    cudaMalloc(&out, sizeof(double2) * sms * 2 * 512);
    cudaFuncSetAttribute(smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    smem_kernel<<<sms * 2, 512, 65536>>>(out, 10);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    cudaEventRecord(e0);
    smem_kernel<<<sms * 2, 512, 65536>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double bytes = 16.0 * 16 * iters * (double)sms * 2 * 512;
    printf("shared memory LDS.128+STS.128: %.1f TB/s aggregate = %.1f B/clk/SM at %.0f MHz\n", bytes / (ms * 1e-3) / 1e12,
           bytes / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
    return 0;
}
