// Microbenchmark: the FP64 pipe's sustained rate on the instruction mix of the FFT passes themselves -- the library's own
// radix-16 register butterflies (fft_core.cuh: dif_regs / dit_regs), the stage-twiddle product tree and the twiddle
// multiplies -- with NO memory traffic: 16 complex points per thread stay in registers.  Gives the FP64-issue floor of
// the headline path in measured (not nominal) pipe cycles.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../toeplitzmatrices.jl_b200/csrc -o fp64_fft_mix fp64_fft_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_core.cuh"
using namespace tb;

// MODE 0: forward stage (butterfly + twiddle generation + twiddle multiply) followed by the mirrored inverse stage
// MODE 1: butterflies only (dif + dit)
template <int MODE>
__global__ void __launch_bounds__(256, 2) mix_kernel(double2* out, int iters, double2 w1) {
    cx<double> a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = mk<double>(threadIdx.x + i, 0.25 * i - threadIdx.x);
    cx<double> ww = w1;
    for (int it = 0; it < iters; ++it) {
        dif_regs<double, 16>(a);
        if (MODE == 0) {
            cx<double> w[16];
            w[1] = ww;
#pragma unroll
            for (int c = 2; c < 16; ++c) w[c] = cmul(w[c / 2], w[c - c / 2]);
#pragma unroll
            for (int c = 1; c < 16; ++c) a[brev(c, 4)] = cmul(a[brev(c, 4)], w[c]);
#pragma unroll
            for (int c = 1; c < 16; ++c) a[brev(c, 4)] = cmulc(a[brev(c, 4)], w[c]);
        }
        dit_regs<double, 16>(a);
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i].x *= 0.0625; a[i].y *= 0.0625; }     // keep the values bounded
        ww = cmul(ww, w1);
    }
    cx<double> s = mk<double>(0, 0);
#pragma unroll
    for (int i = 0; i < 16; ++i) s = cadd(s, a[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(int sms, const char* what, double fp64_per_iter) {
    double2* out;
    const int ctas = sms * 2;
    cudaMalloc(&out, sizeof(double2) * ctas * 256);
    mix_kernel<MODE><<<ctas, 256>>>(out, 10, make_double2(0.999, -0.04));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4000;
    cudaEventRecord(e0);
    mix_kernel<MODE><<<ctas, 256>>>(out, iters, make_double2(0.999, -0.04));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double inst = fp64_per_iter * iters * (double)ctas * 256;           // thread-level FP64 instructions (source count)
    printf("%-44s %.3f ms  %.1f FP64 thread-instr/clk/SM (at %d MHz)\n", what, ms, inst / (ms * 1e-3) / (clk * 1e3) / sms, clk / 1000);
    cudaFree(out);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // source-level FP64 instruction counts: radix-16 butterfly 168 each; tree 14 cmul * 4; two twiddle passes 15 * 4 each; scaling 32; ww 4
    run<0>(sms, "stage mix (butterflies + twiddle gen + mults)", 168 * 2 + 56 + 120 + 32 + 4);
    run<1>(sms, "butterflies only", 168 * 2 + 32 + 4);
    return 0;
}
