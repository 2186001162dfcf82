// Fused BLAS-1 kernels for the cgs / cg iterations (K6): the reference's scalar-indexed
// loops (/root/reference/src/iterativeLinearSolvers.jl:168-177, 185-188, 192-194, 71-73,
// 81-84) and its dot/norm calls (:136,152,163,184,198) as a handful of single-sweep
// kernels with warp-shuffle reductions accumulated in double precision.
#pragma once
#include "fft_core.cuh"

namespace tb {

template <typename T> struct VT;
template <> struct VT<float> { using R = float; static constexpr bool cplx = false; };
template <> struct VT<double> { using R = double; static constexpr bool cplx = false; };
template <> struct VT<float2> { using R = float; static constexpr bool cplx = true; };
template <> struct VT<double2> { using R = double; static constexpr bool cplx = true; };

// scalar (stored as double2) times vector element, in the element's precision
template <typename T> __device__ __forceinline__ T s_mul(double2 s, T v);
template <> __device__ __forceinline__ float s_mul<float>(double2 s, float v) { return (float)s.x * v; }
template <> __device__ __forceinline__ double s_mul<double>(double2 s, double v) { return s.x * v; }
template <> __device__ __forceinline__ float2 s_mul<float2>(double2 s, float2 v) {
    return cmul(make_float2((float)s.x, (float)s.y), v);
}
template <> __device__ __forceinline__ double2 s_mul<double2>(double2 s, double2 v) { return cmul(s, v); }

__device__ __forceinline__ float v_add(float a, float b) { return a + b; }
__device__ __forceinline__ double v_add(double a, double b) { return a + b; }
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 v_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float v_sub(float a, float b) { return a - b; }
__device__ __forceinline__ double v_sub(double a, double b) { return a - b; }
__device__ __forceinline__ float2 v_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 v_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// conj(a)*b accumulated into (re, im) doubles  (Julia's dot conjugates its first argument)
__device__ __forceinline__ void dot_acc(float a, float b, double& re, double& im) { re += (double)a * (double)b; (void)im; }
__device__ __forceinline__ void dot_acc(double a, double b, double& re, double& im) { re = fma(a, b, re); (void)im; }
__device__ __forceinline__ void dot_acc(float2 a, float2 b, double& re, double& im) {
    re += (double)a.x * b.x + (double)a.y * b.y;
    im += (double)a.x * b.y - (double)a.y * b.x;
}
__device__ __forceinline__ void dot_acc(double2 a, double2 b, double& re, double& im) {
    re = fma(a.x, b.x, fma(a.y, b.y, re));
    im = fma(a.x, b.y, fma(-a.y, b.x, im));
}

// Sum of `count` partial records of four doubles by one CTA of 256 threads, in a fixed order (thread t takes records
// t, t + 256, ...; then lanes, then warps): bit-for-bit reproducible, and ~100x less latency than one thread walking
// the records through L2.  `red` is the caller's [4][8] shared scratch; the result is valid in thread 0.
__device__ __forceinline__ double4 final_sum4(const double* __restrict__ partials, unsigned count, double (*red)[8]) {
    double v[4] = {0, 0, 0, 0};
    for (unsigned b = threadIdx.x; b < count; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += __ldcg(&partials[(size_t)b * 4 + k]);
    }
    __syncthreads();                       // the caller's earlier use of `red` is over
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    double4 s = make_double4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int w = 0; w < 8; ++w) { s.x += red[0][w]; s.y += red[1][w]; s.z += red[2][w]; s.w += red[3][w]; }
    }
    return s;
}

// out[0..1] = dot(a1, b1), out[2..3] = dot(a2, b2) (second pair optional).  Deterministic:
// per-CTA partials, the last CTA to finish sums them in index order.
template <typename T>
__global__ void __launch_bounds__(256)
k_dot2(const T* __restrict__ a1, const T* __restrict__ b1, const T* __restrict__ a2, const T* __restrict__ b2,
       long long n, double* __restrict__ partials, unsigned* __restrict__ counter, double* __restrict__ out) {
    double acc[4] = {0, 0, 0, 0};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        dot_acc(a1[i], b1[i], acc[0], acc[1]);
        if (a2) dot_acc(a2[i], b2[i], acc[2], acc[3]);
    }
    __shared__ double red[4][8];
    __shared__ bool is_last;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], o);
        if ((threadIdx.x & 31) == 0) red[v][threadIdx.x >> 5] = acc[v];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int v = 0; v < 4; ++v) {
            double s = 0;
            for (int w = 0; w < 8; ++w) s += red[v][w];
            partials[(size_t)blockIdx.x * 4 + v] = s;
        }
        __threadfence();
        const unsigned done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        const double4 s = final_sum4(partials, gridDim.x, red);
        if (threadIdx.x == 0) { out[0] = s.x; out[1] = s.y; out[2] = s.z; out[3] = s.w; *counter = 0; }
    }
}

// cgs direction update :168-177.  first: u = r; p = u.  else u = r + beta*q; p = u + beta*(q + beta*p)
template <typename T>
__global__ void k_cgs_dir(T* __restrict__ u, T* __restrict__ p, const T* __restrict__ r, const T* __restrict__ q,
                          double2 beta, int first, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (first) { const T rv = r[i]; u[i] = rv; p[i] = rv; }
        else {
            const T qv = q[i];
            const T uv = v_add(r[i], s_mul<T>(beta, qv));
            u[i] = uv;
            p[i] = v_add(uv, s_mul<T>(beta, v_add(qv, s_mul<T>(beta, p[i]))));
        }
    }
}
// :185-188  q = u - alpha*vh ; uh = u + q
template <typename T>
__global__ void k_cgs_q(T* __restrict__ q, T* __restrict__ uh, const T* __restrict__ u, const T* __restrict__ vh,
                        double2 alpha, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T uv = u[i];
        const T qv = v_sub(uv, s_mul<T>(alpha, vh[i]));
        q[i] = qv;
        uh[i] = v_add(uv, qv);
    }
}
// :192-194  x += alpha*y
template <typename T>
__global__ void k_axpy(T* __restrict__ x, double2 alpha, const T* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        x[i] = v_add(x[i], s_mul<T>(alpha, y[i]));
}
// cg :69-76  p = z + beta*p  (first: p = z)
template <typename T>
__global__ void k_cg_dir(T* __restrict__ p, const T* __restrict__ z, double2 beta, int first, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = first ? z[i] : v_add(z[i], s_mul<T>(beta, p[i]));
}
// cg :81-84  x += alpha*p ; r -= alpha*q
template <typename T>
__global__ void k_cg_upd(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p, const T* __restrict__ q,
                         double2 alpha, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        x[i] = v_add(x[i], s_mul<T>(alpha, p[i]));
        r[i] = v_sub(r[i], s_mul<T>(alpha, q[i]));
    }
}

}  // namespace tb
