// ffma2_probe: FFMA vs FFMA2 (fma.rn.f32x2) issue throughput per SM on sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float ffma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
template <int PACKED>
__global__ void k(float* out, int iters, float w0) {
    float a[24], w[8], d[8];
#pragma unroll
    for (int i = 0; i < 24; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = w0 + i * 0.01f; d[i] = 1.0f + threadIdx.x * 1e-3f + i; }
    for (int it = 0; it < iters; ++it) {
        if (PACKED) {
            u64* A = reinterpret_cast<u64*>(a); u64* W = reinterpret_cast<u64*>(w); u64* Dd = reinterpret_cast<u64*>(d);
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 12; ++i) A[i] = ffma2(W[(i + r) & 3], Dd[(i + 2 * r) & 3], A[i]);
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 24; ++i) a[i] = ffma1(w[(i + r) & 7], d[(i + 2 * r) & 7], a[i]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* o; cudaMalloc(&o, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int packed = 0; packed < 2; ++packed) {
            float ms = 0;
            for (int r = 0; r < 3; ++r) {
                cudaEventRecord(e0);
                if (packed) k<1><<<148, threads>>>(o, iters, 0.5f); else k<0><<<148, threads>>>(o, iters, 0.5f);
                cudaEventRecord(e1); cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1);
            }
            double fma = 148.0 * threads * iters * 96.0;
            printf("threads/SM %4d %s: %.3f ms  %.1f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", threads, packed ? "FFMA2" : "FFMA ", ms,
                   fma / ms / 1e9, fma / ms / 1e3 / 148 / 1.965e3 / 1e3 * 1e3 / 1e3);
        }
    }
    return 0;
}
