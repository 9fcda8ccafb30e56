// reuse_probe.cu -- FFMA2 rate against the length of the run of instructions that share one source operand
// (the operand-reuse cache) and against the resident warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o scratch/reuse_probe scratch/reuse_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float lo2(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
constexpr int CH = 16, NW = 16;
template <int RUN>
__global__ void __launch_bounds__(512, 1) probe(float* sink, int iters, float w0) {
    u64 A[CH], D[CH], W[NW];
    for (int i = 0; i < CH; ++i) { A[i] = pack2(threadIdx.x * 0.001f + i, i); D[i] = pack2(1.0f + threadIdx.x * 1e-3f + i, 0.5f - i); }
    for (int r = 0; r < NW; ++r) W[r] = pack2(w0 + r * 0.01f, w0 - r * 0.01f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) A[i] = fma2(W[(i / RUN + r * (CH / RUN)) % NW], D[(i + r) % CH], A[i]);
    }
    float s = 0.0f;
    for (int i = 0; i < CH; ++i) s += lo2(A[i]);
    if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int RUN>
void run(int threads, float* sink) {
    int sms = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    probe<RUN><<<sms, threads>>>(sink, 100, 0.5f);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); probe<RUN><<<sms, threads>>>(sink, iters, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double fl = 2.0 * 2.0 * (double)sms * threads * iters * 4 * CH;
    printf("run %2d  warps/scheduler %d  %7.3f ms  %6.1f TFLOP/s\n", RUN, threads / 128, best, fl / best / 1e9);
}
int main() {
    float* sink; cudaMalloc(&sink, 148 * 512 * 4);
    for (int t : {512, 256, 128}) { run<16>(t, sink); run<8>(t, sink); run<4>(t, sink); run<2>(t, sink); run<1>(t, sink); }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
