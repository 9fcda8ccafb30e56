// probe.cu -- in-process FP32 peak probe: what the SMs of THIS device sustain in packed FFMA2 (fma.rn.f32x2), the
// instruction the degrade kernels issue.  bench.py times one launch with CUDA events and reports it as
// roofline.fp32_peak_tflops, the denominator of roofline.fp32_frac (BASELINE.json north_star: "as a fraction of
// HBM-bandwidth and FP32-FMA rooflines").  Not part of the hot path.
#include "common.cuh"
#include "tma_util.cuh"

namespace kmsr {

namespace {

constexpr int kProbeThreads = 512;       // 4 warps per scheduler
constexpr int kProbeChains = 12;         // independent accumulator pairs per thread
constexpr int kProbeRounds = 4;

__global__ void __launch_bounds__(kProbeThreads, 1) ffma2_probe_kernel(float* sink, int iters, float w0) {
    float a[2 * kProbeChains], w[8], d[8];
#pragma unroll
    for (int i = 0; i < 2 * kProbeChains; ++i) a[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = w0 + i * 0.01f; d[i] = 1.0f + threadIdx.x * 1e-3f + i; }
    u64* A = reinterpret_cast<u64*>(a);
    const u64* W = reinterpret_cast<const u64*>(w);
    const u64* D = reinterpret_cast<const u64*>(d);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < kProbeRounds; ++r)
#pragma unroll
            for (int i = 0; i < kProbeChains; ++i) A[i] = fma2(W[(i + r) & 3], D[(i + 2 * r) & 3], A[i]);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 2 * kProbeChains; ++i) s += a[i];
    if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;      // keeps the chains alive, writes nothing
}

}  // namespace

int launch_fp32_probe(float* sink, int iters, double* fma_count, cudaStream_t st) {
    int dev = 0, sms = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ffma2_probe_kernel<<<sms, kProbeThreads, 0, st>>>(sink, iters, 0.5f);
    KMSR_LAUNCH_CHECK("ffma2_probe_kernel");
    if (fma_count) *fma_count = (double)sms * kProbeThreads * (double)iters * (2.0 * kProbeChains * kProbeRounds);
    return KMSR_OK;
}

}  // namespace kmsr
