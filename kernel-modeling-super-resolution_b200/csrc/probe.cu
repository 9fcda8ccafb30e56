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

// The operand pattern is the degrade kernels': one weight pair held against a run of (pixel pair, accumulator pair)
// operands, so the weight comes from the operand-reuse cache and an FFMA2 reads two 64-bit register operands.  With
// three distinct register-pair operands per instruction the register file sustains one FFMA2 per THREE clocks, not
// two (measured: 49.4 instead of 73 TFLOP/s, SM clock 1965 MHz, 515 W -- not the power cap): a probe written that way
// under-reports the roofline.
__global__ void __launch_bounds__(kProbeThreads, 1) ffma2_probe_kernel(float* sink, int iters, float w0) {
    u64 A[kProbeChains], D[kProbeChains], W[kProbeRounds];
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i) {
        A[i] = pack2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
        D[i] = pack2(1.0f + threadIdx.x * 1e-3f + i, 0.5f - threadIdx.x * 1e-3f - i);
    }
#pragma unroll
    for (int r = 0; r < kProbeRounds; ++r) W[r] = pack2(w0 + r * 0.01f, w0 - r * 0.01f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < kProbeRounds; ++r)
#pragma unroll
            for (int i = 0; i < kProbeChains; ++i) A[i] = fma2(W[r], D[(i + r) % kProbeChains], A[i]);
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < kProbeChains; ++i) s += lo2(A[i]) + hi2(A[i]);
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
