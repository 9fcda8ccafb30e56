// prepare.cu -- kernel-bank preparation: per-band normalisation + box-mean folding.
//
// Reference semantics restated (kernel_from_lr_gan/C_30apply_kernel_to_landsat.py:93-97,
// C_31apply_muti_kernel_to_landsat.py:74-78): a band of a blur kernel is divided by its own
// sum iff that sum is > 0.  The s x s box mean that the cascaded avg_pool2d(2,2) stages apply
// afterwards (C_30:120-122) is linear, so it is folded into the kernel here:
//     K'[u,v] = (1/s^2) * sum_{a,b<s} kn[u-a, v-b]          (size (kh+s-1) x (kw+s-1))
// and the degrade kernels evaluate K' only at output positions (stride s) -- 27x fewer FMAs than
// blur-then-pool at k=13, s=8.
#include "common.cuh"

namespace kmsr {

// one CTA per (kernel, band)
__global__ void __launch_bounds__(128)
prepare_kernels_kernel(const float* __restrict__ kbank, int kh, int kw, int fold, int KH, int KW,
                       int KWp, float* __restrict__ comp, float* __restrict__ dsum) {
    extern __shared__ __align__(16) float kn[];   // kh*kw normalised taps, then the fp64 horizontal sums
    __shared__ double red[128];
    const int kc = blockIdx.x;
    const float* k = kbank + (size_t)kc * kh * kw;
    const int taps = kh * kw;

    // band sum in fp64 (any fp32 summation order of the reference lies within one ulp of it)
    double s = 0.0;
    for (int i = threadIdx.x; i < taps; i += blockDim.x) s += (double)k[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    const float sf = (float)red[0];
    __syncthreads();

    // fp32 division exactly as kernel[i] / kernel_sum does (C_30:96-97)
    for (int i = threadIdx.x; i < taps; i += blockDim.x) kn[i] = sf > 0.0f ? __fdiv_rn(k[i], sf) : k[i];
    __syncthreads();

    // residual of the normalised sum: the degrade kernels add pivot * (sum - 1) back
    double s2 = 0.0;
    for (int i = threadIdx.x; i < taps; i += blockDim.x) s2 += (double)kn[i];
    red[threadIdx.x] = s2;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) dsum[kc] = (float)(red[0] - 1.0);

    // box fold, separable: horizontal running sums of the normalised taps (fp64), then vertical sums of those --
    // fold + fold additions per composite tap instead of fold * fold (what made a per-patch bank of 4096 x 5
    // kernels cost a fifth of the degrade pass)
    double* hs = reinterpret_cast<double*>(kn + ((taps + 1) & ~1));      // [kh][KW] horizontal sums
    for (int i = threadIdx.x; i < kh * KW; i += blockDim.x) {
        const int r = i / KW, v = i % KW;
        double acc = 0.0;
        for (int b = 0; b < fold; ++b) {
            const int kx = v - b;
            if (kx >= 0 && kx < kw) acc += (double)kn[r * kw + kx];
        }
        hs[i] = acc;
    }
    __syncthreads();
    const double inv = 1.0 / ((double)fold * (double)fold);
    float* out = comp + (size_t)kc * KH * KWp;
    for (int i = threadIdx.x; i < KH * KWp; i += blockDim.x) {
        const int u = i / KWp, v = i % KWp;
        double acc = 0.0;
        if (v < KW) {
            for (int a = 0; a < fold; ++a) {
                const int ky = u - a;
                if (ky >= 0 && ky < kh) acc += hs[ky * KW + v];
            }
        }
        out[i] = (float)(acc * inv);
    }
}

int launch_prepare(const float* kbank, long long nK, int C, int kh, int kw, int factor, int down_mode,
                   float* comp, float* dsum, cudaStream_t st) {
    Geometry g;
    int rc = make_geometry(0, 0, kh, kw, factor, down_mode, &g);
    KMSR_REQUIRE(rc == KMSR_OK, rc, "prepare_kernels: bad kernel geometry kh=%d kw=%d factor=%d mode=%d",
                 kh, kw, factor, down_mode);
    KMSR_REQUIRE(nK >= 0 && C >= 1, KMSR_E_INVALID, "prepare_kernels: nK=%lld C=%d", nK, C);
    if (nK == 0) return KMSR_OK;
    KMSR_REQUIRE(kbank && comp && dsum, KMSR_E_INVALID, "prepare_kernels: null pointer");
    const size_t smem = (size_t)((kh * kw + 1) & ~1) * sizeof(float) + (size_t)kh * g.KW * sizeof(double);
    KMSR_REQUIRE(smem <= 48 * 1024, KMSR_E_UNSUPPORTED,
                 "prepare_kernels: kernel %dx%d exceeds 48 KB of shared memory", kh, kw);
    const int fold = down_mode == KMSR_DOWN_BOXMEAN ? g.stride : 1;
    const long long blocks = nK * C;
    KMSR_REQUIRE(blocks < (1ll << 31), KMSR_E_INVALID, "prepare_kernels: nK*C too large");
    prepare_kernels_kernel<<<(unsigned)blocks, 128, smem, st>>>(
        kbank, kh, kw, fold, g.KH, g.KW, g.KWp, comp, dsum);
    KMSR_LAUNCH_CHECK("prepare_kernels_kernel");
    return KMSR_OK;
}

}  // namespace kmsr
