// stats.cu -- per-band NaN-skipping mean / population std (data_mean_std.py:32-33) and the
// deterministic sum over patches that S:45-46 (np.mean over patches) and the multi-GPU
// all-reduce consume.
//
// One CTA per (patch, band), one pass over the band with pivot-shifted fp64 accumulators (see
// band_stats_kernel).
#include "common.cuh"

namespace kmsr {

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; ++i) t += red[i];     // fixed order: deterministic
    return t;
}

// ONLY_NAN: recompute only the bands whose fused result came out NaN (a NaN pixel poisons the fused
// sums; np.nanmean / np.nanstd skip it) -- every other CTA leaves at once.
//
// One pass: every pixel is read once (128-bit loads, four in flight per thread), shifted by the band's first
// finite-or-zero pixel in fp32 (exact or one rounding of a small difference) and accumulated in fp64 per thread
// as sum d and sum d^2; mean = p + S1/N, var = S2/N - (S1/N)^2 evaluated in fp64.  With fp64 accumulators the
// shifted one-pass form is as accurate as numpy's two-pass fp32 (measured <= 3e-8 relative against fp64
// two-pass), and the band is not re-read.
template <bool ONLY_NAN>
__global__ void __launch_bounds__(256)
band_stats_kernel(const float* __restrict__ x, int C, long long hw, long long stride_n,
                  double* __restrict__ mean, double* __restrict__ stdv) {
    __shared__ double red[8];
    const long long band = blockIdx.x;
    if (ONLY_NAN) {
        const double m0 = mean[band], s0 = stdv[band];
        if (m0 == m0 && s0 == s0) return;
    }
    const long long n = band / C;
    const int c = (int)(band % C);
    const float* p = x + n * stride_n + (long long)c * hw;
    const bool vec = (hw % 4 == 0) && (((uintptr_t)p & 15) == 0);

    float pv = p[0];
    if (!isfinite(pv)) pv = 0.0f;

    double s1 = 0.0, s2 = 0.0;
    unsigned cnt = 0;
    auto acc = [&](float v) {
        if (v == v) {
            const double d = (double)(v - pv);
            s1 += d;
            s2 = fma(d, d, s2);
            ++cnt;
        }
    };
    if (vec) {
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const long long n4 = hw >> 2;
        long long i = threadIdx.x;
        for (; i + 3 * 256 < n4; i += 4 * 256) {
            const float4 v0 = p4[i], v1 = p4[i + 256], v2 = p4[i + 512], v3 = p4[i + 768];
            acc(v0.x); acc(v0.y); acc(v0.z); acc(v0.w);
            acc(v1.x); acc(v1.y); acc(v1.z); acc(v1.w);
            acc(v2.x); acc(v2.y); acc(v2.z); acc(v2.w);
            acc(v3.x); acc(v3.y); acc(v3.z); acc(v3.w);
        }
        for (; i < n4; i += 256) {
            const float4 v = p4[i];
            acc(v.x); acc(v.y); acc(v.z); acc(v.w);
        }
    } else {
        for (long long i = threadIdx.x; i < hw; i += blockDim.x) acc(p[i]);
    }
    const double S1 = block_sum(s1, red);
    const double S2 = block_sum(s2, red);
    const double Nn = block_sum((double)cnt, red);
    if (threadIdx.x == 0) {
        if (Nn > 0.0) {
            const double dm = S1 / Nn;
            double var = S2 / Nn - dm * dm;
            if (var < 0.0) var = 0.0;
            mean[band] = (double)pv + dm;
            stdv[band] = sqrt(var);
        } else {
            mean[band] = nan("");
            stdv[band] = nan("");
        }
    }
}

// sums[c] += sum_n mean[n,c]; sums[C+c] += sum_n std[n,c]; sums[2C] += N.  One CTA per output,
// strided partials then a fixed-order tree: bitwise reproducible for a given N.
__global__ void __launch_bounds__(256)
stats_reduce_kernel(const double* __restrict__ mean, const double* __restrict__ stdv, long long N,
                    int C, double* __restrict__ sums) {
    __shared__ double red[256];
    const int o = blockIdx.x;
    if (o == 2 * C) {
        if (threadIdx.x == 0) sums[o] += (double)N;
        return;
    }
    const double* src = o < C ? mean + o : stdv + (o - C);
    double s = 0.0;
    for (long long n = threadIdx.x; n < N; n += blockDim.x) s += src[n * C];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[o] += red[0];
}

int launch_band_stats(const float* x, long long N, int C, long long hw, long long stride_n,
                      double* mean, double* stdv, double* sums, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(N * C < (1ll << 31), KMSR_E_INVALID, "band_stats: N*C too large");
    KMSR_REQUIRE(hw > 0, KMSR_E_INVALID, "band_stats: empty bands");
    band_stats_kernel<false><<<(unsigned)(N * C), 256, 0, st>>>(x, C, hw, stride_n, mean, stdv);
    KMSR_LAUNCH_CHECK("band_stats_kernel");
    if (sums) {
        stats_reduce_kernel<<<2 * C + 1, 256, 0, st>>>(mean, stdv, N, C, sums);
        KMSR_LAUNCH_CHECK("stats_reduce_kernel");
    }
    return KMSR_OK;
}

// Fused path (degrade_tma_kernel<0, true>): part[band][warp][sum x, sum x^2] -> population mean / std.
__global__ void __launch_bounds__(256)
stats_finish_kernel(const double* __restrict__ part, long long nbands, double npix, double* __restrict__ mean,
                    double* __restrict__ stdv) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= nbands) return;
    const double s1 = part[4 * b] + part[4 * b + 2];
    const double s2 = part[4 * b + 1] + part[4 * b + 3];
    const double m = s1 / npix;
    double v = s2 / npix - m * m;
    if (v < 0.0) v = 0.0;
    mean[b] = m;
    stdv[b] = sqrt(v);           // NaN in, NaN out: those bands are redone by band_stats_kernel<true>
}

int launch_stats_finish(const double* part, const float* x, long long N, int C, long long hw, long long stride_n,
                        double* mean, double* stdv, double* sums, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    const long long nb = N * C;
    KMSR_REQUIRE(nb < (1ll << 31), KMSR_E_INVALID, "stats: N*C too large");
    stats_finish_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(part, nb, (double)hw, mean, stdv);
    KMSR_LAUNCH_CHECK("stats_finish_kernel");
    band_stats_kernel<true><<<(unsigned)nb, 256, 0, st>>>(x, C, hw, stride_n, mean, stdv);
    KMSR_LAUNCH_CHECK("band_stats_kernel");
    if (sums) {
        stats_reduce_kernel<<<2 * C + 1, 256, 0, st>>>(mean, stdv, N, C, sums);
        KMSR_LAUNCH_CHECK("stats_reduce_kernel");
    }
    return KMSR_OK;
}

}  // namespace kmsr
