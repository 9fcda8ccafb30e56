// denoise.cu -- the upstream denoise stage: denoise/denoise.py:34-65
// `denoise_band_float_nlm` = NaN fill with the band's nanmean (:43-44), skimage `estimate_sigma` (:47: MAD of
// the finest diagonal db2 detail coefficients), skimage `denoise_nl_means(fast_mode=True, patch_size=7,
// patch_distance=11, h=h_factor*sigma, sigma=sigma)` (:56-63), NaN restore (:66).  Its output is the
// `denoised` group D_build_noise_pool.py:85 and E_make_train_data.py:234 read.
//
// The arithmetic of skimage / PyWavelets is restated from their published algorithms (DESIGN.md 4.7 has the
// derivation and the PARITY-UNPINNED note: neither package is in this image).  For every output pixel p
// the integral-image scatter algorithm of skimage's fast mode reduces to a gather:
//     D_t(p) = sum_{a,b = -2..3} (P(p+(a,b)) - P(p+(a,b)+t))^2 - 36 * 2 sigma^2        (6 x 6: 2*offset rows)
//     w_t(p) = exp(-max(D_t, 0) / (49 h^2))  unless that distance exceeds 5;  w_0 = 2
//     out(p) = sum_t w_t(p) P(p+t) / sum_t w_t(p),   t in [-11, 11]^2,  P = reflect-padded band
// 529 shifts x ~20 flops per pixel: the kernel is FP32-bound (1.1 kflop per byte), not HBM-bound.
//
// nlm_kernel: one CTA per 64 x 64 output tile of one band, 256 threads, a 4 x 4 register tile per thread.
//  * the tile plus its 13 / 14 pixel halo (91 x 91, reflect indexing and NaN fill applied while loading) sits
//    in shared memory FOUR times, copy j shifted left by j floats: whatever the column shift, a lane's 9-float
//    window of a shifted row starts on a 16-byte boundary of one of the copies -- three conflict-free LDS.128
//    per row and shift, no misaligned or scalar shared-memory traffic;
//  * a thread keeps its own 9 x 9 window (the pixels its 16 patches cover) in registers for all 529 shifts and
//    streams the shifted window row by row: 9 differences, the four 6-wide row sums from shared partial sums
//    (13 flops instead of 24), then the four 6-tall column sums the same way -- every squared difference is
//    computed (9 x 9) / 16 = 5 times per pixel instead of 36;
//  * weights: one FFMA folds the bias correction, 1 / (49 h^2) and log2(e); FMNMX, EX2, the cut-off select,
//    and two accumulations.  The weighted sum is accumulated around the pixel's own value
//    (sum w (P(p+t) - P(p))), so fp32 accumulation error scales with the noise, not with the radiance level.
#include "common.cuh"

namespace kmsr {

int launch_band_stats(const float*, long long, int, long long, long long, double*, double*, double*, cudaStream_t);

namespace {

constexpr int kOff = 3;                          // patch_size 7 -> offset 3 (denoise.py:204 hard-codes 7 / 11)
constexpr int kDmax = 11;                        // largest patch_distance the tile halo covers
constexpr int kTile = 64;
constexpr int kHaloL = kOff - 1 + kDmax;         // 13 pixels left / above
constexpr int kHaloR = kOff + kDmax;             // 14 pixels right / below
constexpr int kRows = kTile + kHaloL + kHaloR;   // 91
constexpr int kPitch = 96;
constexpr int kCopyF = kRows * kPitch;
constexpr size_t kNlmSmem = (size_t)4 * kCopyF * sizeof(float);   // 139 776 B
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2_approx(float z) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(z));
    return r;
}

// numpy.pad(mode='reflect') index: mirror without repeating the edge sample
__device__ __forceinline__ int reflecti(int i, int n) {
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}
// PyWavelets mode 'symmetric': half-sample mirror, x[-1] = x[0]
__device__ __forceinline__ int symi(int i, int n) {
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// |dd| of pywt.dwtn(image, 'db2'): high-pass along axis 0, then along axis 1 (out[o] = sum_j dec_hi[j] ext[2o+1-j]).
// One thread per coefficient; NaN pixels read as the band's nanmean (denoise.py:43-44).
// Two evaluations per coefficient.  (a) The VALUE: the filter has zero DC gain, so the band mean is subtracted from
// every pixel first and the float32 rounding of the 16-term sum scales with the noise instead of the radiance level
// (float32 pywt at level 80 carries ~6e-5 relative noise on a sigma of 0.05).  (b) The ZERO TEST of skimage's
// `detail_coeffs[np.nonzero(detail_coeffs)]`: the float32-rounded taps do not sum to zero, so in the reference a
// constant region (a NaN block filled with the mean) yields tiny NON-zero coefficients that stay in the median,
// while (a) yields exact zeros there.  The unshifted sum is therefore evaluated too, in float32 without contraction
// and in pywt's order, only to decide whether the coefficient counts; one that counts but has value 0 is stored as
// the smallest positive magnitude.
__global__ void __launch_bounds__(256)
dwt_dd_abs_kernel(const float* __restrict__ x, int C, int H, int W, long long stride_n, const double* __restrict__ mean,
                  float* __restrict__ dd, int Ho, int Wo, int tiles_x, int tiles) {
    const long long band = blockIdx.x / tiles;
    const int tile = (int)(blockIdx.x - band * tiles);
    const int oy = (tile / tiles_x) * 16 + (threadIdx.x >> 4), ox = (tile % tiles_x) * 16 + (threadIdx.x & 15);
    if (oy >= Ho || ox >= Wo) return;
    const long long n = band / C;
    const float* p = x + n * stride_n + (band - n * C) * (long long)H * W;
    const float fill = (float)mean[band];
    const float f[4] = {-0.4829629131445341f, 0.8365163037378079f, -0.2241438680420134f, -0.12940952255126037f};
    int rows[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) rows[i] = symi(2 * oy + 1 - i, H);
    float acc = 0.0f, raw = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int col = symi(2 * ox + 1 - j, W);
        float s = 0.0f, r = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = p[(long long)rows[i] * W + col];
            if (v != v) v = fill;
            s = fmaf(f[i], v - fill, s);
            r = __fadd_rn(r, __fmul_rn(f[i], v));
        }
        acc = fmaf(f[j], s, acc);
        raw = __fadd_rn(raw, __fmul_rn(f[j], r));
    }
    float mag = fabsf(acc);
    if (raw == 0.0f) mag = 0.0f;
    else if (mag == 0.0f) mag = 1.0e-37f;
    dd[band * (long long)Ho * Wo + (long long)oy * Wo + ox] = mag;
}

// sigma = median(|dd| != 0) / norm.ppf(0.75) (skimage _sigma_est_dwt): one CTA per band, 8-bit radix select on the
// float bits (non-negative floats order like unsigned integers).  All-NaN bands report 0.0 (denoise.py:40-41),
// bands whose coefficients are all zero report NaN (np.median of an empty array).
__global__ void __launch_bounds__(1024)
sigma_median_kernel(const float* __restrict__ dd, long long n, const double* __restrict__ mean, double* __restrict__ sigma) {
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_rank, s_zeros, s_nans;
    const long long band = blockIdx.x;
    const unsigned* v = reinterpret_cast<const unsigned*>(dd + band * n);
    const double m = mean[band];
    if (m != m) {
        if (threadIdx.x == 0) sigma[band] = 0.0;
        return;
    }
    if (threadIdx.x == 0) { s_zeros = 0; s_nans = 0; }
    __syncthreads();
    unsigned z = 0, q = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned b = v[i];
        z += b == 0u;
        q += b > 0x7f800000u;
    }
    if (z) atomicAdd(&s_zeros, z);
    if (q) atomicAdd(&s_nans, q);
    __syncthreads();
    const unsigned zeros = s_zeros, nans = s_nans;
    const long long nz = n - zeros;
    if (nans > 0 || nz == 0) {
        if (threadIdx.x == 0) sigma[band] = nan("");
        return;
    }
    float val[2];
    for (int which = 0; which < 2; ++which) {
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_rank = zeros + (unsigned)(which == 0 ? (nz - 1) / 2 : nz / 2);
        }
        unsigned mask = 0;
        for (int shift = 24; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix;
            for (long long i = threadIdx.x; i < n; i += blockDim.x) {
                const unsigned b = v[i];
                if ((b & mask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned cum = 0, rank = s_rank;
                for (int b = 0; b < 256; ++b) {
                    const unsigned hb = hist[b];
                    if (rank < cum + hb) {
                        s_prefix = prefix | ((unsigned)b << shift);
                        s_rank = rank - cum;
                        break;
                    }
                    cum += hb;
                }
            }
            mask |= 255u << shift;
            __syncthreads();
        }
        val[which] = __uint_as_float(s_prefix);
        __syncthreads();
    }
    if (threadIdx.x == 0) sigma[band] = (double)((val[0] + val[1]) / 2.0f) / 0.6744897501960817;
}

__global__ void __launch_bounds__(256, 1)
nlm_kernel(const float* __restrict__ x, int C, int H, int W, long long stride_n, const double* __restrict__ mean,
           const double* __restrict__ sigma, double h_factor, int d, float* __restrict__ out, int tiles_x, int tiles) {
    extern __shared__ __align__(16) float sm[];
    const long long band = blockIdx.x / tiles;
    const int tile = (int)(blockIdx.x - band * tiles);
    const int y0 = (tile / tiles_x) * kTile, x0 = (tile % tiles_x) * kTile;
    const long long n = band / C;
    const float* p = x + n * stride_n + (band - n * C) * (long long)H * W;
    float* o = out + band * (long long)H * W;
    const float fill = (float)mean[band];

    // tile + halo -> four shifted copies
    for (int e = threadIdx.x; e < kRows * kRows; e += 256) {
        const int r = e / kRows, c = e - r * kRows;
        float v = p[(long long)reflecti(y0 - kHaloL + r, H) * W + reflecti(x0 - kHaloL + c, W)];
        if (v != v) v = fill;
        float* dst = sm + r * kPitch + c;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (c >= j) dst[j * kCopyF - j] = v;
    }
    __syncthreads();

    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    // skimage casts h and var = sigma^2 to the image dtype; h2s2 = h^2 * 49 in float32
    const double sg = sigma[band];
    const float h = (float)(h_factor * sg), var = (float)(sg * sg);
    const float h2s2 = (h * h) * 49.0f;
    const float inv = 1.0f / h2s2;
    const float k1 = -inv * kLog2e;                         // z = -(D / h2s2) log2 e
    const float k0 = (36.0f * (2.0f * var)) * inv * kLog2e;
    const float zcut = -5.0f * kLog2e;                      // DISTANCE_CUTOFF
    const bool bad = !(h2s2 > 0.0f) || !(h2s2 < 3.0e38f);   // sigma NaN / 0: the reference's arithmetic yields NaN

    // own 9 x 9 window: tile rows 4ly + 11 .. + 19, columns 4lx + 11 .. + 19 = copy 3, columns 4lx + 8 ..
    float own[9][9];
    {
        const float* b3 = sm + 3 * kCopyF + (4 * ly + kHaloL - 2) * kPitch + 4 * lx + 8;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(b3 + i * kPitch);
            const float4 b = *reinterpret_cast<const float4*>(b3 + i * kPitch + 4);
            const float c = b3[i * kPitch + 8];
            own[i][0] = a.x; own[i][1] = a.y; own[i][2] = a.z; own[i][3] = a.w;
            own[i][4] = b.x; own[i][5] = b.y; own[i][6] = b.z; own[i][7] = b.w; own[i][8] = c;
        }
    }
    float sw[16], sv[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { sw[k] = 0.0f; sv[k] = 0.0f; }

#pragma unroll 1
    for (int ty = -d; ty <= d; ++ty) {
        const float* rowbase = sm + (4 * ly + kHaloL - 2 + ty) * kPitch + 4 * lx;
#pragma unroll 1
        for (int tx = -d; tx <= d; ++tx) {
            const int s = tx + kHaloL - 2;                   // 0 .. 22: first column of the shifted window, minus 4lx
            const float* base = rowbase + (s & 3) * kCopyF + (s & ~3);
            float hs[9][4];                                  // 6-wide row sums of squared differences
            float ctr[16];                                   // P(p) - P(p+t) at the 16 output pixels
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const float4 a = *reinterpret_cast<const float4*>(base + i * kPitch);
                const float4 b = *reinterpret_cast<const float4*>(base + i * kPitch + 4);
                const float4 c = *reinterpret_cast<const float4*>(base + i * kPitch + 8);
                const float d0 = own[i][0] - a.x, d1 = own[i][1] - a.y, d2 = own[i][2] - a.z, d3 = own[i][3] - a.w;
                const float d4 = own[i][4] - b.x, d5 = own[i][5] - b.y, d6 = own[i][6] - b.z, d7 = own[i][7] - b.w;
                const float d8 = own[i][8] - c.x;
                if (i >= 2 && i <= 5) {
                    ctr[(i - 2) * 4 + 0] = d2; ctr[(i - 2) * 4 + 1] = d3; ctr[(i - 2) * 4 + 2] = d4; ctr[(i - 2) * 4 + 3] = d5;
                }
                // h_x = sum_{k = x .. x+5} d_k^2 for x = 0 .. 3, from the shared partial sums
                const float e2 = d2 * d2, e6 = d6 * d6;
                const float cc = fmaf(d5, d5, fmaf(d4, d4, d3 * d3));
                const float t = cc + fmaf(d1, d1, e2);
                const float u = cc + fmaf(d7, d7, e6);
                hs[i][0] = fmaf(d0, d0, t);
                hs[i][1] = t + e6;
                hs[i][2] = u + e2;
                hs[i][3] = fmaf(d8, d8, u);
            }
#pragma unroll
            for (int xx = 0; xx < 4; ++xx) {
                // v_y = sum_{i = y .. y+5} hs[i][xx] for y = 0 .. 3
                const float cc = hs[3][xx] + hs[4][xx] + hs[5][xx];
                const float t = cc + (hs[1][xx] + hs[2][xx]);
                const float u = cc + (hs[6][xx] + hs[7][xx]);
                const float v[4] = {t + hs[0][xx], t + hs[6][xx], u + hs[2][xx], u + hs[8][xx]};
#pragma unroll
                for (int yy = 0; yy < 4; ++yy) {
                    const float z = fminf(fmaf(v[yy], k1, k0), 0.0f);
                    float w = ex2_approx(z);
                    w = z < zcut ? 0.0f : w;
                    sw[yy * 4 + xx] += w;
                    sv[yy * 4 + xx] = fmaf(-w, ctr[yy * 4 + xx], sv[yy * 4 + xx]);
                }
            }
        }
    }

    // the zero shift is written twice by the scatter algorithm (weight 2): one more unit of weight, no value term
    const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
        const int gy = y0 + 4 * ly + yy;
        if (gy >= H) continue;
#pragma unroll
        for (int xx = 0; xx < 4; ++xx) {
            const int gx = x0 + 4 * lx + xx;
            if (gx >= W) continue;
            float r = own[yy + 2][xx + 2] + sv[yy * 4 + xx] / (sw[yy * 4 + xx] + 1.0f);
            const float orig = p[(long long)gy * W + gx];
            if (bad || orig != orig) r = qnan;               // denoise.py:66: NaN pixels stay NaN
            o[(long long)gy * W + gx] = r;
        }
    }
}

}  // namespace

long long denoise_workspace(long long N, int C, int H, int W) {
    const long long nb = N * C;
    const long long Ho = (H + 3) / 2, Wo = (W + 3) / 2;
    long long bytes = 0;
    bytes += (nb * 2 * (long long)sizeof(double) + 255) / 256 * 256;   // band mean / std
    bytes += (nb * Ho * Wo * (long long)sizeof(float) + 255) / 256 * 256;
    return bytes + 256;
}

// sigma [N, C]: estimate_sigma of every band (NaN pixels filled with the band's nanmean first)
int launch_estimate_sigma(const float* x, long long N, int C, int H, int W, long long stride_n, double* sigma,
                          void* workspace, long long workspace_bytes, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    const long long nb = N * C;
    KMSR_REQUIRE(workspace_bytes >= denoise_workspace(N, C, H, W), KMSR_E_INVALID,
                 "estimate_sigma: workspace of %lld bytes, %lld needed", workspace_bytes, denoise_workspace(N, C, H, W));
    KMSR_REQUIRE(((uintptr_t)workspace & 255) == 0, KMSR_E_ALIGN, "estimate_sigma: workspace not 256-byte aligned");
    const int Ho = (H + 3) / 2, Wo = (W + 3) / 2;
    double* mean = reinterpret_cast<double*>(workspace);
    double* stdv = mean + nb;
    float* dd = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + (nb * 2 * sizeof(double) + 255) / 256 * 256);
    int rc = launch_band_stats(x, N, C, (long long)H * W, stride_n, mean, stdv, nullptr, st);
    if (rc != KMSR_OK) return rc;
    const int tiles_x = (Wo + 15) / 16, tiles = tiles_x * ((Ho + 15) / 16);
    KMSR_REQUIRE(nb * tiles < (1ll << 31), KMSR_E_INVALID, "estimate_sigma: too many bands");
    dwt_dd_abs_kernel<<<(unsigned)(nb * tiles), 256, 0, st>>>(x, C, H, W, stride_n, mean, dd, Ho, Wo, tiles_x, tiles);
    KMSR_LAUNCH_CHECK("dwt_dd_abs_kernel");
    sigma_median_kernel<<<(unsigned)nb, 1024, 0, st>>>(dd, (long long)Ho * Wo, mean, sigma);
    KMSR_LAUNCH_CHECK("sigma_median_kernel");
    return KMSR_OK;
}

// out [N, C, H, W] = denoise_band_float_nlm of every band with h = h_factor * sigma[band]; `mean` is the band
// nanmean left in the workspace by launch_estimate_sigma.
int launch_nlm(const float* x, long long N, int C, int H, int W, long long stride_n, const double* mean,
               const double* sigma, double h_factor, int patch_distance, float* out, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    const long long nb = N * C;
    const int tiles_x = (W + kTile - 1) / kTile, tiles = tiles_x * ((H + kTile - 1) / kTile);
    KMSR_REQUIRE(nb * tiles < (1ll << 31), KMSR_E_INVALID, "nlm: too many tiles");
    KMSR_CUDA_OK(cudaFuncSetAttribute(nlm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNlmSmem));
    nlm_kernel<<<(unsigned)(nb * tiles), 256, kNlmSmem, st>>>(x, C, H, W, stride_n, mean, sigma, h_factor, patch_distance, out,
                                                              tiles_x, tiles);
    KMSR_LAUNCH_CHECK("nlm_kernel");
    return KMSR_OK;
}

bool nlm_shape_ok(int patch_size, int patch_distance, const char** why) {
    *why = "";
    if (patch_size != 2 * kOff + 1 && patch_size != 2 * kOff) { *why = "patch_size 7 (or 6, which skimage rounds up to 7)"; return false; }
    if (patch_distance < 0 || patch_distance > kDmax) { *why = "patch_distance 0 .. 11"; return false; }
    return true;
}

}  // namespace kmsr
