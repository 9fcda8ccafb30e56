// api.cu -- the extern "C" surface of libkmsr.so (include/kmsr.h): argument validation, error
// reporting, kernel selection.  No torch types, no device allocation, nothing retained.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kmsr {

static thread_local char t_error[512] = "";
static thread_local const char* t_algo = "none";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}
void set_algo(const char* name) { t_algo = name; }

int launch_prepare(const float*, long long, int, int, int, int, int, float*, float*, cudaStream_t);
int launch_add_noise(const float*, long long, int, long long, const float*, const int*, const float*,
                     const int*, float*, cudaStream_t);
int launch_crop_sub(const float*, const float*, int, int, int, const int*, const int*, long long, int,
                    float*, cudaStream_t);
int launch_band_stats(const float*, long long, int, long long, long long, double*, double*, double*,
                      cudaStream_t);
int launch_water_mask(float*, int, long long, int, float, float, float, float*, cudaStream_t);
int launch_stats_finish(const double*, const float*, long long, int, long long, long long, double*, double*,
                        double*, cudaStream_t);
long long keep_mask_workspace(int, int, int, int);
int launch_scene_keep_mask(const float*, int, int, int, int, float, float, float, int, int, double, unsigned char*,
                           int*, void*, long long, cudaStream_t);
int launch_keep_mask(const float*, int, int, int, int, int, double, unsigned char*, int*, void*,
                     long long, cudaStream_t);
long long denoise_workspace(long long, int, int, int);
int launch_estimate_sigma(const float*, long long, int, int, int, long long, double*, void*, long long, cudaStream_t);
int launch_nlm(const float*, long long, int, int, int, long long, const double*, const double*, double, int, float*,
               cudaStream_t);
bool nlm_shape_ok(int, int, const char**);
int launch_fp32_probe(float*, int, double*, cudaStream_t);
int launch_validate_indices(const int*, long long, long long, int*, int*, cudaStream_t);
bool selector_umma_shape_ok(int, int);
long long selector_umma_wfloats(int, int);
long long selector_umma_workspace(long long, int, int);
int launch_selector_umma(const float*, long long, int, int, const float*, const float*, const float*, const float*, const float*,
                         const float*, const float*, const float*, float*, void*, long long, cudaStream_t);
long long selector_wsplit_floats(int, int);
long long selector_workspace(long long, int, int);
int launch_selector(const float*, long long, int, int, const float*, const float*, const float*, const float*, const float*,
                    const float*, const float*, const float*, float*, void*, long long, cudaStream_t);

}  // namespace kmsr

using namespace kmsr;

extern "C" {

KMSR_API int kmsr_version(void) { return KMSR_VERSION; }
KMSR_API const char* kmsr_last_error(void) { return t_error; }
KMSR_API int64_t kmsr_launch_count(void) { return (int64_t)g_launches.load(); }
KMSR_API const char* kmsr_last_algo(void) { return t_algo; }

KMSR_API int kmsr_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                              int64_t* l2_bytes, int64_t* smem_optin_bytes) {
    cudaDeviceProp prop;
    KMSR_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    if (smem_optin_bytes) *smem_optin_bytes = (int64_t)prop.sharedMemPerBlockOptin;
    return KMSR_OK;
}

KMSR_API int kmsr_fp32_probe(float* sink, int iters, double* fma_count, void* stream) {
    KMSR_REQUIRE(sink != nullptr && iters >= 1, KMSR_E_INVALID, "fp32_probe: sink=%p iters=%d", (void*)sink, iters);
    return launch_fp32_probe(sink, iters, fma_count, (cudaStream_t)stream);
}

KMSR_API int kmsr_validate_indices(const int32_t* idx, int64_t n, int64_t upper, int32_t* scratch, const char* what,
                                   void* stream) {
    KMSR_REQUIRE(n >= 0 && upper >= 0, KMSR_E_INVALID, "validate_indices: n=%lld upper=%lld", (long long)n, (long long)upper);
    if (n == 0) return KMSR_OK;
    KMSR_REQUIRE(idx && scratch, KMSR_E_INVALID, "validate_indices: null pointer");
    int bad = 0;
    int rc = launch_validate_indices(idx, n, upper, scratch, &bad, (cudaStream_t)stream);
    if (rc != KMSR_OK) return rc;
    KMSR_REQUIRE(bad == 0, KMSR_E_INVALID, "%s: %d of %lld indices lie outside [0, %lld)", what ? what : "indices", bad,
                 (long long)n, (long long)upper);
    return KMSR_OK;
}

KMSR_API int kmsr_degrade_out_size(int H, int W, int kh, int kw, int factor, int down_mode, int* Ho,
                                   int* Wo) {
    Geometry g;
    int rc = make_geometry(H, W, kh, kw, factor, down_mode, &g);
    KMSR_REQUIRE(rc == KMSR_OK, rc, "degrade_out_size: bad geometry H=%d W=%d k=%dx%d factor=%d mode=%d", H, W,
                 kh, kw, factor, down_mode);
    if (Ho) *Ho = g.Ho;
    if (Wo) *Wo = g.Wo;
    return KMSR_OK;
}

KMSR_API int kmsr_composite_size(int kh, int kw, int factor, int down_mode, int* KH, int* KW, int* stride) {
    Geometry g;
    int rc = make_geometry(0, 0, kh, kw, factor, down_mode, &g);
    KMSR_REQUIRE(rc == KMSR_OK, rc, "composite_size: bad geometry k=%dx%d factor=%d mode=%d", kh, kw, factor,
                 down_mode);
    if (KH) *KH = g.KH;
    if (KW) *KW = g.KWp;      /* row pitch of comp: KW rounded up to 4 floats, zero filled */
    if (stride) *stride = g.stride;
    return KMSR_OK;
}

KMSR_API int64_t kmsr_degrade_workspace_bytes(int64_t nK, int C, int kh, int kw, int factor, int down_mode) {
    Geometry g;
    if (make_geometry(0, 0, kh, kw, factor, down_mode, &g) != KMSR_OK || nK < 0 || C < 1) {
        set_error("degrade_workspace_bytes: bad arguments");
        return KMSR_E_INVALID;
    }
    const int64_t comp = nK * C * (int64_t)g.KH * g.KWp * (int64_t)sizeof(float);
    const int64_t ds = ((nK * C * (int64_t)sizeof(float)) + 255) / 256 * 256;
    return comp + ds + 256;
}

KMSR_API int kmsr_prepare_kernels(const float* kbank, int64_t nK, int C, int kh, int kw, int factor,
                                  int down_mode, float* comp, float* dsum, void* stream) {
    return launch_prepare(kbank, nK, C, kh, kw, factor, down_mode, comp, dsum, (cudaStream_t)stream);
}

static int degrade_common(const float* hr, int64_t N, int C, int H, int W, int64_t hr_stride_n,
                          int64_t hr_stride_c, int64_t hr_stride_h, const int64_t* patch_offsets, int scene_h,
                          int scene_w, int x_multiple,
                          const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                          const int32_t* kidx, const float* sigma, const float* pool,
                          int64_t nPool, const int32_t* nidx, int factor, int pad_mode,
                          int down_mode, int noise_mode, float* lr, int algo, void* stream,
                          double* stat_part = nullptr) {
    DegradeArgs a;
    a.stat_part = stat_part;
    int rc = make_geometry(H, W, kh, kw, factor, down_mode, &a.g);
    KMSR_REQUIRE(rc == KMSR_OK, rc, "degrade: bad geometry H=%d W=%d k=%dx%d factor=%d down_mode=%d", H, W, kh,
                 kw, factor, down_mode);
    KMSR_REQUIRE(N >= 0 && C >= 1, KMSR_E_INVALID, "degrade: N=%lld C=%d", (long long)N, C);
    KMSR_REQUIRE(pad_mode == KMSR_PAD_REPLICATE || pad_mode == KMSR_PAD_ZERO, KMSR_E_INVALID,
                 "degrade: pad_mode %d", pad_mode);
    KMSR_REQUIRE(noise_mode >= KMSR_NOISE_NONE && noise_mode <= KMSR_NOISE_SIGMA, KMSR_E_INVALID,
                 "degrade: noise_mode %d", noise_mode);
    KMSR_REQUIRE(algo >= KMSR_ALGO_AUTO && algo <= KMSR_ALGO_BOX, KMSR_E_INVALID, "degrade: algo %d", algo);
    if (N == 0 || a.g.Ho == 0 || a.g.Wo == 0) return KMSR_OK;
    KMSR_REQUIRE(H >= 1 && W >= 1, KMSR_E_INVALID, "degrade: empty patch %dx%d", H, W);
    KMSR_REQUIRE(hr && comp && dsum && lr, KMSR_E_INVALID, "degrade: null pointer");
    KMSR_REQUIRE(nK >= 1, KMSR_E_INVALID, "degrade: empty kernel bank");
    KMSR_REQUIRE(hr_stride_h >= W, KMSR_E_INVALID, "degrade: row stride %lld < W", (long long)hr_stride_h);
    if (noise_mode != KMSR_NOISE_NONE) {
        KMSR_REQUIRE(pool && nidx && nPool >= 1, KMSR_E_INVALID, "degrade: noise requested without pool / nidx");
        if (noise_mode == KMSR_NOISE_SIGMA)
            KMSR_REQUIRE(sigma != nullptr, KMSR_E_INVALID, "degrade: KMSR_NOISE_SIGMA without sigma");
    }
    a.hr = hr; a.N = N; a.C = C; a.H = H; a.W = W;
    a.sN = hr_stride_n; a.sC = hr_stride_c; a.sH = hr_stride_h;
    a.patch_offsets = (const long long*)patch_offsets;
    a.scene_h = scene_h; a.scene_w = scene_w; a.x_multiple = x_multiple;
    a.comp = comp; a.dsum = dsum; a.nK = nK; a.kidx = kidx;
    a.sigma = sigma; a.pool = pool; a.nPool = nPool; a.nidx = nidx;
    a.pad_mode = pad_mode; a.noise_mode = noise_mode; a.lr = lr;
    cudaStream_t st = (cudaStream_t)stream;
    const char* why = "";
    const bool tma_ok = tma_shape_ok(a, &why);
    if (algo == KMSR_ALGO_TMA) {
        KMSR_REQUIRE(tma_ok, KMSR_E_UNSUPPORTED, "degrade: TMA kernel does not cover this call (%s)", why);
        return launch_degrade_tma(a, st);
    }
    if (algo == KMSR_ALGO_AUTO && tma_ok) return launch_degrade_tma(a, st);
    // only the headline kernel writes the fused-statistics partials: nothing below may be handed a non-null stat_part
    KMSR_REQUIRE(stat_part == nullptr, KMSR_E_UNSUPPORTED, "degrade: fused statistics need the TMA kernel (%s)", why);
    const char* why4 = "";
    const bool box_ok = box_shape_ok(a, down_mode, &why4);
    if (algo == KMSR_ALGO_BOX) {
        KMSR_REQUIRE(box_ok, KMSR_E_UNSUPPORTED, "degrade: box-tile kernel does not cover this call (%s)", why4);
        return launch_degrade_box(a, st);
    }
    const char* why3 = "";
    const bool reg_ok = reg_shape_ok(a, down_mode, &why3);
    // factor 2 / 4 (FP32-bound or near the ridge) and 64-wide patches at any factor go to the box-tile kernel
    const bool band = a.W <= 64 && a.H <= 64;
    if (algo == KMSR_ALGO_AUTO && box_ok && (a.g.stride <= 4 || band) && a.W >= 48 && a.H >= 48)
        return launch_degrade_box(a, st);
    if (algo == KMSR_ALGO_REG) {
        KMSR_REQUIRE(reg_ok, KMSR_E_UNSUPPORTED, "degrade: register-tile kernel does not cover this call (%s)", why3);
        return launch_degrade_reg(a, st);
    }
    // FP32-bound shapes: factor 2 (1.3-2.2x over the streaming kernel on every sweep cell), factor 4 on 64-wide patches
    if (algo == KMSR_ALGO_AUTO && reg_ok && (a.g.stride == 2 || a.W <= 64)) return launch_degrade_reg(a, st);
    const char* why2 = "";
    const bool stream_ok = stream_shape_ok(a, down_mode, &why2);
    if (algo == KMSR_ALGO_STREAM) {
        KMSR_REQUIRE(stream_ok, KMSR_E_UNSUPPORTED, "degrade: streaming kernel does not cover this call (%s)", why2);
        return launch_degrade_stream(a, st);
    }
    if (algo == KMSR_ALGO_AUTO && stream_ok) return launch_degrade_stream(a, st);
    if (algo == KMSR_ALGO_AUTO && reg_ok) return launch_degrade_reg(a, st);      // factor-4 shapes the streaming kernel refuses
    return launch_degrade_tiled(a, st);
}

KMSR_API int kmsr_degrade_prepared(const float* hr, int64_t N, int C, int H, int W, int64_t hr_stride_n,
                                   int64_t hr_stride_c, int64_t hr_stride_h, const int64_t* patch_offsets,
                                   const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                                   const int32_t* kidx, const float* sigma, const float* pool,
                                   int64_t nPool, const int32_t* nidx, int factor, int pad_mode,
                                   int down_mode, int noise_mode, float* lr, int algo, void* stream) {
    return degrade_common(hr, N, C, H, W, hr_stride_n, hr_stride_c, hr_stride_h, patch_offsets, 0, 0, 1, comp, dsum,
                          nK, kh, kw, kidx, sigma, pool, nPool, nidx, factor, pad_mode, down_mode, noise_mode, lr,
                          algo, stream);
}

KMSR_API int kmsr_degrade_windows(const float* scene, int C, int scene_h, int scene_w, int64_t hr_stride_c,
                                  int64_t hr_stride_h, const int64_t* patch_offsets, int64_t N, int H, int W,
                                  int x_multiple, const float* comp, const float* dsum, int64_t nK, int kh,
                                  int kw, const int32_t* kidx, const float* sigma, const float* pool,
                                  int64_t nPool, const int32_t* nidx, int factor, int pad_mode, int down_mode,
                                  int noise_mode, float* lr, int algo, void* stream) {
    KMSR_REQUIRE(scene_h >= H && scene_w >= W && H >= 1 && W >= 1, KMSR_E_INVALID,
                 "degrade_windows: window %dx%d does not fit the scene %dx%d", H, W, scene_h, scene_w);
    KMSR_REQUIRE(hr_stride_h >= scene_w && hr_stride_c >= (int64_t)scene_h * hr_stride_h - (hr_stride_h - scene_w),
                 KMSR_E_INVALID, "degrade_windows: strides (%lld, %lld) do not hold a %dx%d scene",
                 (long long)hr_stride_c, (long long)hr_stride_h, scene_h, scene_w);
    KMSR_REQUIRE(N == 0 || patch_offsets != nullptr, KMSR_E_INVALID, "degrade_windows: null patch_offsets");
    KMSR_REQUIRE(x_multiple >= 1, KMSR_E_INVALID, "degrade_windows: x_multiple %d", x_multiple);
    return degrade_common(scene, N, C, H, W, 0, hr_stride_c, hr_stride_h, patch_offsets, scene_h, scene_w, x_multiple,
                          comp, dsum, nK, kh, kw, kidx, sigma, pool, nPool, nidx, factor, pad_mode, down_mode,
                          noise_mode, lr, algo, stream);
}

KMSR_API int64_t kmsr_degrade_stats_workspace_bytes(int64_t N, int C) {
    if (N < 0 || C < 1) { set_error("degrade_stats_workspace_bytes: N=%lld C=%d", (long long)N, C); return KMSR_E_INVALID; }
    return N * C * 4 * (int64_t)sizeof(double) + 256;
}

KMSR_API int kmsr_degrade_stats_prepared(const float* hr, int64_t N, int C, int H, int W, int64_t hr_stride_n,
                                         const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                                         const int32_t* kidx, const float* sigma, const float* pool,
                                         int64_t nPool, const int32_t* nidx, int factor, int pad_mode,
                                         int down_mode, int noise_mode, float* lr, double* mean, double* std,
                                         double* sums, void* workspace, int64_t workspace_bytes, int algo,
                                         void* stream) {
    KMSR_REQUIRE(mean && std, KMSR_E_INVALID, "degrade_stats: null mean / std");
    const int64_t hw = (int64_t)H * W;
    KMSR_REQUIRE(N <= 1 || hr_stride_n >= (int64_t)C * hw, KMSR_E_INVALID, "degrade_stats: patch stride %lld < C*H*W",
                 (long long)hr_stride_n);
    const int64_t need = kmsr_degrade_stats_workspace_bytes(N, C);
    KMSR_REQUIRE(workspace && workspace_bytes >= need, KMSR_E_INVALID,
                 "degrade_stats: workspace of %lld B needed, %lld given", (long long)need, (long long)workspace_bytes);
    double* part = (double*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    // fused when the TMA kernel takes the call, otherwise the two kernels run back to back (same results contract)
    DegradeArgs probe{};
    probe.stat_part = part;
    int rc = make_geometry(H, W, kh, kw, factor, down_mode, &probe.g);
    KMSR_REQUIRE(rc == KMSR_OK, rc, "degrade_stats: bad geometry");
    probe.hr = hr; probe.N = N; probe.C = C; probe.H = H; probe.W = W; probe.sN = hr_stride_n; probe.sC = hw; probe.sH = W;
    probe.patch_offsets = nullptr; probe.scene_h = probe.scene_w = 0; probe.x_multiple = 1; probe.pad_mode = pad_mode;
    const char* why = "";
    const bool fused = (algo == KMSR_ALGO_AUTO || algo == KMSR_ALGO_TMA) && N > 0 && tma_shape_ok(probe, &why);
    rc = degrade_common(hr, N, C, H, W, hr_stride_n, hw, W, nullptr, 0, 0, 1, comp, dsum, nK, kh, kw, kidx, sigma, pool,
                        nPool, nidx, factor, pad_mode, down_mode, noise_mode, lr, algo, stream, fused ? part : nullptr);
    if (rc != KMSR_OK || N == 0) return rc;
    if (fused) return launch_stats_finish(part, hr, N, C, hw, hr_stride_n, mean, std, sums, (cudaStream_t)stream);
    return launch_band_stats(hr, N, C, hw, hr_stride_n, mean, std, sums, (cudaStream_t)stream);
}

KMSR_API int kmsr_degrade_batch(const float* hr, int64_t N, int C, int H, int W, int64_t hr_stride_n,
                                int64_t hr_stride_c, int64_t hr_stride_h, const int64_t* patch_offsets,
                                const float* kbank, int64_t nK, int kh, int kw, const int32_t* kidx,
                                const float* sigma, const float* pool, int64_t nPool, const int32_t* nidx,
                                int factor, int pad_mode, int down_mode, int noise_mode, float* lr,
                                void* workspace, int64_t workspace_bytes, int algo, void* stream) {
    const int64_t need = kmsr_degrade_workspace_bytes(nK, C, kh, kw, factor, down_mode);
    if (need < 0) return (int)need;
    KMSR_REQUIRE(workspace && workspace_bytes >= need, KMSR_E_INVALID,
                 "degrade_batch: workspace of %lld B needed, %lld given", (long long)need,
                 (long long)workspace_bytes);
    Geometry g;
    make_geometry(0, 0, kh, kw, factor, down_mode, &g);
    // carve: [dsum (256-aligned) | comp (256-aligned)]
    uintptr_t base = ((uintptr_t)workspace + 255) & ~(uintptr_t)255;
    float* dsum = (float*)base;
    float* comp = (float*)(base + (((size_t)nK * C * sizeof(float)) + 255) / 256 * 256);
    int rc = kmsr_prepare_kernels(kbank, nK, C, kh, kw, factor, down_mode, comp, dsum, stream);
    if (rc != KMSR_OK) return rc;
    return kmsr_degrade_prepared(hr, N, C, H, W, hr_stride_n, hr_stride_c, hr_stride_h, patch_offsets, comp,
                                 dsum, nK, kh, kw, kidx, sigma, pool, nPool, nidx, factor, pad_mode,
                                 down_mode, noise_mode, lr, algo, stream);
}

KMSR_API int kmsr_add_noise(const float* blurred, int64_t N, int C, int64_t hw, const float* pool,
                            int64_t nPool, const int32_t* nidx, const float* sigma, const int32_t* kidx,
                            float* out, void* stream) {
    KMSR_REQUIRE(N >= 0 && C >= 1 && hw >= 0, KMSR_E_INVALID, "add_noise: N=%lld C=%d hw=%lld", (long long)N, C,
                 (long long)hw);
    if (N == 0 || hw == 0) return KMSR_OK;
    KMSR_REQUIRE(blurred && pool && nidx && out && nPool >= 1, KMSR_E_INVALID, "add_noise: null pointer / empty pool");
    return launch_add_noise(blurred, N, C, hw, pool, nidx, sigma, kidx, out, (cudaStream_t)stream);
}

KMSR_API int kmsr_crop_sub(const float* geo, const float* den, int C, int H, int W, const int32_t* top,
                           const int32_t* left, int64_t n_samples, int crop, float* pool, void* stream) {
    KMSR_REQUIRE(C >= 1 && H >= 1 && W >= 1 && crop >= 1 && n_samples >= 0, KMSR_E_INVALID,
                 "crop_sub: C=%d H=%d W=%d crop=%d n=%lld", C, H, W, crop, (long long)n_samples);
    KMSR_REQUIRE(H >= crop && W >= crop, KMSR_E_INVALID, "crop_sub: image %dx%d smaller than crop %d", H, W, crop);
    if (n_samples == 0) return KMSR_OK;
    KMSR_REQUIRE(geo && den && top && left && pool, KMSR_E_INVALID, "crop_sub: null pointer");
    return launch_crop_sub(geo, den, C, H, W, top, left, n_samples, crop, pool, (cudaStream_t)stream);
}

KMSR_API int kmsr_band_stats(const float* x, int64_t N, int C, int64_t hw, int64_t x_stride_n, double* mean,
                             double* std, double* sums, void* stream) {
    KMSR_REQUIRE(N >= 0 && C >= 1 && hw >= 1, KMSR_E_INVALID, "band_stats: N=%lld C=%d hw=%lld", (long long)N, C,
                 (long long)hw);
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(x && mean && std, KMSR_E_INVALID, "band_stats: null pointer");
    KMSR_REQUIRE(x_stride_n >= (int64_t)C * hw, KMSR_E_INVALID, "band_stats: patch stride %lld < C*hw",
                 (long long)x_stride_n);
    return launch_band_stats(x, N, C, hw, x_stride_n, mean, std, sums, (cudaStream_t)stream);
}

KMSR_API int kmsr_water_mask(float* data, int C, int64_t hw, int nir, float invalid, float tmin, float tmax,
                             float* masked, void* stream) {
    KMSR_REQUIRE(C >= 1 && hw >= 0 && nir >= 0 && nir < C, KMSR_E_INVALID, "water_mask: C=%d hw=%lld nir=%d", C,
                 (long long)hw, nir);
    if (hw == 0) return KMSR_OK;
    KMSR_REQUIRE(data && masked, KMSR_E_INVALID, "water_mask: null pointer");
    return launch_water_mask(data, C, hw, nir, invalid, tmin, tmax, masked, (cudaStream_t)stream);
}

KMSR_API int64_t kmsr_keep_mask_workspace_bytes(int H, int W, int P, int stride) {
    return keep_mask_workspace(H, W, P, stride);
}

KMSR_API int kmsr_scene_keep_mask(const float* data, int C, int H, int W, int nir, float invalid, float tmin,
                                  float tmax, int P, int stride, double nan_threshold, uint8_t* keep,
                                  int32_t* nan_count, void* workspace, int64_t workspace_bytes, void* stream) {
    KMSR_REQUIRE(C >= 1 && H >= 0 && W >= 0 && P >= 1 && stride >= 1 && nir >= 0 && nir < C, KMSR_E_INVALID,
                 "scene_keep_mask: C=%d H=%d W=%d P=%d stride=%d nir=%d", C, H, W, P, stride, nir);
    if (H < P || W < P) return KMSR_OK;
    KMSR_REQUIRE(data && keep, KMSR_E_INVALID, "scene_keep_mask: null pointer");
    return launch_scene_keep_mask(data, C, H, W, nir, invalid, tmin, tmax, P, stride, nan_threshold, keep, nan_count,
                                  workspace, workspace_bytes, (cudaStream_t)stream);
}

KMSR_API int kmsr_keep_mask(const float* masked, int C, int H, int W, int P, int stride, double nan_threshold,
                            uint8_t* keep, int32_t* nan_count, void* workspace, int64_t workspace_bytes,
                            void* stream) {
    KMSR_REQUIRE(C >= 1 && H >= 0 && W >= 0 && P >= 1 && stride >= 1, KMSR_E_INVALID,
                 "keep_mask: C=%d H=%d W=%d P=%d stride=%d", C, H, W, P, stride);
    if (H < P || W < P) return KMSR_OK;
    KMSR_REQUIRE(masked && keep, KMSR_E_INVALID, "keep_mask: null pointer");
    return launch_keep_mask(masked, C, H, W, P, stride, nan_threshold, keep, nan_count, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

KMSR_API int64_t kmsr_denoise_workspace_bytes(int64_t N, int C, int H, int W) {
    if (N < 0 || C < 1 || H < 2 || W < 2) {
        set_error("denoise_workspace_bytes: N=%lld C=%d H=%d W=%d", (long long)N, C, H, W);
        return KMSR_E_INVALID;
    }
    return denoise_workspace(N, C, H, W);
}

KMSR_API int kmsr_estimate_sigma(const float* x, int64_t N, int C, int H, int W, int64_t x_stride_n, double* sigma,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    KMSR_REQUIRE(N >= 0 && C >= 1 && H >= 2 && W >= 2, KMSR_E_INVALID, "estimate_sigma: N=%lld C=%d H=%d W=%d",
                 (long long)N, C, H, W);
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(x && sigma && workspace, KMSR_E_INVALID, "estimate_sigma: null pointer");
    KMSR_REQUIRE(x_stride_n >= (int64_t)C * H * W, KMSR_E_INVALID, "estimate_sigma: patch stride %lld < C*H*W",
                 (long long)x_stride_n);
    return launch_estimate_sigma(x, N, C, H, W, x_stride_n, sigma, workspace, workspace_bytes, (cudaStream_t)stream);
}

KMSR_API int kmsr_denoise_nlm(const float* x, int64_t N, int C, int H, int W, int64_t x_stride_n, double h_factor,
                              int patch_size, int patch_distance, float* out, double* sigma, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    KMSR_REQUIRE(N >= 0 && C >= 1 && H >= 2 && W >= 2, KMSR_E_INVALID, "denoise_nlm: N=%lld C=%d H=%d W=%d", (long long)N,
                 C, H, W);
    const char* why = "";
    KMSR_REQUIRE(nlm_shape_ok(patch_size, patch_distance, &why), KMSR_E_UNSUPPORTED,
                 "denoise_nlm: patch_size=%d patch_distance=%d not covered: %s", patch_size, patch_distance, why);
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(x && out && sigma && workspace, KMSR_E_INVALID, "denoise_nlm: null pointer");
    KMSR_REQUIRE(x != out, KMSR_E_INVALID, "denoise_nlm: out must not alias x (every tile reads its neighbours' pixels)");
    KMSR_REQUIRE(x_stride_n >= (int64_t)C * H * W, KMSR_E_INVALID, "denoise_nlm: patch stride %lld < C*H*W",
                 (long long)x_stride_n);
    int rc = launch_estimate_sigma(x, N, C, H, W, x_stride_n, sigma, workspace, workspace_bytes, (cudaStream_t)stream);
    if (rc != KMSR_OK) return rc;
    return launch_nlm(x, N, C, H, W, x_stride_n, reinterpret_cast<const double*>(workspace), sigma, h_factor, patch_distance,
                      out, (cudaStream_t)stream);
}

KMSR_API int64_t kmsr_selector_weight_floats(int cin, int cout) {
    if (cin < 1 || cout < 32 || cout % 32 != 0) {
        set_error("selector_weight_floats: cin=%d cout=%d", cin, cout);
        return KMSR_E_INVALID;
    }
    return selector_wsplit_floats(cin, cout);
}

KMSR_API int64_t kmsr_selector_workspace_bytes(int64_t N, int H, int W) {
    if (N < 0 || H < 1 || W < 1) {
        set_error("selector_workspace_bytes: N=%lld H=%d W=%d", (long long)N, H, W);
        return KMSR_E_INVALID;
    }
    return selector_workspace(N, H, W);
}

KMSR_API int kmsr_selector_logits(const float* x, int64_t N, int H, int W, const float* w1, const float* b1, const float* w2,
                                  const float* b2, const float* w3, const float* b3, const float* fc_w, const float* fc_b,
                                  float* logits, void* workspace, int64_t workspace_bytes, void* stream) {
    KMSR_REQUIRE(N >= 0 && H >= 1 && W >= 1, KMSR_E_INVALID, "selector_logits: N=%lld H=%d W=%d", (long long)N, H, W);
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && fc_w && fc_b && logits && workspace, KMSR_E_INVALID,
                 "selector_logits: null pointer");
    return launch_selector(x, N, H, W, w1, b1, w2, b2, w3, b3, fc_w, fc_b, logits, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

KMSR_API int kmsr_selector_umma_supported(int H, int W) { return selector_umma_shape_ok(H, W) ? 1 : 0; }

KMSR_API int64_t kmsr_selector_umma_weight_floats(int cin, int cout) {
    if (!((cin == 5 || (cin >= 16 && cin % 16 == 0)) && cout >= 16 && cout <= 128 && cout % 16 == 0)) {
        set_error("selector_umma_weight_floats: cin=%d cout=%d", cin, cout);
        return KMSR_E_INVALID;
    }
    return selector_umma_wfloats(cin, cout);
}

KMSR_API int64_t kmsr_selector_umma_workspace_bytes(int64_t N, int H, int W) {
    if (N < 0 || !selector_umma_shape_ok(H, W)) {
        set_error("selector_umma_workspace_bytes: N=%lld H=%d W=%d", (long long)N, H, W);
        return KMSR_E_INVALID;
    }
    return selector_umma_workspace(N, H, W);
}

KMSR_API int kmsr_selector_logits_umma(const float* x, int64_t N, int H, int W, const float* w1, const float* b1, const float* w2,
                                       const float* b2, const float* w3, const float* b3, const float* fc_w, const float* fc_b,
                                       float* logits, void* workspace, int64_t workspace_bytes, void* stream) {
    KMSR_REQUIRE(N >= 0 && H >= 1 && W >= 1, KMSR_E_INVALID, "selector_logits_umma: N=%lld H=%d W=%d", (long long)N, H, W);
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(x && w1 && b1 && w2 && b2 && w3 && b3 && fc_w && fc_b && logits && workspace, KMSR_E_INVALID,
                 "selector_logits_umma: null pointer");
    return launch_selector_umma(x, N, H, W, w1, b1, w2, b2, w3, b3, fc_w, fc_b, logits, workspace, workspace_bytes,
                                (cudaStream_t)stream);
}

}  // extern "C"
