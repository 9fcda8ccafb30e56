/*
 * kmsr.h -- C ABI of libkmsr.so: the B200 (sm_100a) LR/HR pair-synthesis hot path.
 *
 * The reference (Zhiyyeah/Kernel-Modeling-Super-Resolution) has no FFI: its boundary for this
 * path is a set of Python functions that call torch/numpy on the CPU.  Each entry point below
 * replaces the arithmetic of one of those call sites; the Python drop-ins in
 * kernel-modeling-super-resolution_b200/ keep the reference signatures and bind these symbols
 * with ctypes (INTEGRATION.md shows the stub).  Paths are relative to
 * /root/reference/kernel_from_lr_gan/.
 *
 * Conventions
 *   - every pointer named below is a DEVICE pointer unless it says "host";
 *   - the caller (PyTorch's allocator) owns every buffer, outputs and workspaces included:
 *     the library never allocates or frees device memory and keeps no pointer after a call;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *     stream-ordered and asynchronous with respect to the host;
 *   - return value 0 = success, negative = KMSR_E_*; the message of the last failure on the
 *     calling thread is available from kmsr_last_error().  Nothing throws, nothing exits;
 *   - float32 CHW images, band order 443/490/555/660/865 nm (C_30apply_kernel_to_landsat.py:49).
 */
#ifndef KMSR_H_
#define KMSR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define KMSR_API
#else
#define KMSR_API __attribute__((visibility("default")))
#endif

#define KMSR_VERSION 100 /* 0.1.0 */

enum {
    KMSR_OK = 0,
    KMSR_E_INVALID = -1,     /* bad argument (shape, mode, null pointer) */
    KMSR_E_UNSUPPORTED = -2, /* valid request this build has no kernel for */
    KMSR_E_CUDA = -3,        /* CUDA runtime / driver error (message has the cudaError string) */
    KMSR_E_ALIGN = -4        /* pointer or stride not aligned as the chosen kernel requires */
};

/* pad_mode: how the blur sees pixels outside the patch */
enum {
    KMSR_PAD_REPLICATE = 0, /* F.pad(mode='replicate')  C_30:107-109, C_31:85-87 */
    KMSR_PAD_ZERO = 1       /* F.conv2d(padding=k//2)   muti_kernel/train_gemini.py:128 */
};
/* down_mode: how the blurred image is reduced */
enum {
    KMSR_DOWN_BOXMEAN = 0, /* int(log2(f)) cascaded avg_pool2d(2,2)  C_30:120-122, C_31:92-95 */
    KMSR_DOWN_DECIMATE = 1 /* out[:, :, ::f, ::f]                     muti_kernel/train_gemini.py:134 */
};
/* noise_mode: what is added to the degraded patch */
enum {
    KMSR_NOISE_NONE = 0,  /* C_30 / C_31 output ('blurred' / 'lr' group) */
    KMSR_NOISE_ADD = 1,   /* blurred + noise_pool[idx]                E_make_train_data.py:72-74 */
    KMSR_NOISE_SIGMA = 2  /* blurred + sigma[kidx,c] * noise          muti_kernel/train_gemini.py:137 */
};
/* algo: kernel selection for kmsr_degrade_* */
enum {
    KMSR_ALGO_AUTO = 0,    /* the fastest kernel that covers the call: TMA (headline shape), BOX (factor 2 / 4 and
                              patches of at most 64 x 64), STREAM (the other box-mean calls with k in 11..31 at factor 8),
                              REG (what BOX refuses), TILED (everything else); kmsr_last_algo() says which ran          */
    KMSR_ALGO_TILED = 1,   /* polyphase shared-memory tile kernel (any k, factor, H, W, pad / down mode)   */
    KMSR_ALGO_TMA = 2,     /* headline TMA row-streaming kernel (k = 13, factor 8, W a multiple of 256, H % 8 == 0,
                              H <= 512); KMSR_E_UNSUPPORTED if the shape does not qualify                      */
    KMSR_ALGO_STREAM = 3,  /* generic TMA row-streaming kernel (k in 11/13/15/21/31, factor 2/4/8, box mean,
                              W in 64/128/256*m); KMSR_E_UNSUPPORTED otherwise                             */
    KMSR_ALGO_REG = 4,     /* register-tile stencil kernel for the FP32-bound shapes (k in 11/13/15/21/31,
                              factor 2/4, box mean, any H / W / strides); KMSR_E_UNSUPPORTED otherwise       */
    KMSR_ALGO_BOX = 5      /* TMA box-tile register stencil (k in 11/13/15/21/31, factor 2/4/8, box mean, H % 8 == 0,
                              W % 16 == 0, 16-byte aligned strides): the factor-2 / factor-4 sweep shapes and 64-wide
                              patches; KMSR_E_UNSUPPORTED otherwise                                          */
};

/* ---- library ------------------------------------------------------------------------------- */
KMSR_API int kmsr_version(void);
KMSR_API const char* kmsr_last_error(void);
/* Device facts used to size grids / report rooflines (cudaGetDeviceProperties). */
KMSR_API int kmsr_device_info(int device, int* sm_count, int* cc_major, int* cc_minor,
                              int64_t* l2_bytes, int64_t* smem_optin_bytes);
/* Measurement aid (bench.py roofline.fp32_peak_tflops): one launch of dependent packed-FFMA2 chains on every SM of
 * the current device, `iters` rounds; *fma_count (host) receives the fp32 multiply-adds the launch performs.  The
 * caller times it with CUDA events on `stream`.  `sink` is a device buffer of >= 512 * sm_count floats (never
 * written in practice).  Not part of the reference path.                                                        */
KMSR_API int kmsr_fp32_probe(float* sink, int iters, double* fma_count, void* stream);

/* Opt-in range check for DEVICE-resident int32 indices (kidx < nK, nidx < nPool, crop offsets): the degrade / gather
 * kernels trust their indices (an index >= nK / nPool is an out-of-bounds device read).  Host-side callers check host
 * arrays for free (ops.py does); for arrays that only exist on the device this entry counts the entries outside
 * [0, upper) on `stream`, SYNCHRONISES it, and returns KMSR_E_INVALID (message names `what` and the count) if any.
 * `scratch` is one caller-owned device int32.  Mirrors the reference's IndexError on noise_pool[idx]
 * (E_make_train_data.py:72-74) and its crop contract (D_build_noise_pool.py:44-45).                              */
KMSR_API int kmsr_validate_indices(const int32_t* idx, int64_t n, int64_t upper, int32_t* scratch,
                                   const char* what, void* stream);

/* ---- a2/a3: blur + downsample (+ noise) ------------------------------------------------------
 * Replaces apply_kernel_degradation (C_30:68-124 == C_31:59-97) for a batch of patches, the
 * multi-kernel + sigma composition of SURVEY.md 8a row 3 (train_gemini.py:107-138) and the
 * fused E.add_noise (E_make_train_data.py:65-74).
 *
 * Shape helpers (host only, no device work):
 *   effective factor f' = 2^floor(log2(factor))      (C_30:121  int(np.log2(f)))
 *   blurred size  Hb = H + 2*(kh/2) - kh + 1         (C_30:107-117; H for odd kh, H+1 for even)
 *   BOXMEAN : Ho = Hb / f' (floor, equals the cascaded floors), composite window KH = kh + f' - 1,
 *             stride f'
 *   DECIMATE: Ho = ceil(Hb / factor), composite window = kernel, stride = factor
 */
KMSR_API int kmsr_degrade_out_size(int H, int W, int kh, int kw, int factor, int down_mode,
                                   int* Ho, int* Wo);
KMSR_API int kmsr_composite_size(int kh, int kw, int factor, int down_mode,
                                 int* KH, int* KW, int* stride);
/* bytes of the `workspace` kmsr_degrade_batch needs (composite bank + per-band sum residuals) */
KMSR_API int64_t kmsr_degrade_workspace_bytes(int64_t nK, int C, int kh, int kw, int factor,
                                              int down_mode);

/* Normalise every band of every kernel as C_30:93-97 does (divide by the band sum iff it is > 0),
 * fold the box mean into it (composite stride-f' kernel) and record
 * dsum[k,c] = sum(normalised band) - 1.
 *   kbank [nK, C, kh, kw]  ->  comp [nK, C, KH, KW],  dsum [nK, C]                              */
KMSR_API int kmsr_prepare_kernels(const float* kbank, int64_t nK, int C, int kh, int kw,
                                  int factor, int down_mode, float* comp, float* dsum,
                                  void* stream);

/* lr[n,c] = degrade(hr[n], K[kidx[n]])[c]  (+ scale[n,c] * pool[nidx[n], c])
 *   hr            patch n band c row y starts at hr + off(n) + c*hr_stride_c + y*hr_stride_h
 *                 (element strides; rows contiguous); off(n) = patch_offsets ? patch_offsets[n]
 *                 : n*hr_stride_n.  patch_offsets lets patches be overlapping views of a scene
 *                 (A_00_patch_cutter_universal.py:176).
 *   comp, dsum    from kmsr_prepare_kernels (same kh, kw, factor, down_mode)
 *   kidx  [N]     kernel per patch, NULL = kernel 0 for all (C_30 / C_31 single-kernel apply)
 *   sigma [nK,C]  only for KMSR_NOISE_SIGMA
 *   pool  [nPool, C, Ho, Wo], nidx [N]   only when noise_mode != NONE
 *   lr    [N, C, Ho, Wo] contiguous
 */
KMSR_API int kmsr_degrade_prepared(const float* hr, int64_t N, int C, int H, int W,
                                   int64_t hr_stride_n, int64_t hr_stride_c, int64_t hr_stride_h,
                                   const int64_t* patch_offsets,
                                   const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                                   const int32_t* kidx,
                                   const float* sigma, const float* pool, int64_t nPool,
                                   const int32_t* nidx,
                                   int factor, int pad_mode, int down_mode, int noise_mode,
                                   float* lr, int algo, void* stream);

/* Scene-scale form (BASELINE config 4): the N patches are H x W WINDOWS of one resident scene
 * [C, scene_h, scene_w] (band stride hr_stride_c, row stride hr_stride_h), window n starting at element
 * offset patch_offsets[n] of band 0 -- the windows A_00_patch_cutter_universal.py:166-176 cuts, never
 * materialised.  Replicate padding clamps to the WINDOW (patches are cut first, then blurred).  Knowing
 * the scene extents lets the TMA kernel stream windows straight from the scene (overlapping windows are
 * served by L2); `x_multiple` is the caller's promise that every window's left column is a multiple of
 * it (the TMA kernel needs a multiple of 4; pass 1 if unknown and the tiled kernel runs).            */
KMSR_API int kmsr_degrade_windows(const float* scene, int C, int scene_h, int scene_w,
                                  int64_t hr_stride_c, int64_t hr_stride_h,
                                  const int64_t* patch_offsets, int64_t N, int H, int W, int x_multiple,
                                  const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                                  const int32_t* kidx,
                                  const float* sigma, const float* pool, int64_t nPool,
                                  const int32_t* nidx,
                                  int factor, int pad_mode, int down_mode, int noise_mode,
                                  float* lr, int algo, void* stream);

/* Pair generation AND per-band statistics of the HR patches in one pass (BASELINE config 3:
 * E_make_train_data.py:223-250 + data_mean_std.py:32-33): same as kmsr_degrade_prepared on contiguous
 * [N, C, H, W] patches (patch stride hr_stride_n), plus mean / std [N, C] float64 and the accumulated
 * `sums` [2C+1] of kmsr_band_stats.  When the TMA kernel takes the call the sums of x and x^2 are
 * gathered from the pixels while they sit in registers (HR crosses HBM once); bands that contain NaN are
 * redone by the exact NaN-skipping two-pass kernel.  Otherwise degrade and band_stats run back to back.
 * workspace: kmsr_degrade_stats_workspace_bytes(N, C).                                              */
KMSR_API int64_t kmsr_degrade_stats_workspace_bytes(int64_t N, int C);
KMSR_API int kmsr_degrade_stats_prepared(const float* hr, int64_t N, int C, int H, int W,
                                         int64_t hr_stride_n,
                                         const float* comp, const float* dsum, int64_t nK, int kh, int kw,
                                         const int32_t* kidx,
                                         const float* sigma, const float* pool, int64_t nPool,
                                         const int32_t* nidx,
                                         int factor, int pad_mode, int down_mode, int noise_mode,
                                         float* lr, double* mean, double* std, double* sums,
                                         void* workspace, int64_t workspace_bytes, int algo, void* stream);

/* kmsr_prepare_kernels into `workspace`, then kmsr_degrade_prepared, on the same stream. */
KMSR_API int kmsr_degrade_batch(const float* hr, int64_t N, int C, int H, int W,
                                int64_t hr_stride_n, int64_t hr_stride_c, int64_t hr_stride_h,
                                const int64_t* patch_offsets,
                                const float* kbank, int64_t nK, int kh, int kw,
                                const int32_t* kidx,
                                const float* sigma, const float* pool, int64_t nPool,
                                const int32_t* nidx,
                                int factor, int pad_mode, int down_mode, int noise_mode,
                                float* lr, void* workspace, int64_t workspace_bytes,
                                int algo, void* stream);

/* ---- a5: noise-pool gather + add (E_make_train_data.py:65-74) -------------------------------
 * out[n] = blurred[n] + scale * pool[nidx[n]],  scale = 1 (sigma NULL) or sigma[kidx[n], c].
 * blurred/out [N, C, hw], pool [nPool, C, hw]; indices are drawn on the host (bit-exact MT19937). */
KMSR_API int kmsr_add_noise(const float* blurred, int64_t N, int C, int64_t hw,
                            const float* pool, int64_t nPool, const int32_t* nidx,
                            const float* sigma, const int32_t* kidx, float* out, void* stream);

/* ---- a4: noise-pool construction (D_build_noise_pool.py:88 + :41-53) ------------------------
 * pool[m, c, y, x] = geo[c, top[m]+y, left[m]+x] - den[c, top[m]+y, left[m]+x]  for one file;
 * top/left [n_samples] are host-drawn (CPython random.randint, inclusive bounds).             */
KMSR_API int kmsr_crop_sub(const float* geo, const float* den, int C, int H, int W,
                           const int32_t* top, const int32_t* left, int64_t n_samples, int crop,
                           float* pool, void* stream);

/* ---- a7: per-band statistics (data_mean_std.py:32-33, 45-46) --------------------------------
 * For every patch n and band c: NaN-skipping mean and population std over hw pixels.
 *   x [N, C, hw] (patch stride x_stride_n elements, bands contiguous)
 *   mean, std [N, C] float64 (per patch, what np.nanmean / np.nanstd return before S:41-46)
 *   sums [2*C + 1] float64, ACCUMULATED (+=): sum over patches of mean[c], of std[c], and N --
 *        the 88-byte vector the multi-GPU path all-reduces.  May be NULL.        */
KMSR_API int kmsr_band_stats(const float* x, int64_t N, int C, int64_t hw, int64_t x_stride_n,
                             double* mean, double* std, double* sums, void* stream);

/* ---- a8: scene mask + patch keep-mask (A_00_patch_cutter_universal.py:89-123, 152-183) ------
 * kmsr_water_mask: data [C, hw] is updated IN PLACE (invalid -> NaN, CUT:102) and masked [C, hw]
 * receives NaN wherever NIR (band nir) is outside [tmin, tmax] or NaN (CUT:108-113).
 * `masked` may alias `data`.                                                                   */
KMSR_API int kmsr_water_mask(float* data, int C, int64_t hw, int nir, float invalid,
                             float tmin, float tmax, float* masked, void* stream);
/* keep[i, j] = 1 iff the P x P window at (i*stride, j*stride) over all C bands holds no more than
 * nan_threshold * C*P*P NaNs (CUT:179-183; 0.0 = none).  nan_count [hp, wp] int32 optional.
 * workspace: kmsr_keep_mask_workspace_bytes(H, W, P, stride).                                   */
KMSR_API int64_t kmsr_keep_mask_workspace_bytes(int H, int W, int P, int stride);
KMSR_API int kmsr_keep_mask(const float* masked, int C, int H, int W, int P, int stride,
                            double nan_threshold, uint8_t* keep, int32_t* nan_count,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* Keep grid of the masked scene computed straight from the RAW scene: nothing is written, the scene is read once.
 * Same result as kmsr_water_mask followed by kmsr_keep_mask.  With nan_threshold = 0 (the reference's value,
 * CUT:31) a kept window contains no masked pixel, so its pixels ARE the raw scene's and kmsr_degrade_windows can
 * run on the raw scene: the masked copy of A_00_patch_cutter_universal.py:112-113 is never materialised.
 * Needs P % stride == 0 (KMSR_E_UNSUPPORTED otherwise); workspace: kmsr_keep_mask_workspace_bytes(H, W, P, stride). */
KMSR_API int kmsr_scene_keep_mask(const float* data, int C, int H, int W, int nir, float invalid,
                                  float tmin, float tmax, int P, int stride, double nan_threshold,
                                  uint8_t* keep, int32_t* nan_count,
                                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- f4: upstream denoise stage (denoise/denoise.py:34-65 denoise_band_float_nlm) -----------------
 * Produces the `denoised` group that D_build_noise_pool.py:85 and E_make_train_data.py:234 read.  Per band:
 * NaN pixels are filled with the band's nanmean (:43-44), sigma = skimage estimate_sigma (:47: MAD of the db2
 * 'dd' coefficients of pywt.dwtn / 0.6745), out = skimage denoise_nl_means(fast_mode=True, patch_size,
 * patch_distance, h = h_factor * sigma, sigma = sigma) (:56-63), NaN pixels restored (:66).  An all-NaN band
 * is returned unchanged with sigma 0.0 (:40-41).  skimage / PyWavelets are third-party and absent from the
 * build image: their published algorithms are restated ("parity unpinned").
 *   x, out [N, C, H, W] float32 (patch stride x_stride_n elements; out contiguous, must not alias x)
 *   sigma  [N, C] float64  (what the reference returns next to the band and stores as `<band>_sigma`, :248)
 *   patch_size 7 (6 is rounded up to 7 as skimage does), patch_distance 0..11: KMSR_E_UNSUPPORTED otherwise
 *   workspace: kmsr_denoise_workspace_bytes(N, C, H, W), 256-byte aligned                                */
KMSR_API int64_t kmsr_denoise_workspace_bytes(int64_t N, int C, int H, int W);
KMSR_API int kmsr_estimate_sigma(const float* x, int64_t N, int C, int H, int W, int64_t x_stride_n,
                                 double* sigma, void* workspace, int64_t workspace_bytes, void* stream);
KMSR_API int kmsr_denoise_nlm(const float* x, int64_t N, int C, int H, int W, int64_t x_stride_n,
                              double h_factor, int patch_size, int patch_distance, float* out, double* sigma,
                              void* workspace, int64_t workspace_bytes, void* stream);

/* ---- f2: learned kernel pick (SelectorNet, muti_kernel/train_gemini.py:14-39) ----------------------
 * logits [N, 10] of the selector CNN for patches x [N, 5, H, W] (contiguous): three 3x3 / stride-2 / pad-1
 * convolutions (5 -> 32 -> 64 -> 128) with eval-mode BatchNorm folded in and ReLU, global average pooling, a
 * 128 -> 10 linear layer.  argmax over the 10 logits is the kernel index (kidx) of kmsr_degrade_*.
 * The convolutions run as 3xTF32-split tensor-core MMAs (fp32-level accuracy).  Weight blobs w1 / w2 / w3 are
 * prepared on the host (kmsr_b200/selector.py: BN fold, fragment-major [chunk][channel block][tap][quad][lane][4] layout,
 * TF32 hi / lo split), kmsr_selector_weight_floats(cin, cout) floats each, 16-byte aligned; b1 / b2 / b3 are
 * the folded biases [32] / [64] / [128]; fc_w [10, 128], fc_b [10].  N <= 65535 per call.
 * workspace: kmsr_selector_workspace_bytes(N, H, W), 256-byte aligned (the two intermediate activations). */
KMSR_API int64_t kmsr_selector_weight_floats(int cin, int cout);
KMSR_API int64_t kmsr_selector_workspace_bytes(int64_t N, int H, int W);
KMSR_API int kmsr_selector_logits(const float* x, int64_t N, int H, int W,
                                  const float* w1, const float* b1, const float* w2, const float* b2,
                                  const float* w3, const float* b3, const float* fc_w, const float* fc_b,
                                  float* logits, void* workspace, int64_t workspace_bytes, void* stream);

/* The same logits through the 5th-generation tensor cores (csrc/selector_umma.cuh: tcgen05.mma kind::tf32, accumulators
 * and the A operand in tensor memory, the activations of layers 2 / 3 fetched as strided TMA boxes, 3xTF32 split).
 * Shapes: kmsr_selector_umma_supported(H, W) != 0 (H, W multiples of 8, W in {64, 128, 256}, an even number of 128-pixel
 * tiles per layer; 256 x 256 and 128 x 128 patches qualify).  Weight blobs are the per-stage shared-memory images
 * [stage][4][2 cout / 8][8][4] (rows < cout: TF32 hi part, rows >= cout: lo part; kmsr_b200/selector.py builds them),
 * kmsr_selector_umma_weight_floats(cin, cout) floats each, 16-byte aligned; x 16-byte aligned; biases / fc as above.
 * workspace: kmsr_selector_umma_workspace_bytes(N, H, W), 256-byte aligned.  Replaces the same reference forward as
 * kmsr_selector_logits (train_gemini.py:35-39); other shapes return KMSR_E_UNSUPPORTED. */
KMSR_API int kmsr_selector_umma_supported(int H, int W);
KMSR_API int64_t kmsr_selector_umma_weight_floats(int cin, int cout);
KMSR_API int64_t kmsr_selector_umma_workspace_bytes(int64_t N, int H, int W);
KMSR_API int kmsr_selector_logits_umma(const float* x, int64_t N, int H, int W,
                                       const float* w1, const float* b1, const float* w2, const float* b2,
                                       const float* w3, const float* b3, const float* fc_w, const float* fc_b,
                                       float* logits, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------
 * Launch counter: number of kernels this library launched on the calling process since load.   */
KMSR_API int64_t kmsr_launch_count(void);
/* Name of the kernel the last kmsr_degrade_* call on this thread selected ("tiled" | "tma" | "stream"). */
KMSR_API const char* kmsr_last_algo(void);

#ifdef __cplusplus
}
#endif
#endif /* KMSR_H_ */
