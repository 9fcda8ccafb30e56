/*
 * oracle_nlm.c -- library-free CPU restatement of the upstream denoise stage:
 * denoise/denoise.py:34-65 `denoise_band_float_nlm` = skimage.restoration.estimate_sigma +
 * skimage.restoration.denoise_nl_means(fast_mode=True, patch_size=7, patch_distance=11).
 *
 * TEST INFRASTRUCTURE ONLY: loaded solely by tests/, __graft_entry__.smoke() and the cpu-baseline legs of
 * bench.py / tools/bench_denoise.py.  The product path never touches it.
 *
 * PARITY UNPINNED.  The arithmetic of this stage lives in two third-party packages the reference neither
 * vendors nor pins (no requirements file in the tree) and that are ABSENT from this image (no network):
 * scikit-image (skimage/restoration/_nl_means_denoising_cy.pyx `_fast_nl_means_denoising_2d`,
 * skimage/restoration/_denoise.py `estimate_sigma` / `_sigma_est_dwt`; API of 0.19 and later, where
 * `channel_axis=None` makes a 2-D image a single-channel one) and PyWavelets (`pywt.dwtn(image, 'db2')`,
 * default mode 'symmetric').  What follows restates their PUBLISHED algorithms -- Darbon et al. 2008 / Froment
 * 2015 integral-image non-local means with the 2*sigma^2 bias correction of Buades et al., and Donoho &
 * Johnstone's MAD estimator on the finest diagonal db2 detail coefficients -- anchored on the reference's call
 * site (denoise.py:47, :56-63).  There are no golden vectors of the reference for this stage and the real
 * libraries cannot be executed here, so nothing below is pinned against them.
 *
 * Two evaluations are provided, as for the degrade path:
 *   orc_nlm_fast_f32   the integral-image algorithm in float32, loop for loop (what skimage's fused-type
 *                      kernel does for a float32 image): integral of (squared difference - 2 var) per shift,
 *                      patch distance from four corners, symmetric scatter of the weight to both pixels;
 *   orc_nlm_exact_f64  the value of the same formula in float64 for every output pixel, plus the
 *                      sensitivity of each pixel to the hard cut-off `distance > 5 -> skip` (a weight of
 *                      e^-5 appears or disappears when rounding moves a distance across the threshold).
 *
 * For an output pixel p (all of them lie >= pad = offset + d + 1 inside the reflect-padded image P) the
 * scatter form reduces to
 *   D_t(p)   = sum_{a,b = -offset+1 .. offset} (P(p+(a,b)) - P(p+(a,b)+t))^2  -  (2*offset)^2 * 2 var
 *              (an integral image differenced at +-offset covers 2*offset = 6 rows, not 7)
 *   dist     = max(D_t(p), 0) / (h^2 s^2),   w_t(p) = dist > 5 ? 0 : exp(-dist),   w_0 = 2
 *   out(p)   = sum_t w_t(p) P(p+t) / sum_t w_t(p),     t in [-d, d]^2
 * (shifts with t_col < 0 arrive as the mirrored write of the pair (p+t, -t); the 0.5 + 0.5 of the
 * t_col = 0 column adds up to 1; the zero shift is written twice with weight 1).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* numpy.pad(mode='reflect'): mirror without repeating the edge sample; period 2(n-1) */
static int reflecti(int i, int n) {
    if (n == 1) return 0;
    int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}

/* half-sample symmetric extension of PyWavelets' mode 'symmetric': x[-1] = x[0], x[n] = x[n-1] */
static int symi(int i, int n) {
    int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

/* db2 decomposition high-pass filter (pywt.Wavelet('db2').dec_hi) */
static const double kDecHi[4] = {-0.4829629131445341, 0.8365163037378079, -0.2241438680420134, -0.12940952255126037};

static int orc_cmp_double_(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

static int cmp_float(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

/* ---- skimage estimate_sigma(image) for a 2-D float32 image (denoise.py:47):
 * coeffs = pywt.dwtn(image, 'db2') -> 'dd' = high-pass along axis 0, then along axis 1, each
 * out[o] = sum_j dec_hi[j] * ext[2o + 1 - j], o < (n + 3) / 2, float32 like pywt's float32 path;
 * _sigma_est_dwt: drop exact zeros, sigma = median(|dd|) / norm.ppf(0.75).
 * dd_abs (optional) receives |dd| [(H+3)/2, (W+3)/2].  Returns NaN when every coefficient is zero
 * (np.median of an empty array). */
ORC_API double orc_estimate_sigma(const float* img, int H, int W, float* dd_abs) {
    const int Ho = (H + 3) / 2, Wo = (W + 3) / 2;
    float f[4];
    for (int j = 0; j < 4; ++j) f[j] = (float)kDecHi[j];
    float* d0 = (float*)malloc(sizeof(float) * (size_t)Ho * W);
    float* dd = (float*)malloc(sizeof(float) * (size_t)Ho * Wo);
    if (!d0 || !dd) { free(d0); free(dd); return -1.0; }
    for (int o = 0; o < Ho; ++o)
        for (int x = 0; x < W; ++x) {
            float s = 0.0f;
            for (int j = 0; j < 4; ++j) s += f[j] * img[(size_t)symi(2 * o + 1 - j, H) * W + x];
            d0[(size_t)o * W + x] = s;
        }
    size_t n = 0;
    for (int y = 0; y < Ho; ++y)
        for (int o = 0; o < Wo; ++o) {
            float s = 0.0f;
            for (int j = 0; j < 4; ++j) s += f[j] * d0[(size_t)y * W + symi(2 * o + 1 - j, W)];
            float a = fabsf(s);
            if (dd_abs) dd_abs[(size_t)y * Wo + o] = a;
            if (a != 0.0f) dd[n++] = a;
        }
    double sigma;
    if (n == 0) {
        sigma = NAN;
    } else {
        qsort(dd, n, sizeof(float), cmp_float);
        /* np.median of float32: mean of the two middle values, in float32 */
        float med = (n & 1) ? dd[n / 2] : (dd[n / 2 - 1] + dd[n / 2]) / 2.0f;
        sigma = (double)med / 0.6744897501960817;      /* scipy.stats.norm.ppf(0.75) */
    }
    free(d0); free(dd);
    return sigma;
}

/* The same estimator with the transform evaluated in float64 (float32-rounded filter taps, as pywt uses for a
 * float32 image): the value the float32 evaluation above approximates. */
ORC_API double orc_estimate_sigma_f64(const float* img, int H, int W) {
    const int Ho = (H + 3) / 2, Wo = (W + 3) / 2;
    double f[4];
    for (int j = 0; j < 4; ++j) f[j] = (double)(float)kDecHi[j];
    double* d0 = (double*)malloc(sizeof(double) * (size_t)Ho * W);
    double* dd = (double*)malloc(sizeof(double) * (size_t)Ho * Wo);
    if (!d0 || !dd) { free(d0); free(dd); return -1.0; }
    for (int o = 0; o < Ho; ++o)
        for (int x = 0; x < W; ++x) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += f[j] * (double)img[(size_t)symi(2 * o + 1 - j, H) * W + x];
            d0[(size_t)o * W + x] = s;
        }
    size_t n = 0;
    for (int y = 0; y < Ho; ++y)
        for (int o = 0; o < Wo; ++o) {
            double s = 0.0;
            for (int j = 0; j < 4; ++j) s += f[j] * d0[(size_t)y * W + symi(2 * o + 1 - j, W)];
            if (s != 0.0) dd[n++] = fabs(s);
        }
    double sigma = NAN;
    if (n > 0) {
        qsort(dd, n, sizeof(double), orc_cmp_double_);
        const double med = (n & 1) ? dd[n / 2] : 0.5 * (dd[n / 2 - 1] + dd[n / 2]);
        sigma = med / 0.6744897501960817;
    }
    free(d0); free(dd);
    return sigma;
}

/* ---- skimage _fast_nl_means_denoising_2d for one float32 channel, loop for loop.
 * s = patch_size, d = patch_distance, h and var = sigma^2 as the wrapper passes them (cast to float32). */
ORC_API int orc_nlm_fast_f32(const float* img, int H, int W, int s, int d, float h, float var, float* out) {
    if (s % 2 == 0) s += 1;
    const int offset = s / 2;
    const int pad = offset + d + 1;
    const int nr = H + 2 * pad, nc = W + 2 * pad;
    const size_t np_ = (size_t)nr * nc;
    float* P = (float*)malloc(sizeof(float) * np_);
    float* res = (float*)calloc(np_, sizeof(float));
    float* wts = (float*)calloc(np_, sizeof(float));
    float* integ = (float*)malloc(sizeof(float) * np_);
    if (!P || !res || !wts || !integ) { free(P); free(res); free(wts); free(integ); return -1; }
    for (int r = 0; r < nr; ++r)
        for (int c = 0; c < nc; ++c) P[(size_t)r * nc + c] = img[(size_t)reflecti(r - pad, H) * W + reflecti(c - pad, W)];
    const float h2 = (float)pow((double)h, 2.0);
    const float s2 = (float)(s * s);
    const float h2s2 = 1.0f * h2 * s2;                   /* n_channels = 1 */
    const float var_diff = 2.0f * var;                   /* variance of the difference of two noisy pixels */
    for (int tr = -d; tr <= d; ++tr) {
        const int row_start = offset > offset - tr ? offset : offset - tr;
        const int row_end = nr - offset < nr - offset - tr ? nr - offset : nr - offset - tr;
        for (int tc = 0; tc <= d; ++tc) {
            const float alpha = (tc == 0 && tr != 0) ? 0.5f : 1.0f;
            memset(integ, 0, sizeof(float) * np_);
            {   /* _integral_image_2d */
                const int r0 = 1 > -tr ? 1 : -tr;
                const int r1 = nr < nr - tr ? nr : nr - tr;
                for (int r = r0; r < r1; ++r)
                    for (int c = 1; c < nc - tc; ++c) {
                        const float t = P[(size_t)r * nc + c] - P[(size_t)(r + tr) * nc + c + tc];
                        float dist = t * t;
                        dist -= var_diff;
                        integ[(size_t)r * nc + c] = dist + integ[(size_t)(r - 1) * nc + c] + integ[(size_t)r * nc + c - 1]
                                                    - integ[(size_t)(r - 1) * nc + c - 1];
                    }
            }
            for (int r = row_start; r < row_end; ++r) {
                const int rs = r + tr;
                for (int c = offset; c < nc - offset - tc; ++c) {
                    float dist = integ[(size_t)(r + offset) * nc + c + offset] + integ[(size_t)(r - offset) * nc + c - offset]
                                 - integ[(size_t)(r - offset) * nc + c + offset] - integ[(size_t)(r + offset) * nc + c - offset];
                    dist = (dist > 0.0f ? dist : 0.0f) / h2s2;
                    if (dist > 5.0f) continue;           /* DISTANCE_CUTOFF */
                    const int cs = c + tc;
                    const float w = (float)((double)alpha * exp(-(double)dist));
                    wts[(size_t)r * nc + c] += w;
                    wts[(size_t)rs * nc + cs] += w;
                    res[(size_t)r * nc + c] += w * P[(size_t)rs * nc + cs];
                    res[(size_t)rs * nc + cs] += w * P[(size_t)r * nc + c];
                }
            }
        }
    }
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const size_t q = (size_t)(r + pad) * nc + c + pad;
            out[(size_t)r * W + c] = res[q] / wts[q];
        }
    free(P); free(res); free(wts); free(integ);
    return 0;
}

/* ---- the same formula in float64, gather form (header).  h and var are the float32-rounded parameters.
 * flip (optional) [H, W]: upper bound of what the hard cut-off can change at this pixel when a distance is
 * perturbed by `eps` (absolute, in units of dist): sum over shifts with |dist - 5| < eps of
 * e^-5 |P(p+t) - out(p)| / (sum_t w - e^-5 * count). */
ORC_API int orc_nlm_exact_f64(const float* img, int H, int W, int s, int d, float h, float var, double eps,
                              double* out, double* flip) {
    if (s % 2 == 0) s += 1;
    const int offset = s / 2;
    const int pad = offset + d + 1;
    const int nr = H + 2 * pad, nc = W + 2 * pad;
    const size_t np_ = (size_t)nr * nc, no = (size_t)H * W;
    double* P = (double*)malloc(sizeof(double) * np_);
    double* integ = (double*)malloc(sizeof(double) * np_);
    double* sw = (double*)calloc(no, sizeof(double));
    double* sv = (double*)calloc(no, sizeof(double));
    /* shifts near the cut-off, kept per pixel as a running sum of |P(p+t)| candidates: two passes */
    double* amb_n = (double*)calloc(no, sizeof(double));
    if (!P || !integ || !sw || !sv || !amb_n) { free(P); free(integ); free(sw); free(sv); free(amb_n); return -1; }
    for (int r = 0; r < nr; ++r)
        for (int c = 0; c < nc; ++c) P[(size_t)r * nc + c] = (double)img[(size_t)reflecti(r - pad, H) * W + reflecti(c - pad, W)];
    const double h2s2 = (double)h * (double)h * (double)(s * s);
    const double var_diff = 2.0 * (double)var;
    const double npatch = (double)(2 * offset) * (double)(2 * offset);
    const double wcut = exp(-5.0);
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1 && !flip) break;
        if (pass == 1) memset(flip, 0, sizeof(double) * no);
        for (int tr = -d; tr <= d; ++tr)
            for (int tc = -d; tc <= d; ++tc) {
                /* integral image of the squared differences for this shift, over the rows / columns where both
                 * q and q + t are inside the padded image */
                memset(integ, 0, sizeof(double) * np_);
                const int r0 = 1 > 1 - tr ? 1 : 1 - tr, r1 = nr < nr - tr ? nr : nr - tr;
                const int c0 = 1 > 1 - tc ? 1 : 1 - tc, c1 = nc < nc - tc ? nc : nc - tc;
                for (int r = r0; r < r1; ++r)
                    for (int c = c0; c < c1; ++c) {
                        const double t = P[(size_t)r * nc + c] - P[(size_t)(r + tr) * nc + c + tc];
                        integ[(size_t)r * nc + c] = t * t + integ[(size_t)(r - 1) * nc + c] + integ[(size_t)r * nc + c - 1]
                                                    - integ[(size_t)(r - 1) * nc + c - 1];
                    }
                for (int y = 0; y < H; ++y)
                    for (int x = 0; x < W; ++x) {
                        const int r = y + pad, c = x + pad;
                        double D = integ[(size_t)(r + offset) * nc + c + offset] + integ[(size_t)(r - offset) * nc + c - offset]
                                   - integ[(size_t)(r - offset) * nc + c + offset] - integ[(size_t)(r + offset) * nc + c - offset];
                        D -= npatch * var_diff;
                        const double dist = (D > 0.0 ? D : 0.0) / h2s2;
                        const size_t o = (size_t)y * W + x;
                        const double pv = P[(size_t)(r + tr) * nc + c + tc];
                        if (pass == 0) {
                            if (!(dist > 5.0)) {
                                const double w = (tr == 0 && tc == 0) ? 2.0 * exp(-dist) : exp(-dist);
                                sw[o] += w;
                                sv[o] += w * pv;
                            }
                            if (fabs(dist - 5.0) < eps) amb_n[o] += 1.0;
                        } else if (fabs(dist - 5.0) < eps) {
                            const double den = sw[o] - wcut * amb_n[o];
                            flip[o] += wcut * fabs(pv - out[o]) / (den > 1.0 ? den : 1.0);
                        }
                    }
            }
        if (pass == 0)
            for (size_t o = 0; o < no; ++o) out[o] = sv[o] / sw[o];
    }
    free(P); free(integ); free(sw); free(sv); free(amb_n);
    return 0;
}
