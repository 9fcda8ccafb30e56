/*
 * oracle.c -- library-free CPU restatement of the pair-synthesis hot path.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded solely by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg.  The product path never touches it.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/kernel_from_lr_gan/).  The reference delegates its arithmetic to
 * PyTorch / NumPy / CPython; this file restates the published algorithms of those
 * calls (direct cross-correlation, 2x2 mean pooling, MT19937 + numpy's masked
 * rejection sampling, CPython's getrandbits rejection sampling, two-pass
 * mean / population std).  Parity pin: checked against outputs of the real
 * reference functions stored in tests/golden/ (tests/test_oracle_golden.py).
 *
 * Build: make -C oracle   (gcc -O2 -fPIC -shared -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- C_30:93-97 / C_31:74-78: per-band normalisation when the band sum is > 0.
 * torch sums in fp32 with a vectorised cascade whose order is ISA specific; the
 * sum is restated here in fp64 and rounded once to fp32 (differs from any fp32
 * order by at most one ulp of the sum).  Division is fp32 as in the reference. */
ORC_API void orc_normalize_kernel(const float* k, int C, int kh, int kw, float* kn) {
    for (int c = 0; c < C; ++c) {
        double s = 0.0;
        for (int i = 0; i < kh * kw; ++i) s += (double)k[c * kh * kw + i];
        float sf = (float)s;
        for (int i = 0; i < kh * kw; ++i)
            kn[c * kh * kw + i] = sf > 0.0f ? k[c * kh * kw + i] / sf : k[c * kh * kw + i];
    }
}

/* ---- C_30:104-124 in fp32, reference operation order:
 * replicate pad (clamp), cross-correlation (no flip) tap order row-major,
 * then int(log2(f)) cascaded 2x2 mean pools each computed ((a+b)+c)+d then *0.25
 * (ATen avg_pool2d CPU sums the window in raster order and divides by 4; /4 is exact).
 * kernel is used as given (already normalised by the caller).
 * scratch: none (allocates H*W floats internally). */
ORC_API int orc_degrade_f32(const float* img, int C, int H, int W, const float* kn, int kh, int kw,
                            int factor, int zero_pad, int decimate, float* out) {
    int ph = kh / 2, pw = kw / 2;
    int steps = 0;
    while ((1 << (steps + 1)) <= factor) ++steps;          /* int(np.log2(f)) C_30:121 */
    float* cur = (float*)malloc(sizeof(float) * (size_t)H * W);
    float* nxt = (float*)malloc(sizeof(float) * (size_t)H * W);
    if (!cur || !nxt) { free(cur); free(nxt); return -1; }
    for (int c = 0; c < C; ++c) {
        const float* x = img + (size_t)c * H * W;
        const float* k = kn + (size_t)c * kh * kw;
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                float acc = 0.0f;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) {
                        int sy = y + i - ph, sx = xx + j - pw;
                        float v;
                        if (zero_pad) v = (sy < 0 || sy >= H || sx < 0 || sx >= W) ? 0.0f : x[sy * W + sx];
                        else v = x[clampi(sy, 0, H - 1) * W + clampi(sx, 0, W - 1)];
                        acc = acc + k[i * kw + j] * v;
                    }
                cur[y * W + xx] = acc;
            }
        int h = H, w = W;
        if (decimate) {                                   /* train_gemini.py:134  out[:, :, ::f, ::f] */
            int ho = (H + factor - 1) / factor, wo = (W + factor - 1) / factor;
            for (int y = 0; y < ho; ++y)
                for (int xx = 0; xx < wo; ++xx) out[((size_t)c * ho + y) * wo + xx] = cur[(y * factor) * W + xx * factor];
            continue;
        }
        for (int s = 0; s < steps; ++s) {
            int h2 = h / 2, w2 = w / 2;
            for (int y = 0; y < h2; ++y)
                for (int xx = 0; xx < w2; ++xx) {
                    float a = cur[(2 * y) * w + 2 * xx], b = cur[(2 * y) * w + 2 * xx + 1];
                    float cc = cur[(2 * y + 1) * w + 2 * xx], d = cur[(2 * y + 1) * w + 2 * xx + 1];
                    nxt[y * w2 + xx] = (((a + b) + cc) + d) * 0.25f;
                }
            float* t = cur; cur = nxt; nxt = t;
            h = h2; w = w2;
        }
        memcpy(out + (size_t)c * h * w, cur, sizeof(float) * (size_t)h * w);
    }
    free(cur); free(nxt);
    return 0;
}

/* ---- same operator in fp64 ("truth" for the error budget): exact-in-double products of the
 * fp32-normalised kernel and fp32 pixels, double accumulation, double pooling. */
ORC_API int orc_degrade_f64(const float* img, int C, int H, int W, const float* kn, int kh, int kw,
                            int factor, int zero_pad, int decimate, double* out) {
    int ph = kh / 2, pw = kw / 2;
    int steps = 0;
    while ((1 << (steps + 1)) <= factor) ++steps;
    double* cur = (double*)malloc(sizeof(double) * (size_t)H * W);
    double* nxt = (double*)malloc(sizeof(double) * (size_t)H * W);
    if (!cur || !nxt) { free(cur); free(nxt); return -1; }
    for (int c = 0; c < C; ++c) {
        const float* x = img + (size_t)c * H * W;
        const float* k = kn + (size_t)c * kh * kw;
        for (int y = 0; y < H; ++y)
            for (int xx = 0; xx < W; ++xx) {
                double acc = 0.0;
                for (int i = 0; i < kh; ++i)
                    for (int j = 0; j < kw; ++j) {
                        int sy = y + i - ph, sx = xx + j - pw;
                        double v;
                        if (zero_pad) v = (sy < 0 || sy >= H || sx < 0 || sx >= W) ? 0.0 : (double)x[sy * W + sx];
                        else v = (double)x[clampi(sy, 0, H - 1) * W + clampi(sx, 0, W - 1)];
                        acc += (double)k[i * kw + j] * v;
                    }
                cur[y * W + xx] = acc;
            }
        int h = H, w = W;
        if (decimate) {
            int ho = (H + factor - 1) / factor, wo = (W + factor - 1) / factor;
            for (int y = 0; y < ho; ++y)
                for (int xx = 0; xx < wo; ++xx) out[((size_t)c * ho + y) * wo + xx] = cur[(y * factor) * W + xx * factor];
            continue;
        }
        for (int s = 0; s < steps; ++s) {
            int h2 = h / 2, w2 = w / 2;
            for (int y = 0; y < h2; ++y)
                for (int xx = 0; xx < w2; ++xx)
                    nxt[y * w2 + xx] = (cur[(2 * y) * w + 2 * xx] + cur[(2 * y) * w + 2 * xx + 1] +
                                        cur[(2 * y + 1) * w + 2 * xx] + cur[(2 * y + 1) * w + 2 * xx + 1]) * 0.25;
            double* t = cur; cur = nxt; nxt = t;
            h = h2; w = w2;
        }
        memcpy(out + (size_t)c * h * w, cur, sizeof(double) * (size_t)h * w);
    }
    free(cur); free(nxt);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * MT19937 (Matsumoto & Nishimura 2002 reference algorithm) -- shared by numpy's legacy
 * RandomState and CPython's `random`.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t mt[624]; int idx; } orc_mt;

static void mt_init_genrand(orc_mt* s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}
static void mt_init_by_array(orc_mt* s, const uint32_t* key, int len) {
    mt_init_genrand(s, 19650218u);
    int i = 1, j = 0;
    int k = 624 > len ? 624 : len;
    for (; k; --k) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
        if (j >= len) j = 0;
    }
    for (k = 623; k; --k) {
        s->mt[i] = (s->mt[i] ^ ((s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { s->mt[0] = s->mt[623]; i = 1; }
    }
    s->mt[0] = 0x80000000u;
    s->idx = 624;
}
static uint32_t mt_next(orc_mt* s) {
    if (s->idx >= 624) {
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (s->mt[k] & 0x80000000u) | (s->mt[(k + 1) % 624] & 0x7fffffffu);
            s->mt[k] = s->mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

/* ---- E:190 + E:72: np.random.seed(seed); n scalar draws np.random.randint(0, high).
 * numpy legacy path: seed -> init_genrand(seed); randint(0,high) with range rng=high-1 < 2^32 ->
 * smallest all-ones mask >= rng, draw 32-bit words, keep the first with (word & mask) <= rng. */
ORC_API void orc_numpy_randint_stream(uint32_t seed, int64_t high, int64_t n, int64_t* out) {
    orc_mt s; mt_init_genrand(&s, seed);
    uint32_t rng = (uint32_t)(high - 1);
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    for (int64_t i = 0; i < n; ++i) {
        if (rng == 0) { out[i] = 0; continue; }
        uint32_t v;
        do { v = mt_next(&s) & mask; } while (v > rng);
        out[i] = (int64_t)v;
    }
}
/* rs=RandomState(seed); kidx=rs.randint(0,nk,n); nidx=rs.randint(0,npool,n) -- one stream. */
ORC_API void orc_numpy_two_randint_vectors(uint32_t seed, int64_t nk, int64_t npool, int64_t n,
                                            int32_t* kidx, int32_t* nidx) {
    orc_mt s; mt_init_genrand(&s, seed);
    int64_t highs[2] = {nk, npool};
    int32_t* outs[2] = {kidx, nidx};
    for (int p = 0; p < 2; ++p) {
        uint32_t rng = (uint32_t)(highs[p] - 1), mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        for (int64_t i = 0; i < n; ++i) {
            if (rng == 0) { outs[p][i] = 0; continue; }
            uint32_t v;
            do { v = mt_next(&s) & mask; } while (v > rng);
            outs[p][i] = (int32_t)v;
        }
    }
}

/* ---- D:65 + D:49-50: random.seed(seed) (int seed -> init_by_array of its 32-bit limbs),
 * random.randint(0, m) = _randbelow(m+1): k = bit_length(m+1); r = getrandbits(k) = word >> (32-k);
 * redraw while r >= m+1.  Per sample: top (bound H-crop) then left (bound W-crop). */
static uint32_t py_randbelow(orc_mt* s, uint32_t n) {
    int k = 0; for (uint32_t t = n; t; t >>= 1) ++k;
    uint32_t r;
    do { r = mt_next(s) >> (32 - k); } while (r >= n);
    return r;
}
ORC_API void orc_python_crop_offsets(uint32_t seed, const int32_t* hw /* [files][2] */, int64_t files,
                                      int crop, int samples_per_file, int32_t* out /* [files*spf][2] */) {
    orc_mt s; uint32_t key[1] = {seed}; mt_init_by_array(&s, key, 1);
    int64_t o = 0;
    for (int64_t f = 0; f < files; ++f)
        for (int k = 0; k < samples_per_file; ++k) {
            out[2 * o] = (int32_t)py_randbelow(&s, (uint32_t)(hw[2 * f] - crop + 1));
            out[2 * o + 1] = (int32_t)py_randbelow(&s, (uint32_t)(hw[2 * f + 1] - crop + 1));
            ++o;
        }
}

/* ---- D:88 + D:51: pool[m] = (geo - den)[:, top:top+crop, left:left+crop] */
ORC_API void orc_crop_sub(const float* geo, const float* den, int C, int H, int W, int top, int left,
                          int crop, float* out) {
    for (int c = 0; c < C; ++c)
        for (int y = 0; y < crop; ++y)
            for (int x = 0; x < crop; ++x) {
                size_t s = ((size_t)c * H + top + y) * W + left + x;
                out[((size_t)c * crop + y) * crop + x] = geo[s] - den[s];
            }
}

/* ---- E:72-74: blurred + pool[idx]  (scale==1) / sigma-scaled variant as one fused multiply-add. */
ORC_API void orc_add_noise(const float* blurred, const float* noise, const float* scale /* [C] or NULL */,
                           int C, int hw, float* out) {
    for (int c = 0; c < C; ++c)
        for (int i = 0; i < hw; ++i)
            out[c * hw + i] = scale ? fmaf(scale[c], noise[c * hw + i], blurred[c * hw + i])
                                    : blurred[c * hw + i] + noise[c * hw + i];
}

/* ---- data_mean_std.py:32-33: nanmean / nanstd (ddof=0) over (H,W) per band, two-pass, in fp64. */
ORC_API void orc_band_stats_f64(const float* x, int C, int64_t hw, double* mean, double* std) {
    for (int c = 0; c < C; ++c) {
        const float* p = x + (size_t)c * hw;
        double s = 0.0; int64_t n = 0;
        for (int64_t i = 0; i < hw; ++i) if (!isnan(p[i])) { s += p[i]; ++n; }
        double m = n ? s / (double)n : NAN;
        double q = 0.0;
        for (int64_t i = 0; i < hw; ++i) if (!isnan(p[i])) { double d = p[i] - m; q += d * d; }
        mean[c] = m;
        std[c] = n ? sqrt(q / (double)n) : NAN;
    }
}

/* ---- A_00_patch_cutter_universal.py:102-113: -9999 -> NaN, NIR window, NaN all bands outside. */
ORC_API void orc_water_mask(const float* data, int C, int64_t hw, int nir, float tmin, float tmax,
                            float* masked) {
    for (int64_t i = 0; i < hw; ++i) {
        float v = data[(size_t)nir * hw + i];
        if (v == -9999.0f) v = NAN;
        int water = (v >= tmin) && (v <= tmax);           /* NaN compares false */
        for (int c = 0; c < C; ++c) {
            float d = data[(size_t)c * hw + i];
            if (d == -9999.0f) d = NAN;
            masked[(size_t)c * hw + i] = water ? d : NAN;
        }
    }
}

/* ---- A_00_patch_cutter_universal.py:152-183: stride=int(P*ratio); grid (H-P)//stride+1;
 * keep[i][j] = (NaN count of the window == 0) for nan_threshold 0 (general: ratio > thr drops). */
ORC_API void orc_keep_mask(const float* masked, int C, int H, int W, int P, int stride, double nan_thr,
                           uint8_t* keep, int hp, int wp) {
    for (int i = 0; i < hp; ++i)
        for (int j = 0; j < wp; ++j) {
            int64_t nan = 0;
            for (int c = 0; c < C; ++c)
                for (int y = 0; y < P; ++y)
                    for (int x = 0; x < P; ++x)
                        nan += isnan(masked[((size_t)c * H + i * stride + y) * W + j * stride + x]) ? 1 : 0;
            double ratio = (double)nan / ((double)C * P * P);
            keep[i * wp + j] = !(ratio > nan_thr);
        }
}
