// degrade_tiled.cu -- generic fused blur + downsample + noise kernel (any k, factor, H, W, pad mode).
//
// Restates apply_kernel_degradation (kernel_from_lr_gan/C_30apply_kernel_to_landsat.py:68-124 ==
// C_31apply_muti_kernel_to_landsat.py:59-97) with the box mean folded into the kernel
// (prepare.cu) and E_make_train_data.py:72-74 / train_gemini.py:137 fused into the epilogue:
//
//   lr[n,c,Y,X] = sum_{u<KH, v<KW} K'[u,v] * hr[n,c, pad(s*Y+u-pt), pad(s*X+v-pl)]  (+ scale*noise)
//
// One CTA computes a (8*TY) x 32 output tile of one band.  The input window is staged in shared
// memory in POLYPHASE order -- column cc of the window lives at [row][cc % s][cc / s] -- so that
// the 32 lanes of a warp (32 adjacent outputs, s input pixels apart) read 32 adjacent words:
// conflict-free LDS without vector loads.  Pixels are stored as (x - pivot), pivot = one pixel of
// the tile, so that the fp32 accumulation error scales with the patch's dynamic range instead of
// its radiance level; pivot * sum(K') is added back in the epilogue.
//
// This is the fallback / sweep kernel; the headline shape (k=13, s=8, W=256) runs degrade_tma.cu.
#include "common.cuh"

namespace kmsr {

struct TiledParams {
    const float* hr;
    const long long* patch_offsets;
    const float* comp;
    const float* dsum;
    const int* kidx;
    const float* sigma;
    const float* pool;
    const int* nidx;
    float* lr;
    long long sN, sC, sH;
    int C, H, W, Ho, Wo;
    int KH, KW, KWp, S, pt, pl;
    int pad_mode, noise_mode;
    int tiles_x, tiles_y;
    int IR;        // staged rows
    int ICp;       // staged columns (window of the last lane, weights padded to KWp)
    int IC;        // columns that carry real data
    int PM;        // words per phase plane row
    int RP;        // words per staged row  (= S * PM)
};

template <int S_>
__device__ __forceinline__ void split_col(int cc, int S, int& t, int& m) {
    if (S_ > 0) { t = cc % S_; m = cc / S_; }
    else        { t = cc % S;  m = cc / S;  }
}

template <int TY, int S_>
__global__ void __launch_bounds__(256)
degrade_tiled_kernel(const TiledParams p) {
    extern __shared__ __align__(16) float smem[];
    float* wts = smem;                               // [KH][KWp]
    float* tile = smem + ((p.KH * p.KWp + 3) & ~3);  // [IR][S][PM]
    const int S = S_ > 0 ? S_ : p.S;

    const int tiles = p.tiles_x * p.tiles_y;
    const long long band = blockIdx.x / tiles;
    const int tl = blockIdx.x % tiles;
    const int oy0 = (tl / p.tiles_x) * (8 * TY);
    const int ox0 = (tl % p.tiles_x) * 32;
    const long long n = band / p.C;
    const int c = (int)(band % p.C);
    const int kid = p.kidx ? p.kidx[n] : 0;

    const float* img = p.hr + (p.patch_offsets ? p.patch_offsets[n] : n * p.sN) + (long long)c * p.sC;
    const float* kc = p.comp + ((long long)kid * p.C + c) * p.KH * p.KWp;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // pivot: the (clamped) top-left pixel of this tile's window
    const int gy0 = min(max(oy0 * S - p.pt, 0), p.H - 1);
    const int gx0 = min(max(ox0 * S - p.pl, 0), p.W - 1);
    float pv = __ldg(img + (long long)gy0 * p.sH + gx0);
    if (!isfinite(pv)) pv = 0.0f;

    for (int i = threadIdx.x; i < p.KH * p.KWp; i += 256) wts[i] = __ldg(kc + i);

    // stage the window: warps over row pairs, lanes over columns (coalesced global reads).  The loads
    // of a 2 x 256 element slab are all issued before the first store so that 16 requests per thread
    // are in flight (one-at-a-time loads left the kernel latency-bound at 3 % of HBM bandwidth).
    const bool zero_pad = p.pad_mode == KMSR_PAD_ZERO;
    for (int r0 = warp * 2; r0 < p.IR; r0 += 16) {
        for (int cc0 = 0; cc0 < p.ICp; cc0 += 256) {
            float v[2][8];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = r0 + rr;
                const int gy = oy0 * S + r - p.pt;
                const bool yin = gy >= 0 && gy < p.H;
                const float* row = img + (long long)min(max(gy, 0), p.H - 1) * p.sH;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cc = cc0 + lane + 32 * j;
                    const int gx = ox0 * S + cc - p.pl;
                    const bool xin = gx >= 0 && gx < p.W;
                    float t = 0.0f;
                    if (r < p.IR && cc < p.IC) {
                        if (!zero_pad || (yin && xin)) t = __ldg(row + min(max(gx, 0), p.W - 1)) - pv;
                        else t = -pv;                   // a zero-padded pixel, pivot-shifted
                    }
                    v[rr][j] = t;
                }
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = r0 + rr;
                if (r >= p.IR) continue;
                float* trow = tile + r * p.RP;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cc = cc0 + lane + 32 * j;
                    if (cc >= p.ICp) continue;
                    int t, m;
                    split_col<S_>(cc, S, t, m);
                    trow[t * p.PM + m] = v[rr][j];
                }
            }
        }
    }
    __syncthreads();

    float acc[TY];
#pragma unroll
    for (int ty = 0; ty < TY; ++ty) acc[ty] = 0.0f;

    const int row0 = S * warp * TY;
    const int nrows = S * (TY - 1) + p.KH;
    const int nv4 = p.KWp >> 2;
#pragma unroll 1
    for (int r = 0; r < nrows; ++r) {
        const float* trow = tile + (row0 + r) * p.RP + lane;
#pragma unroll 1
        for (int v4 = 0; v4 < nv4; ++v4) {
            float d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int t, m;
                split_col<S_>(4 * v4 + j, S, t, m);
                // taps KW .. KWp-1 are row padding (weight 0): they must not meet the neighbouring window's pixels, or a
                // NaN / Inf up to three columns right of an output's window would poison it (0 * NaN)
                d[j] = 4 * v4 + j < p.KW ? trow[t * p.PM + m] : 0.0f;
            }
#pragma unroll
            for (int ty = 0; ty < TY; ++ty) {
                const int u = r - S * ty;
                if (u >= 0 && u < p.KH) {
                    const float4 w = *reinterpret_cast<const float4*>(wts + u * p.KWp + 4 * v4);
                    acc[ty] = fmaf(w.x, d[0], acc[ty]);
                    acc[ty] = fmaf(w.y, d[1], acc[ty]);
                    acc[ty] = fmaf(w.z, d[2], acc[ty]);
                    acc[ty] = fmaf(w.w, d[3], acc[ty]);
                }
            }
        }
    }

    const int ox = ox0 + lane;
    if (ox >= p.Wo) return;
    const float ds = __ldg(p.dsum + (long long)kid * p.C + c);
    const long long ohw = (long long)p.Ho * p.Wo;
    float scale = 1.0f;
    const float* nz = nullptr;
    if (p.noise_mode != KMSR_NOISE_NONE) {
        nz = p.pool + ((long long)p.nidx[n] * p.C + c) * ohw;
        if (p.noise_mode == KMSR_NOISE_SIGMA) scale = __ldg(p.sigma + (long long)kid * p.C + c);
    }
    float* out = p.lr + band * ohw;
#pragma unroll
    for (int ty = 0; ty < TY; ++ty) {
        const int oy = oy0 + warp * TY + ty;
        if (oy >= p.Ho) continue;
        float res = pv + fmaf(pv, ds, acc[ty]);
        if (nz) res = fmaf(scale, __ldg(nz + (long long)oy * p.Wo + ox), res);
        out[(long long)oy * p.Wo + ox] = res;
    }
}

template <int TY, int S_>
static int launch_one(const TiledParams& p, long long blocks, size_t smem, cudaStream_t st) {
    auto kern = degrade_tiled_kernel<TY, S_>;
    KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)blocks, 256, smem, st>>>(p);
    KMSR_LAUNCH_CHECK("degrade_tiled_kernel");
    return KMSR_OK;
}

template <int TY>
static int launch_s(const TiledParams& p, long long blocks, size_t smem, cudaStream_t st) {
    switch (p.S) {
        case 1: return launch_one<TY, 1>(p, blocks, smem, st);
        case 2: return launch_one<TY, 2>(p, blocks, smem, st);
        case 4: return launch_one<TY, 4>(p, blocks, smem, st);
        case 8: return launch_one<TY, 8>(p, blocks, smem, st);
        case 16: return launch_one<TY, 16>(p, blocks, smem, st);
        default: return launch_one<TY, 0>(p, blocks, smem, st);
    }
}

int launch_degrade_tiled(const DegradeArgs& a, cudaStream_t st) {
    const Geometry& g = a.g;
    TiledParams p;
    p.hr = a.hr; p.patch_offsets = a.patch_offsets; p.comp = a.comp; p.dsum = a.dsum;
    p.kidx = a.kidx; p.sigma = a.sigma; p.pool = a.pool; p.nidx = a.nidx; p.lr = a.lr;
    p.sN = a.sN; p.sC = a.sC; p.sH = a.sH;
    p.C = a.C; p.H = a.H; p.W = a.W; p.Ho = g.Ho; p.Wo = g.Wo;
    p.KH = g.KH; p.KW = g.KW; p.KWp = g.KWp; p.S = g.stride; p.pt = g.pt; p.pl = g.pl;
    p.pad_mode = a.pad_mode; p.noise_mode = a.noise_mode;

    const int S = g.stride;
    p.IC = S * 31 + g.KW;
    p.ICp = S * 31 + g.KWp;
    const int M = (p.ICp + S - 1) / S;
    // words per phase plane: M rounded so that the staging stores of 32 adjacent columns
    // (S phases x 32/S words) fall in 32 different banks when S divides 32, odd otherwise
    if (S <= 32 && 32 % S == 0) p.PM = M + (((32 / S) % 32 - M) % 32 + 32) % 32;     // smallest >= M, == 32/S (mod 32)
    else p.PM = M | 1;
    p.RP = S * p.PM;

    int dev = 0, max_smem = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));

    const size_t wbytes = (size_t)((g.KH * g.KWp + 3) & ~3) * sizeof(float);
    auto tile_bytes = [&](int ty) { return (size_t)(S * (8 * ty - 1) + g.KH) * p.RP * sizeof(float); };
    int TY = 1;
    if (g.Ho > 8 && wbytes + tile_bytes(2) <= 100 * 1024) TY = 2;
    const size_t smem = wbytes + tile_bytes(TY);
    KMSR_REQUIRE(smem <= (size_t)max_smem, KMSR_E_UNSUPPORTED,
                 "degrade (tiled): window %dx%d at stride %d needs %zu B of shared memory (max %d)",
                 g.KH, g.KW, S, smem, max_smem);
    p.IR = S * (8 * TY - 1) + g.KH;
    p.tiles_x = (g.Wo + 31) / 32;
    p.tiles_y = (g.Ho + 8 * TY - 1) / (8 * TY);
    const long long blocks = a.N * a.C * p.tiles_x * p.tiles_y;
    KMSR_REQUIRE(blocks < (1ll << 31), KMSR_E_INVALID, "degrade (tiled): %lld CTAs exceed the grid limit", blocks);
    if (blocks == 0) return KMSR_OK;
    set_algo("tiled");
    return TY == 2 ? launch_s<2>(p, blocks, smem, st) : launch_s<1>(p, blocks, smem, st);
}

}  // namespace kmsr
