// selector.cu -- SelectorNet inference (muti_kernel/train_gemini.py:14-39): the learned, content-adaptive kernel pick
// of SURVEY.md 8f row f2.  Three 3x3 / stride-2 / pad-1 convolutions (5 -> 32 -> 64 -> 128 channels) with eval-mode
// BatchNorm and ReLU, global average pooling, a 128 -> 10 linear layer; argmax of the logits is the kernel index the
// fused degrade kernel takes.  BatchNorm is folded into the convolution weights on the host (selector.py).
//
// The convolutions are GEMM-shaped (0.35 GFLOP per 256 x 256 patch, 91 x the time of the degrade kernel when left to
// the fp32 library path), so they run on the tensor cores -- but the pick must not depend on reduced precision: every
// product is evaluated as a 3xTF32 split (a = a_hi + a_lo, b = b_hi + b_lo in TF32; a_lo b_hi + a_hi b_lo + a_hi b_hi
// accumulated in fp32), which restores ~fp32 accuracy (measured against the fp32 library path in the tests) at three
// MMAs per product.  `mma.sync.m16n8k8.tf32` is used: a 128-pixel x 64-channel tile per CTA with K = 72 per step is
// far below what a tcgen05 / TMEM pipeline needs to pay off, and the layer is not the bottleneck of any headline path.
//
// conv_mma_kernel<NT>: one CTA = 32 x 8 output pixels x (8 NT) output channels of one patch, 8 warps, each warp two
// m16 tiles (4 output rows x 8 columns) x NT n8 tiles.  The input channels are walked in chunks of 8 (zero-filled
// beyond CIN: the 5-channel first layer uses one chunk); per chunk the 65 x 17 input pixels sit in shared memory
// channel-last with a pitch of 10 floats (conflict-free for the stride-2 A fragments) and the pre-split weights
// fragment-major (each lane's B values of a tap as NT conflict-free LDS.128); single-buffered, two CTAs per SM overlap
// each other's staging (cp.async) and MMA phases.
// Epilogue: bias + ReLU, NCHW store -- or, for the last layer, the per-CTA sum over its pixels (fixed order:
// deterministic) that pool_fc_kernel turns into logits.
#include "common.cuh"
#include "tma_util.cuh"

namespace kmsr {

namespace {

constexpr int kWarps = 8, kThreads = 32 * kWarps;
constexpr int kTH = 4 * kWarps, kTW = 8;         // output pixels per CTA: four rows per warp
constexpr int kIH = 2 * kTH + 1, kIW = 2 * kTW + 1;   // input pixels per CTA (stride 2, 3 x 3)
constexpr int kCS = 10;                          // floats per staged pixel (8 channels + 2: bank-conflict-free at stride 2)
constexpr int kInF = (kIH * kIW * kCS + 3) / 4 * 4;   // floats per staged input chunk (the weights behind it take 16-byte copies)

// ReLU as torch evaluates it: a NaN stays a NaN (fmaxf would return 0)
__device__ __forceinline__ float relu_keep_nan(float v) { return v < 0.0f ? 0.0f : v; }
__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct ConvArgs {
    const float* in;        // [N, CIN, H, W]
    const float* wsplit;    // [chunks][nblk][9 taps][NT quads][32 lanes][4]  (tf32 hi / lo parts, fragment-major, host-prepared)
    const float* bias;      // [COUT]
    float* out;             // [N, COUT, Ho, Wo], or nullptr when pooling
    float* pool_part;       // [N, tiles, COUT] per-CTA channel sums (last layer), or nullptr
    int CIN, COUT, H, W, Ho, Wo;
    int chunks;             // ceil(CIN / 8)
    int tiles_x, tiles;     // CTA tiles per patch
    int nblk;               // COUT / (8 NT)
};

template <int NT>
__global__ void __launch_bounds__(kThreads, 2)
conv_mma_kernel(const ConvArgs a) {
    // weights arrive fragment-major: [tap][quad NT][lane 32][4]; lane (g, t) finds its 4 NT B values of a tap -- index
    // part * 2 NT + 2 j + h for (hi | lo, n-tile j, k = t + 4 h) -- as NT conflict-free LDS.128
    constexpr int kWF = 9 * NT * 128;            // floats per staged weight chunk (hi and lo)
    extern __shared__ __align__(16) float csm[];
    float* in_s = csm;                           // [kInF]
    float* w_s = csm + kInF;                     // [kWF]   (single-buffered: two CTAs per SM overlap each other)
    const int tile = blockIdx.x % a.tiles, nb = blockIdx.x / a.tiles;
    const long long n = blockIdx.y;
    const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
    const int oy0 = ty * kTH, ox0 = tx * kTW;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;          // input coordinates of staged pixel (0, 0)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    auto stage = [&](int chunk) {
        // input: 8 channels x 65 x 17 pixels, zero outside the image and beyond CIN (cp.async src-size 0)
        const float* src = a.in + (n * a.CIN + 8 * chunk) * (long long)a.H * a.W;
        const uint32_t dst = smem_u32(in_s);
        const int nc = min(8, a.CIN - 8 * chunk);                 // channels of this chunk that exist (the rest stay zero)
        for (int e = tid; e < nc * kIH * kIW; e += kThreads) {
            const int c = e / (kIH * kIW), p = e - c * (kIH * kIW);
            const int py = p / kIW, px = p - py * kIW;
            const int gy = iy0 + py, gx = ix0 + px;
            const bool ok = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
            const float* s = ok ? src + ((long long)c * a.H + gy) * a.W + gx : a.in;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst + 4u * (p * kCS + c)), "l"(s), "r"(ok ? 4 : 0)
                         : "memory");
        }
        // weights of this chunk and channel block: contiguous, 16-byte copies
        const float* wsrc = a.wsplit + ((long long)chunk * a.nblk + nb) * kWF;
        const uint32_t wdst = smem_u32(w_s);
        for (int e = tid; e < kWF / 4; e += kThreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wdst + 16u * e), "l"(wsrc + 4 * e) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    float acc[2][NT][4];
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[m][j][r] = 0.0f;

    if (a.CIN % 8 != 0) {                        // channels beyond CIN are never staged: zero them once
        for (int e = tid; e < kInF; e += kThreads) in_s[e] = 0.0f;
        __syncthreads();
    }
    for (int chunk = 0; chunk < a.chunks; ++chunk) {
        stage(chunk);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const float* is = in_s;
        const float* ws = w_s;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - 3 * ky;
            // A fragments: warp rows 4 warp .. 4 warp + 3; m-tile m covers output rows 4 warp + 2 m (+0 for g, +1 for g + 8)
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int r0 = 4 * warp + 2 * m;
                const float* p0 = is + ((2 * r0 + ky) * kIW + 2 * g + kx) * kCS;          // output pixel (r0, g)
                const float* p1 = p0 + 2 * kIW * kCS;                                      // output pixel (r0 + 1, g)
                const float v[4] = {p0[t], p1[t], p0[t + 4], p1[t + 4]};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    ahi[m][r] = to_tf32(v[r]);
                    alo[m][r] = to_tf32(v[r] - __uint_as_float(ahi[m][r]));
                }
            }
            uint32_t bh[NT][2], bl[NT][2];
            {
                const float4* wq = reinterpret_cast<const float4*>(ws + tap * NT * 128) + lane;
#pragma unroll
                for (int q = 0; q < NT / 2; ++q) {
                    const float4 h4 = wq[q * 32], l4 = wq[(q + NT / 2) * 32];
                    bh[2 * q][0] = __float_as_uint(h4.x); bh[2 * q][1] = __float_as_uint(h4.y);
                    bh[2 * q + 1][0] = __float_as_uint(h4.z); bh[2 * q + 1][1] = __float_as_uint(h4.w);
                    bl[2 * q][0] = __float_as_uint(l4.x); bl[2 * q][1] = __float_as_uint(l4.y);
                    bl[2 * q + 1][0] = __float_as_uint(l4.z); bl[2 * q + 1][1] = __float_as_uint(l4.w);
                }
            }
            // three passes over the 2 NT independent accumulator tiles (small terms first): an MMA never waits for the
            // one issued just before it
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32(acc[m][j], alo[m], bh[j][0], bh[j][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32(acc[m][j], ahi[m], bl[j][0], bl[j][1]);
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int m = 0; m < 2; ++m) mma_tf32(acc[m][j], ahi[m], bh[j][0], bh[j][1]);
        }
        __syncthreads();                                         // the staging buffers are refilled by the next chunk
    }

    // ---- epilogue: acc[m][j] = {(row g, col 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)}: rows = output pixels, cols = channels
    const int cbase = nb * 8 * NT;
    if (a.pool_part == nullptr) {
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int oy = oy0 + 4 * warp + 2 * m + (r >> 1), ox = ox0 + g;
                    const int co = cbase + 8 * j + 2 * t + (r & 1);
                    if (oy < a.Ho && ox < a.Wo)
                        a.out[((n * a.COUT + co) * a.Ho + oy) * (long long)a.Wo + ox] = relu_keep_nan(acc[m][j][r] + __ldg(a.bias + co));
                }
        return;
    }
    // last layer: ReLU, then the sum over this CTA's pixels per channel, in a fixed order
    __shared__ float red[kWarps][64];
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int co = cbase + 8 * j + 2 * t + q;
            const float b = __ldg(a.bias + co);
            float s = 0.0f;
#pragma unroll
            for (int m = 0; m < 2; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int oy = oy0 + 4 * warp + 2 * m + h, ox = ox0 + g;
                    if (oy < a.Ho && ox < a.Wo) s += relu_keep_nan(acc[m][j][2 * h + q] + b);
                }
            // over the 8 pixel columns g (lane bits 2..4)
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (g == 0) red[warp][8 * j + 2 * t + q] = s;
        }
    __syncthreads();
    if (tid < 8 * NT) {
        float s = red[0][tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) s += red[w][tid];
        a.pool_part[(n * a.tiles + tile) * (long long)a.COUT + cbase + tid] = s;
    }
}

// logits[n] = fc_w . (sum_tiles pool_part[n] / (Ho Wo)) + fc_b; one warp per patch, fixed summation order
__global__ void __launch_bounds__(128)
pool_fc_kernel(const float* __restrict__ part, int tiles, int C, float inv_area, const float* __restrict__ fc_w,
               const float* __restrict__ fc_b, int classes, long long N, float* __restrict__ logits) {
    const long long n = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    for (int c = lane; c < C; c += 32) {
        float s = 0.0f;
        for (int tl = 0; tl < tiles; ++tl) s += part[(n * tiles + tl) * (long long)C + c];
        s *= inv_area;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            if (k < classes) acc[k] = fmaf(__ldg(fc_w + k * C + c), s, acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && k < classes) logits[n * classes + k] = v + __ldg(fc_b + k);
    }
}

template <int NT>
int launch_conv(const ConvArgs& a, long long N, cudaStream_t st) {
    constexpr size_t smem = (size_t)(kInF + 9 * NT * 128) * sizeof(float);
    auto kern = conv_mma_kernel<NT>;
    KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KMSR_REQUIRE(N <= 65535, KMSR_E_INVALID, "selector: more than 65535 patches per call");
    dim3 grid((unsigned)(a.tiles * a.nblk), (unsigned)N);
    kern<<<grid, kThreads, smem, st>>>(a);
    KMSR_LAUNCH_CHECK("conv_mma_kernel");
    return KMSR_OK;
}

inline int conv_out(int h) { return (h - 1) / 2 + 1; }       // (h + 2 - 3) / 2 + 1

}  // namespace

// pooling + linear layer for the tcgen05 path (selector_umma.cu), which produces the same per-tile channel sums
int launch_pool_fc(const float* part, int tiles, int C, float inv_area, const float* fc_w, const float* fc_b, int classes, long long N,
                   float* logits, cudaStream_t st) {
    pool_fc_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(part, tiles, C, inv_area, fc_w, fc_b, classes, N, logits);
    KMSR_LAUNCH_CHECK("pool_fc_kernel");
    return KMSR_OK;
}

// floats of the split weight blob of one layer: chunks x nblk x 9 taps x NT quads x 32 lanes x 4
long long selector_wsplit_floats(int cin, int cout) {
    const int NT = cout >= 64 ? 8 : 4;
    const int chunks = (cin + 7) / 8, nblk = cout / (8 * NT);
    return (long long)chunks * nblk * 9 * NT * 128;
}

long long selector_workspace(long long N, int H, int W) {
    const int h1 = conv_out(H), w1 = conv_out(W), h2 = conv_out(h1), w2 = conv_out(w1), h3 = conv_out(h2), w3 = conv_out(w2);
    const long long tiles3 = (long long)((h3 + kTH - 1) / kTH) * ((w3 + kTW - 1) / kTW);
    long long bytes = 0;
    bytes += (N * 32 * h1 * w1 * 4 + 255) / 256 * 256;
    bytes += (N * 64 * h2 * w2 * 4 + 255) / 256 * 256;
    bytes += (N * tiles3 * 128 * 4 + 255) / 256 * 256;
    return bytes + 256;
}

int launch_selector(const float* x, long long N, int H, int W, const float* w1, const float* b1, const float* w2,
                    const float* b2, const float* w3, const float* b3, const float* fc_w, const float* fc_b,
                    float* logits, void* workspace, long long workspace_bytes, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(workspace_bytes >= selector_workspace(N, H, W), KMSR_E_INVALID, "selector: workspace of %lld bytes, %lld needed",
                 workspace_bytes, selector_workspace(N, H, W));
    KMSR_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)w1 & 15) == 0 && ((uintptr_t)w2 & 15) == 0 && ((uintptr_t)w3 & 15) == 0,
                 KMSR_E_ALIGN, "selector: workspace / weight blobs not aligned");
    const int h1 = conv_out(H), wd1 = conv_out(W), h2 = conv_out(h1), wd2 = conv_out(wd1), h3 = conv_out(h2), wd3 = conv_out(wd2);
    char* ws = reinterpret_cast<char*>(workspace);
    float* a1 = reinterpret_cast<float*>(ws);
    ws += (N * 32 * h1 * wd1 * 4 + 255) / 256 * 256;
    float* a2 = reinterpret_cast<float*>(ws);
    ws += (N * 64 * h2 * wd2 * 4 + 255) / 256 * 256;
    float* part = reinterpret_cast<float*>(ws);
    auto fill = [](ConvArgs& c, const float* in, const float* w, const float* b, float* out, float* part, int cin, int cout, int h,
                   int wd, int nt) {
        c.in = in; c.wsplit = w; c.bias = b; c.out = out; c.pool_part = part;
        c.CIN = cin; c.COUT = cout; c.H = h; c.W = wd; c.Ho = conv_out(h); c.Wo = conv_out(wd);
        c.chunks = (cin + 7) / 8;
        c.tiles_x = (c.Wo + kTW - 1) / kTW;
        c.tiles = c.tiles_x * ((c.Ho + kTH - 1) / kTH);
        c.nblk = cout / (8 * nt);
    };
    ConvArgs c;
    fill(c, x, w1, b1, a1, nullptr, 5, 32, H, W, 4);
    int rc = launch_conv<4>(c, N, st);
    if (rc != KMSR_OK) return rc;
    fill(c, a1, w2, b2, a2, nullptr, 32, 64, h1, wd1, 8);
    rc = launch_conv<8>(c, N, st);
    if (rc != KMSR_OK) return rc;
    fill(c, a2, w3, b3, nullptr, part, 64, 128, h2, wd2, 8);
    rc = launch_conv<8>(c, N, st);
    if (rc != KMSR_OK) return rc;
    pool_fc_kernel<<<(unsigned)((N + 3) / 4), 128, 0, st>>>(part, c.tiles, 128, 1.0f / (float)(h3 * wd3), fc_w, fc_b, 10, N, logits);
    KMSR_LAUNCH_CHECK("pool_fc_kernel");
    return KMSR_OK;
}

}  // namespace kmsr
