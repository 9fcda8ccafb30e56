// degrade_reg.cu -- register-tile fused blur + box-mean downsample + noise kernel for the FP32-bound shapes of the
// sweep (BASELINE config 5: factor 2 with any kernel, factor 4 with the larger kernels).
//
// Same arithmetic as the other degrade kernels (C_30apply_kernel_to_landsat.py:68-124 with the box mean folded into
// a stride-S composite kernel K' of (k + S - 1)^2 taps, E_make_train_data.py:72-74 / train_gemini.py:137 noise in the
// epilogue).  At factor 2 a composite window is (k + 1)^2 / 4 FMAs per HR pixel -- 36 at k = 11, 256 at k = 31 -- and
// the path is bound by FP32 issue slots, not by HBM.  The row-streaming kernel (degrade_stream.cu) spends its step on
// bookkeeping there: S input rows per step give each lane Q * TP * TX FFMA2 (144 at k = 11) against ~450 other
// instructions (ring waits, edge substitution, re-pairing, the rotation of Q accumulator sets, the shuffle
// reduce-scatter): FFMA2 is a quarter of what it issues.  This kernel is a plain register-tiled stencil instead:
//
//  * a group of threads owns one HR tile (256 x 64 pixels of one band, or the whole band when it is smaller) plus its
//    halo in shared memory; the halo is made while loading (cp.async per element with clamped or zero-filled
//    coordinates), so replicate / zero padding, any W and H, any strides cost nothing later;
//  * a thread owns TXO x TYO LR outputs (8 x 4 at factor 2, 4 x 2 at factor 4) in registers and walks the input rows
//    of its window once: one row segment (pairs of pixels, LDS.64), pivot shift (FADD2), then for each of its output
//    rows that this input row meets the composite-kernel row u = rr - S*yo from shared memory (broadcast LDS.128) and
//    TP x TXO FFMA2; the next row's segment is loaded while the current one is multiplied.  No ring, no shuffles, no
//    accumulator rotation: ~75 % of the issued instructions are FFMA2;
//  * shared-memory layout: a row is stored in blocks of 8 pixel pairs, "block-column major" (pair p sits at
//    (p mod 8) * NB + p div 8), so the 16 lanes of a half-warp -- whose segments start 8 pairs apart -- read
//    consecutive 8-byte words: conflict-free LDS.64 without any alignment requirement between taps and pairs;
//  * two CTAs of 128 threads per SM, each single-buffered: one loads while the other computes.
#include <stdlib.h>

#include "common.cuh"
#include "tma_util.cuh"

namespace kmsr {

namespace {

template <int K, int S, int TYO_ = (S == 2 ? 4 : 2), int THREADS_ = 128>
struct RCfg {
    static constexpr int THREADS = THREADS_;
    static constexpr int KW = K + S - 1;           // composite taps per row / column (even)
    static constexpr int TP = KW / 2;              // tap pairs
    static constexpr int PAD = K / 2;
    static constexpr int HALO = KW - S;            // extra input rows / columns of a window beyond S per output
    static constexpr int TXO = S == 2 ? 8 : 4;     // LR columns per thread: 16 HR columns per lane either way
    static constexpr int TYO = TYO_;               // LR rows per thread (8 HR rows per thread row at 4 / 2)
    static constexpr int NPX = (S * (TXO - 1) + KW) / 2;   // pixel pairs of a thread's row segment
    static_assert(K % 2 == 1 && (S == 2 || S == 4), "odd kernel, factor 2 or 4");
};

struct RegArgs {
    const float* hr;
    long long sN, sC, sH;
    const float* comp;      // [nK, C, KW, KWp]
    int KWp;
    const float* dsum;
    const int* kidx;
    const float* sigma;
    const float* pool;
    const int* nidx;
    float* lr;
    long long nitems;       // bands * tiles
    int C, H, W, Ho, Wo;
    int tiles_x, tiles;     // tiles per band
    int GX, GY;             // thread grid of a group
    int gthreads, groups;   // threads per group (multiple of 32), groups per CTA
    int TWl, THl;           // LR tile extent (GX * TXO, GY * TYO)
    int RW, RH;             // staged region in HR pixels (tile + halo)
    int NB;                 // 8-pair blocks per staged row
    int pitchF;             // floats per staged row (16 * NB, + 2 for narrow groups)
    int groupFloats;        // shared-memory floats per group (tile + composite kernel)
    int pad_mode, noise_mode;
};

// groups of 64 or more threads meet at a named barrier (at most two such groups per CTA); a group of one warp at __syncwarp
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    if (nthreads == 32) __syncwarp();
    else if (id == 0) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory");
}

template <int K, int S, int TYO_, int THREADS_>
__global__ void __launch_bounds__(THREADS_, 2)
degrade_reg_kernel(const RegArgs a) {
    using G = RCfg<K, S, TYO_, THREADS_>;
    extern __shared__ __align__(16) float rsm[];
    const int grp = threadIdx.x / a.gthreads, tid = threadIdx.x - grp * a.gthreads;
    const long long item = (long long)blockIdx.x * a.groups + grp;
    if (grp >= a.groups || item >= a.nitems) return;       // whole groups leave together
    float* tile = rsm + (size_t)grp * a.groupFloats;
    float* wsm = tile + (size_t)a.RH * a.pitchF;

    const long long band = item / a.tiles;
    const int t = (int)(item - band * a.tiles);
    const int ty = t / a.tiles_x, tx = t - ty * a.tiles_x;
    const long long n = band / a.C;
    const int c = (int)(band - n * a.C);
    const int X0 = tx * a.TWl, Y0 = ty * a.THl;             // LR origin of the tile
    const int hx0 = S * X0 - G::PAD, hy0 = S * Y0 - G::PAD; // HR coordinates of staged (row 0, column 0)
    const float* src = a.hr + n * a.sN + (long long)c * a.sC;
    const bool replicate = a.pad_mode == KMSR_PAD_REPLICATE;
    const bool noisy = a.noise_mode != KMSR_NOISE_NONE;

    // ---- per-band parameters and the composite kernel ----
    const int kid = a.kidx ? __ldg(a.kidx + n) : 0;
    {
        const float* kc = a.comp + ((long long)kid * a.C + c) * (G::KW * a.KWp);
        const uint32_t dst = smem_u32(wsm);
        for (int e = tid; e < G::KW * a.KWp / 4; e += a.gthreads)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * e), "l"(kc + 4 * e) : "memory");
    }
    // ---- tile + halo: column-parallel, row loop; clamp (replicate) or zero fill while loading.  The rows are split
    // into the halo above the image, the image rows (pointer increments only) and the halo below it.
    {
        const uint32_t tdst = smem_u32(tile);
        const int rtop = min(max(-hy0, 0), a.RH);              // staged rows above image row 0
        const int rbot = min(max(a.H - hy0, rtop), a.RH);      // first staged row below the image
        const uint32_t pitchB = 4u * a.pitchF;
        auto cp4 = [](uint32_t dst, const float* s, int nbytes) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(s), "r"(nbytes) : "memory");
        };
        for (int cc = tid; cc < a.RW; cc += a.gthreads) {
            const int gx = hx0 + cc;
            const bool xin = gx >= 0 && gx < a.W;
            const int p = cc >> 1;
            uint32_t d = tdst + 4u * (2 * ((p & 7) * a.NB + (p >> 3)) + (cc & 1));
            const float* s0 = src + min(max(gx, 0), a.W - 1);
            const int halo_bytes = replicate ? 4 : 0;                    // src-size 0: zero fill (zero padding)
            const int row_bytes = (replicate || xin) ? 4 : 0;
            int r = 0;
            for (; r < rtop; ++r, d += pitchB) cp4(d, s0, halo_bytes);
            const float* sp = s0 + (long long)(hy0 + r) * a.sH;
#pragma unroll 4
            for (; r < rbot; ++r, d += pitchB, sp += a.sH) cp4(d, sp, row_bytes);
            const float* sl = s0 + (long long)(a.H - 1) * a.sH;
            for (; r < a.RH; ++r, d += pitchB) cp4(d, sl, halo_bytes);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float ds = __ldg(a.dsum + (long long)kid * a.C + c);
    const float scale = a.noise_mode == KMSR_NOISE_SIGMA ? __ldg(a.sigma + (long long)kid * a.C + c) : 1.0f;
    const int nid = noisy ? __ldg(a.nidx + n) : 0;
    asm volatile("cp.async.wait_all;" ::: "memory");
    group_sync(grp, a.gthreads);

    const int lx = tid % a.GX, ly = tid / a.GX;
    if (ly >= a.GY) return;                                  // padding threads of a group (gthreads rounded up to 32)
    // physical float offset of logical pair (8 lx + i): 2 * ((i & 7) * NB + lx + (i >> 3))
    const float* trow = tile + (size_t)(S * G::TYO * ly) * a.pitchF + 2 * lx;
    // pivot: the pixel under the centre tap of the thread's first output
    float pv;
    {
        const int pc = S * G::TXO * lx + G::PAD, pp = pc >> 1;
        pv = tile[(size_t)(S * G::TYO * ly + G::PAD) * a.pitchF + 2 * ((pp & 7) * a.NB + (pp >> 3)) + (pc & 1)];
        if (!isfinite(pv)) pv = 0.0f;
    }
    const u64 npv2 = pack2(-pv, -pv);

    u64 acc[G::TYO][G::TXO];
#pragma unroll
    for (int yo = 0; yo < G::TYO; ++yo)
#pragma unroll
        for (int x = 0; x < G::TXO; ++x) acc[yo][x] = 0ull;

    constexpr int NROWS = S * G::TYO + G::HALO;              // input rows a thread's outputs touch (even: K is odd)
    static_assert(NROWS % 2 == 0, "row loop is unrolled by two");
    // row rr of the thread's window -> pivot-shifted pixel pairs
    auto load_row = [&](const int rr, u64 (&P)[G::NPX]) {
        const float* prow = trow + (size_t)rr * a.pitchF;
#pragma unroll
        for (int i = 0; i < G::NPX; ++i)
            P[i] = add2(*reinterpret_cast<const u64*>(prow + 2 * ((i & 7) * a.NB + (i >> 3))), npv2);
    };
    auto mac_row = [&](const int rr, const u64 (&P)[G::NPX]) {
#pragma unroll
        for (int yo = 0; yo < G::TYO; ++yo) {
            const int u = rr - S * yo;                       // composite row this input row meets for output row yo
            if (u >= 0 && u < G::KW) {                       // warp-uniform
                const float* wrow = wsm + u * a.KWp;
#pragma unroll
                for (int t4 = 0; t4 < (G::TP + 1) / 2; ++t4) {
                    const ulonglong2 w2 = reinterpret_cast<const ulonglong2*>(wrow)[t4];
#pragma unroll
                    for (int x = 0; x < G::TXO; ++x) {
                        acc[yo][x] = fma2(w2.x, P[(S * x) / 2 + 2 * t4], acc[yo][x]);
                        if (2 * t4 + 1 < G::TP) acc[yo][x] = fma2(w2.y, P[(S * x) / 2 + 2 * t4 + 1], acc[yo][x]);
                    }
                }
            }
        }
    };
    // two row buffers: the next row's LDS.64 are in flight while the current row is multiplied
    u64 PA[G::NPX], PB[G::NPX];
    load_row(0, PA);
#pragma unroll 1
    for (int rr = 0; rr < NROWS; rr += 2) {
        load_row(rr + 1, PB);
        mac_row(rr, PA);
        if (rr + 2 < NROWS) load_row(rr + 2, PA);
        mac_row(rr + 1, PB);
    }

    // ---- epilogue: even + odd taps, pivot back, noise, store ----
    const long long ohw = (long long)a.Ho * a.Wo;
    float* outb = a.lr + band * ohw;
    const float* nz = a.pool + ((long long)nid * a.C + c) * ohw;
    const int Xb = X0 + G::TXO * lx;
    const bool vec = (a.Wo % 4 == 0) && Xb + G::TXO <= a.Wo && (((uintptr_t)a.lr & 15) == 0) &&
                     (!noisy || ((uintptr_t)a.pool & 15) == 0);
#pragma unroll
    for (int yo = 0; yo < G::TYO; ++yo) {
        const int Y = Y0 + G::TYO * ly + yo;
        if (Y >= a.Ho) continue;
        float res[G::TXO];
#pragma unroll
        for (int x = 0; x < G::TXO; ++x) res[x] = pv + fmaf(pv, ds, lo2(acc[yo][x]) + hi2(acc[yo][x]));
        float* o = outb + (long long)Y * a.Wo + Xb;
        const float* z = nz + (long long)Y * a.Wo + Xb;
        if (vec) {
#pragma unroll
            for (int x4 = 0; x4 < G::TXO / 4; ++x4) {
                float4 r4 = make_float4(res[4 * x4], res[4 * x4 + 1], res[4 * x4 + 2], res[4 * x4 + 3]);
                if (noisy) {
                    const float4 q = __ldg(reinterpret_cast<const float4*>(z) + x4);
                    r4.x = fmaf(scale, q.x, r4.x); r4.y = fmaf(scale, q.y, r4.y);
                    r4.z = fmaf(scale, q.z, r4.z); r4.w = fmaf(scale, q.w, r4.w);
                }
                reinterpret_cast<float4*>(o)[x4] = r4;
            }
        } else {
#pragma unroll
            for (int x = 0; x < G::TXO; ++x) {
                if (Xb + x < a.Wo) o[x] = noisy ? fmaf(scale, __ldg(z + x), res[x]) : res[x];
            }
        }
    }
}

template <int K, int S, int TYO_, int THREADS_>
int launch_reg(const DegradeArgs& a, RegArgs& t, cudaStream_t st) {
    using G = RCfg<K, S, TYO_, THREADS_>;
    int dev = 0, max_smem = 0, sm_smem = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&sm_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    // thread grid of a group: up to 16 lanes across (16 HR columns each), up to THREADS / 16 rows of threads.
    t.GX = (t.Wo + G::TXO - 1) / G::TXO;
    if (t.GX > 16) t.GX = 16;
    int gy_max = (t.Ho + G::TYO - 1) / G::TYO;
    if (gy_max > G::THREADS / 16) gy_max = G::THREADS / 16;
    auto geometry = [&](int gy) {
        t.GY = gy;
        t.TWl = t.GX * G::TXO; t.THl = t.GY * G::TYO;
        t.RW = S * t.TWl + G::HALO; t.RH = S * t.THl + G::HALO;
        t.NB = ((t.RW + 1) / 2 + 7) / 8;
        // a thread's last pair is 8 (GX - 1) + NPX - 1: its block row (NPX - 1) >> 3 must exist
        if (t.NB < t.GX + ((G::NPX - 1) >> 3)) t.NB = t.GX + ((G::NPX - 1) >> 3);
        // narrow groups put several thread rows into one half-warp: two extra floats per row keep them on different banks
        t.pitchF = 16 * t.NB + (t.GX < 16 ? 2 : 0);
        t.groupFloats = t.RH * t.pitchF + G::KW * t.KWp;
        t.gthreads = (t.GX * t.GY + 31) / 32 * 32;
    };
    // the tallest tile that fits; then the groups per CTA that keep the most threads resident (ties: larger CTAs).
    // Shorter tiles for more CTAs were measured slower (partial tiles, idle thread rows, more halo rows per output row).
    long long best = -1;
    int best_gy = 1, best_groups = 1;
    for (int gy = gy_max; gy >= 1 && best < 0; --gy) {
        geometry(gy);
        const long long gbytes = (long long)t.groupFloats * 4;
        for (int groups = G::THREADS / t.gthreads; groups >= 1; --groups) {
            const long long bytes = groups * gbytes;
            if (bytes > max_smem) continue;
            long long ctas = sm_smem / (bytes + 1024);
            const long long by_threads = 320 / (groups * t.gthreads);         // register file: up to 196 registers per thread
            if (ctas > by_threads) ctas = by_threads;
            if (ctas > 16) ctas = 16;
            const long long resident = ctas * groups * t.gthreads;
            if (resident > best) { best = resident; best_gy = gy; best_groups = groups; }
        }
    }
    KMSR_REQUIRE(best > 0, KMSR_E_UNSUPPORTED, "degrade (reg): k=%d factor=%d W=%d does not fit shared memory", K, S, t.W);
    geometry(best_gy);
    t.groups = best_groups;
    const size_t smem = (size_t)t.groups * t.groupFloats * 4;
    t.tiles_x = (t.Wo + t.TWl - 1) / t.TWl;
    t.tiles = t.tiles_x * ((t.Ho + t.THl - 1) / t.THl);
    t.nitems = a.N * a.C * t.tiles;
    const long long grid = (t.nitems + t.groups - 1) / t.groups;
    KMSR_REQUIRE(grid < (1ll << 31), KMSR_E_INVALID, "degrade (reg): too many tiles");
    auto kern = degrade_reg_kernel<K, S, TYO_, THREADS_>;
    KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)grid, t.groups * t.gthreads, smem, st>>>(t);
    KMSR_LAUNCH_CHECK("degrade_reg_kernel");
    return KMSR_OK;
}

template <int K>
int launch_reg_k(const DegradeArgs& a, RegArgs& t, int S, cudaStream_t st) {
    // factor 4 with 4 x 4 outputs per thread and 64-thread CTAs (the same staged tile per half as many threads) was
    // measured 15-30 % slower than 4 x 2 / 128 threads (r62): four warps per SM do not hide the shared-memory latency
    // hoisting all composite-kernel rows of an input row ahead of the FFMA2 blocks (k <= 15: 64 more registers, 245-255
    // in all) measured 3-20 % slower (r74).
    // factor 2 with 8 x 2 outputs per thread and 256-thread CTAs (twice the warps per staged tile) measured no faster at
    // P = 256 and 15-20 % slower at P = 128 (r69); tap-major FFMA2 ordering is what the compiler already emits (r68).
    // factor 8 (2 x 1 outputs per thread) was measured at half the streaming kernel's speed on every cell, 64-wide
    // patches included (r67): every staged pixel is re-read 4.4 times from shared memory there
    if (S == 2) return launch_reg<K, 2, 4, 128>(a, t, st);
    return launch_reg<K, 4, 2, 128>(a, t, st);
}

}  // namespace

bool reg_shape_ok(const DegradeArgs& a, int down_mode, const char** why) {
    const Geometry& g = a.g;
    *why = "";
    if (down_mode != KMSR_DOWN_BOXMEAN) { *why = "box-mean downsampling only"; return false; }
    const int k = g.kh;
    if (g.kh != g.kw || !(k == 11 || k == 13 || k == 15 || k == 21 || k == 31)) { *why = "square kernels of size 11, 13, 15, 21, 31"; return false; }
    if (g.stride != 2 && g.stride != 4) { *why = "factor 2 or 4"; return false; }
    if (a.patch_offsets) { *why = "scene windows not covered"; return false; }
    if (a.N * a.C >= (1ll << 40)) { *why = "too many patches"; return false; }
    return true;
}

int launch_degrade_reg(const DegradeArgs& a, cudaStream_t st) {
    const Geometry& g = a.g;
    RegArgs t;
    t.hr = a.hr; t.sN = a.N > 1 ? a.sN : 0; t.sC = a.sC; t.sH = a.sH;
    t.comp = a.comp; t.KWp = g.KWp; t.dsum = a.dsum; t.kidx = a.kidx; t.sigma = a.sigma; t.pool = a.pool; t.nidx = a.nidx;
    t.lr = a.lr; t.C = a.C; t.H = a.H; t.W = a.W; t.Ho = g.Ho; t.Wo = g.Wo;
    t.pad_mode = a.pad_mode; t.noise_mode = a.noise_mode;
    set_algo("reg");
    switch (g.kh) {
        case 11: return launch_reg_k<11>(a, t, g.stride, st);
        case 13: return launch_reg_k<13>(a, t, g.stride, st);
        case 15: return launch_reg_k<15>(a, t, g.stride, st);
        case 21: return launch_reg_k<21>(a, t, g.stride, st);
        default: return launch_reg_k<31>(a, t, g.stride, st);
    }
}

}  // namespace kmsr
