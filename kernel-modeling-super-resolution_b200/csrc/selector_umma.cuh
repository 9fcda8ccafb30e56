// selector_umma.cuh -- SelectorNet convolutions (muti_kernel/train_gemini.py:14-39) on the 5th-generation tensor cores:
// `tcgen05.mma.kind::tf32` with the accumulators AND the A operand in tensor memory, SURVEY.md 8f row f2, round 2.
// DESIGN.md 4.8 has the design, the probes behind it and the measurements.
//
// A 3x3 / stride-2 / pad-1 convolution is the GEMM  D[pixel, cout] = sum_k A[pixel, k] B[cout, k],  k = (tap, cin).
// One MMA tile is 128 output pixels (raster order inside a patch) x COUT channels, M = 128, K = 8 per instruction.
// fp32-level accuracy comes from the 3xTF32 split a_hi b_hi + a_hi b_lo + a_lo b_hi.  The tensor core truncates an fp32
// operand to TF32 (scratch/umma_round_probe.cu), so the raw activations ARE A_hi and A_lo = x - trunc_tf32(x) (exact).
// Builder threads (one row of one tile each) write 16 hi and 16 lo values per K = 16 sub-stage into a short ring of A
// slots in tensor memory (tcgen05.st); the MMAs are TS-form (A from tensor memory, the weights [B_hi ; B_lo] from shared
// memory in the K-major canonical layout without swizzle, host-built, streamed per stage with one bulk copy).
// Layers 2 / 3 (channel-last activations): the builders read TMA boxes -- 32 channels x Wo pixels (every second column)
// x 128 / Wo rows (every second row), 128-byte swizzle, zero fill outside the image = the padding.  First layer
// ([N, 5, H, W]): the input rows of a pass arrive as one TMA box, the builders gather their taps from it.
//
// Warp roles (480 threads, one persistent CTA per SM, a contiguous run of passes each):
//   warps 0-3   epilogue: tcgen05.ld of the tile (TMEM lane quadrant = warp), + bias, ReLU, then either the tile staged
//               in swizzled shared memory and written by TMA stores (two staging buffers in rotation) or -- last layer --
//               the per-tile channel sums (shuffle reduce-scatter, fixed order, deterministic) that pool_fc_kernel turns
//               into logits
//   warps 4, 14 MMA issue, one warp per M tile of the pass, warp-uniform code (elect.sync once): UTCHMMA back to back,
//               tcgen05.commit releases ring slots / publishes the accumulator buffer
//   warp  5     producer: a lane per TMA request of a stage (boxes, weight stage; first layer: input rows per pass)
//   warps 6-13  A-operand builders
// A pass = two M tiles sharing every weight stage; accumulators are double-buffered where 512 columns allow it.
#pragma once
#include <string.h>

#include "common.cuh"
#include "tma_util.cuh"

#ifndef UMMA_DBG
#define UMMA_DBG 0      // measurement builds only (scratch/conv_umma_test.cu): 1 no MMA issue, 2 no operand build, 4 no weight copy
#endif

namespace kmsr {
namespace umma {

constexpr int kEpiWarps = 4, kBuildWarps = 8;
constexpr int kBuildThreads = 32 * kBuildWarps;
constexpr int kMmaWarp = kEpiWarps, kTmaWarp = kEpiWarps + 1;
constexpr int kBuild0 = 32 * (kEpiWarps + 2);                  // first builder thread
constexpr int kMma2Warp = kEpiWarps + 2 + kBuildWarps;         // second MMA issuer (the pass's second M tile)
constexpr int kThreads = 32 * (kMma2Warp + 1);                 // 480
constexpr int kTPP = 2;                                        // M tiles per pass
constexpr uint32_t kTileBytes = 8192;                          // one 128 x 16 fp32 operand tile: [4 chunks][16 groups][8 rows][16 B]
constexpr uint32_t kChunkBytes = 2048;                         // LBO of the A tiles (core matrices adjacent in K)

struct ConvUArgs {
    const float* in;        // first layer: [N, 5, H, W]; others: [N, H, W, CIN] (channel-last)
    const float* wst;       // [stage][4 chunks][2 COUT / 8][8][4]: shared-memory image of [B_hi ; B_lo] per stage (host-prepared)
    const float* bias;      // [COUT]
    float* out;             // [N, Ho, Wo, COUT] channel-last, or nullptr when pooling
    float* pool_part;       // [N, tiles, COUT] per-tile channel sums (last layer), or nullptr
    int H, W, Ho, Wo;
    int tiles;              // Ho Wo / 128, even
    long long passes;       // N tiles / 2
};

// ReLU as torch evaluates it: a NaN stays a NaN (fmaxf would return 0)
__device__ __forceinline__ float relu_keep_nan(float v) { return v < 0.0f ? 0.0f : v; }
// shared-memory matrix descriptor, K-major, no swizzle: start address, LBO (K direction), SBO (M / N direction), version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
// the same with the A operand in tensor memory (lane = row, one 32-bit column per value of K)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc),
                 "r"(accumulate)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                 "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
                 : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(float (&v)[16], uint32_t taddr) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// one stage of the butterfly reduce-scatter over the lanes of a warp: 2 H live values per lane become H
template <int H>
__device__ __forceinline__ void rs_stage(float (&v)[32], int lane) {
    const bool up = (lane & H) != 0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float send = up ? v[j] : v[j + H];
        const float recv = __shfl_xor_sync(0xffffffffu, send, H);
        v[j] = (up ? v[j + H] : v[j]) + recv;
    }
}

template <int CIN, int COUT, bool POOL>
struct Shape {
    static constexpr int G = CIN >= 16 ? CIN / 16 : 1;                 // channel groups of 16 per tap
    static constexpr int SUB = CIN == 5 ? 1 : 2;                       // K = 16 sub-stages per stage: the issuing thread's per-stage cost
                                                                       // (two barrier waits, two commits) is paid once per 12 MMAs, not per 6
    static constexpr int S = CIN == 5 ? 3 : 9 * G / SUB;               // stages per pass (first layer: K = 45 -> 48)
    static constexpr uint32_t B_LBO = 32u * COUT;                      // (2 COUT / 8) row groups x 128 B
    static constexpr uint32_t B_BYTES = 4 * B_LBO;
    // TS form: the A operand lives in TENSOR memory, a short ring of NL slots of 64 columns (per M tile 16 hi + 16 lo
    // columns, one 32-bit column per value of K) written by the builder warps; the MMAs read only the weights from shared
    // memory.  Layers 2 / 3: a ring of TMA landing slots [hi 0 | hi 1 | B] in shared memory feeds the builders; first
    // layer: the builders gather from global memory and a slot holds the stage's weights only.
    // CAT (two-MMA form, D = 2 COUT columns per tile: hi x [B_hi ; B_lo], lo x B_hi) where tensor memory has the room,
    // otherwise three MMAs of N = COUT per k-step (hi x B_hi, hi x B_lo, lo x B_hi) into COUT columns.
    static constexpr bool GATHER = CIN == 5;
    static constexpr uint32_t BOFF = GATHER ? 0u : SUB * kTPP * kTileBytes; // weights behind the tiles [sub-stage][tile]
    static constexpr uint32_t SLOT = BOFF + SUB * B_BYTES;
    static constexpr int NH = GATHER ? 8 : (COUT <= 64 ? 4 : 3);
    static constexpr int NL = GATHER ? 4 : 2;                               // A-operand slots in tensor memory
    static constexpr int ACOLS = 64 * SUB;                                  // columns per A slot: [tile][sub-stage][hi 16 | lo 16]
    static constexpr bool CAT = GATHER;
    static constexpr int DCOLS = CAT ? 2 * COUT : COUT;                     // accumulator columns per tile
    static constexpr int TCOLS = DCOLS * kTPP;                              // ... per buffer
    static constexpr int NBUF = (512 - NL * ACOLS) / TCOLS >= 2 ? 2 : 1;    // accumulator buffers
    static constexpr int ABASE = NBUF * TCOLS;                              // first column of the A ring
    static constexpr uint32_t RAW_SLOT = GATHER ? 26u * 1024u : 0u;         // first layer: the input rows of one pass (W <= 256: 5 bands x 5 rows)
    static constexpr int NR = GATHER ? 5 : 0;                               // ... and how many passes are in flight
    static constexpr uint32_t OUT_BYTES = POOL ? 0u : 2u * 16384u;                 // two staging buffers of 32 channels x 128 pixels for the TMA stores
    static constexpr size_t SMEM = (size_t)NH * SLOT + (size_t)NR * RAW_SLOT + OUT_BYTES + 512 /* barriers */ + COUT * 4 + (POOL ? 4 * 128 * 4 : 0) + 1024 /* alignment */;
};

template <int CIN, int COUT, bool POOL>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap omap, const ConvUArgs a) {
    using Sh = Shape<CIN, COUT, POOL>;
    constexpr bool TMA = CIN != 5;                                     // channel-last input: the hi operand is a TMA box
    constexpr int S = Sh::S, G = Sh::G;
    extern __shared__ uint8_t umma_smem_raw[];
    const uint32_t raw = smem_u32(umma_smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* basep = umma_smem_raw + (base - raw);
    constexpr int NH = Sh::NH, NL = Sh::NL > 0 ? Sh::NL : 1;
    const uint32_t rawbase = base + NH * Sh::SLOT;                     // first layer: ring of staged input rows
    const uint32_t ostage = rawbase + Sh::NR * Sh::RAW_SLOT;           // 2 x [128 rows][128 B], 128-byte swizzle (store epilogue)
    const uint32_t bars = ostage + Sh::OUT_BYTES;                      // full[4] empty[4] tfull[2] tempty[2] | tmem holder | lofull[4]
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 64u + 8u * s; };      // 8 barriers each: full, empty, lofull, loempty, tfull | tempty (4 + 4)
    auto lofull_bar = [&](int s) { return bars + 128u + 8u * s; };
    auto loempty_bar = [&](int s) { return bars + 192u + 8u * s; };
    auto tfull_bar = [&](int b) { return bars + 256u + 8u * b; };
    auto tempty_bar = [&](int b) { return bars + 288u + 8u * b; };
    auto rawfull_bar = [&](int s) { return bars + 352u + 8u * s; };
    auto rawempty_bar = [&](int s) { return bars + 416u + 8u * s; };
    volatile uint32_t* tmem_holder = reinterpret_cast<volatile uint32_t*>(basep + (ostage - base) + Sh::OUT_BYTES + 320);
    float* bias_s = reinterpret_cast<float*>(basep + (ostage - base) + Sh::OUT_BYTES + 512);
    float* red_s = bias_s + COUT;                                      // [4 warps][128] (pooling only)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < NH; ++s) {
            mbar_init(full_bar(s), 1);                                 // the producer's expect_tx
            mbar_init(empty_bar(s), kTPP);                             // one tcgen05.commit per MMA issuer
        }
        for (int s = 0; s < NL; ++s) {
            mbar_init(lofull_bar(s), kBuildThreads);                   // every builder after its lo stores
            mbar_init(loempty_bar(s), kTPP);
        }
        for (int s = 0; s < Sh::NR; ++s) {
            mbar_init(rawfull_bar(s), 1);
            mbar_init(rawempty_bar(s), kBuildThreads);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), kTPP);
            mbar_init(tempty_bar(b), 32 * kEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int c = tid; c < COUT; c += kThreads) bias_s[c] = a.bias[c];
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_holder)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_holder;
    const int half_tiles = a.tiles / kTPP;
    // a CTA takes a contiguous run of passes: consecutive passes are neighbouring rows of one patch, so the rows two tiles
    // share and the taps' re-reads stay hot in L2 (a grid-strided assignment jumped 2.3 patches per pass: the first-layer
    // gather waited a third of its time on DRAM / TLB misses)
    const long long p_begin = a.passes * (long long)blockIdx.x / gridDim.x, p_end = a.passes * (long long)(blockIdx.x + 1) / gridDim.x;

    if (warp < kEpiWarps) {
        // ------------------------------------------------------------------------------------------ epilogue
        const int row = tid;                                           // row of the M tile = TMEM lane
        const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
        uint32_t pc = 0, unit = 0;
        // 16 accumulator columns of this thread's row (SS form: the hi*hi + lo*hi and the hi*lo halves added)
        auto acc16 = [&](float (&p)[16], uint32_t col) {
            tmem_ld16(p, col);
            if constexpr (Sh::CAT) {
                float q[16];
                tmem_ld16(q, col + COUT);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) p[j] += q[j];
            } else {
                tmem_ld_wait();
            }
        };
        for (long long pass = p_begin; pass < p_end; ++pass, ++pc) {
            const int buf = (int)(pc % Sh::NBUF);
            const long long n = pass / half_tiles;
            const int tile0 = (int)(pass - n * half_tiles) * kTPP;
            mbar_wait_relaxed(tfull_bar(buf), (pc / Sh::NBUF) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < kTPP; ++t) {
                const uint32_t tcol = lane_addr + (uint32_t)(buf * Sh::TCOLS + t * Sh::DCOLS);
                if constexpr (!POOL) {
                    // the tile's output is 128 x COUT contiguous floats: staged in shared memory (row = pixel, 128-byte rows of
                    // 32 channels, 16-byte chunks XOR-swizzled with the row so the per-row stores are conflict free) and written by
                    // one TMA store per 32 channels, which undoes the swizzle -- per-thread 16-byte global stores 128 B apart cost
                    // 2.3 ms per 4096 patches on the first layer
                    // two staging buffers of 32 channels x 128 pixels in rotation, one bulk group per buffer: before a
                    // buffer is rewritten only the store before the last one has to have read it
                    const int pix0 = (int)((n * a.tiles + tile0 + t) * 128);
#pragma unroll 1
                    for (int h = 0; h < COUT / 32; ++h, ++unit) {
                        const uint32_t obuf = ostage + (unit & 1u) * 16384u;
                        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                        for (int j0 = 0; j0 < 32; j0 += 16) {
                            float p[16];
                            acc16(p, tcol + 32 * h + j0);
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const int c = 32 * h + j0 + j;                           // first channel of this 16-byte chunk
                                const uint32_t o0 = __float_as_uint(relu_keep_nan(p[j] + bias_s[c]));
                                const uint32_t o1 = __float_as_uint(relu_keep_nan(p[j + 1] + bias_s[c + 1]));
                                const uint32_t o2 = __float_as_uint(relu_keep_nan(p[j + 2] + bias_s[c + 2]));
                                const uint32_t o3 = __float_as_uint(relu_keep_nan(p[j + 3] + bias_s[c + 3]));
                                const uint32_t addr = obuf + (uint32_t)row * 128u + ((((uint32_t)((j0 + j) >> 2)) ^ ((uint32_t)row & 7u)) << 4);
                                if (!(UMMA_DBG & 8)) st_shared_v4(addr, o0, o1, o2, o3);
                            }
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (tid == 0) {
                            if (!(UMMA_DBG & 8))
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&omap), "r"(32 * h),
                                             "r"(pix0), "r"(obuf)
                                             : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                    }
                } else {
#pragma unroll 1
                    for (int j0 = 0; j0 < COUT; j0 += 32) {
                        float v[32];
                        {
                            float p[16];
                            acc16(p, tcol + j0);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = relu_keep_nan(p[j] + bias_s[j0 + j]);
                            acc16(p, tcol + j0 + 16);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[16 + j] = relu_keep_nan(p[j] + bias_s[j0 + 16 + j]);
                        }
                        rs_stage<16>(v, lane);
                        rs_stage<8>(v, lane);
                        rs_stage<4>(v, lane);
                        rs_stage<2>(v, lane);
                        rs_stage<1>(v, lane);
                        red_s[warp * 128 + j0 + lane] = v[0];         // column j0 + lane, summed over this warp's 32 pixels
                    }
                    if (t == kTPP - 1) {                                 // every TMEM read of the pass is done: hand the buffer back
                        tc_fence_before();
                        mbar_arrive(tempty_bar(buf));
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int c = tid; c < COUT; c += 128)
                        a.pool_part[((long long)n * a.tiles + tile0 + t) * COUT + c] =
                            ((red_s[c] + red_s[128 + c]) + red_s[256 + c]) + red_s[384 + c];
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
            }
            if constexpr (!POOL) {
                tc_fence_before();
                mbar_arrive(tempty_bar(buf));
            }
        }
        if (!POOL && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else if (warp == kMmaWarp || warp == kMma2Warp) {
        // ------------------------------------------------------------------------------------------ MMA issue
        // One thread per M tile of the pass: the MMAs here are small (32-128 clocks of tensor time each) and a thread needs
        // ~50-90 clocks to issue one (uniform-register descriptor moves, elect, branch), so a single issuer was the limiter
        // (ncu: 80 % of its samples in the issue sequence).  Descriptors are the slot-0 descriptor plus constants.
        uint32_t leader;
        asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(leader));
        {
            const int t = warp == kMmaWarp ? 0 : 1;
            constexpr uint32_t idesc_cat = idesc_tf32(2 * COUT), idesc_one = idesc_tf32(COUT);
            constexpr uint64_t kJB = (2 * Sh::B_LBO) >> 4;                                      // second k-step of a stage, B
            constexpr uint64_t kBlo = (16u * COUT) >> 4;                                        // rows COUT.. of the weight image: the lo part
            const uint64_t b0 = smem_desc(base + Sh::BOFF, Sh::B_LBO, 128);
            uint32_t it = 0, pc = 0;
            for (long long pass = p_begin; pass < p_end; ++pass, ++pc) {
                const int buf = (int)(pc % Sh::NBUF);
                mbar_wait(tempty_bar(buf), ((pc / Sh::NBUF) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t dcol = tm + (uint32_t)(buf * Sh::TCOLS + t * Sh::DCOLS);
#pragma unroll 1
                for (int s = 0; s < S; ++s, ++it) {
                    const int slot = (int)(it % NH), ls = (int)(it % NL);
                    mbar_wait(full_bar(slot), (it / NH) & 1u);
                    mbar_wait(lofull_bar(ls), (it / NL) & 1u);
                    tc_fence_after();
                    const uint64_t bd = b0 + (uint64_t)((slot * Sh::SLOT) >> 4);
                    if (!(UMMA_DBG & 1) && leader) {
                        // A from tensor memory: columns [hi 16 | lo 16] per sub-stage of this tile in slot ls, 8 per k-step
#pragma unroll
                        for (int h = 0; h < Sh::SUB; ++h) {
                            const uint32_t ac = tm + (uint32_t)(Sh::ABASE + ls * Sh::ACOLS + t * 32 * Sh::SUB + h * 32);
                            const uint64_t bh = bd + (uint64_t)((h * Sh::B_BYTES) >> 4);
                            if constexpr (Sh::CAT) {
                                mma_tf32_ts(dcol, ac, bh, idesc_cat, (s | h) != 0);
                                mma_tf32_ts(dcol, ac + 16, bh, idesc_one, 1u);
                                mma_tf32_ts(dcol, ac + 8, bh + kJB, idesc_cat, 1u);
                                mma_tf32_ts(dcol, ac + 24, bh + kJB, idesc_one, 1u);
                            } else {
                                mma_tf32_ts(dcol, ac, bh, idesc_one, (s | h) != 0);
                                mma_tf32_ts(dcol, ac, bh + kBlo, idesc_one, 1u);
                                mma_tf32_ts(dcol, ac + 16, bh, idesc_one, 1u);
                                mma_tf32_ts(dcol, ac + 8, bh + kJB, idesc_one, 1u);
                                mma_tf32_ts(dcol, ac + 8, bh + kJB + kBlo, idesc_one, 1u);
                                mma_tf32_ts(dcol, ac + 24, bh + kJB, idesc_one, 1u);
                            }
                        }
                    }
                    if (leader) {
                        mma_commit(empty_bar(slot));                   // arrives once these MMAs have read the slot's weights
                        mma_commit(loempty_bar(ls));                   // ... and the A slot
                    }
                    __syncwarp();
                }
                if (leader) mma_commit(tfull_bar(buf));                // ... and once the accumulators are final
            }
        }
    } else if (warp == kTmaWarp) {
        // ------------------------------------------------------------------------------------------ TMA producer (layers 2 / 3)
        if constexpr (!TMA) {
            // first layer: lane 0 fetches the input rows of a pass (all 5 bands, the 4 (128 / Wo) + 1 rows its two tiles touch,
            // full width; rows above the image are zero-filled = the padding) as one TMA box per pass, Sh::NR passes ahead of
            // the gather; lane 2 streams the weight stages
            if (lane == 0) {
                const int rpt = 128 / a.Wo;
                const uint32_t bytes = (uint32_t)(a.W * (4 * rpt + 1) * CIN * 4);
                uint32_t pc = 0;
                for (long long pass = p_begin; pass < p_end; ++pass, ++pc) {
                    const long long n = pass / half_tiles;
                    const int tile0 = (int)(pass - n * half_tiles) * kTPP;
                    const int rs = (int)(pc % Sh::NR);
                    mbar_wait(rawempty_bar(rs), ((pc / Sh::NR) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(rawfull_bar(rs), bytes);
                    tma_load_4d(rawbase + rs * Sh::RAW_SLOT, &tmap, 0, 2 * tile0 * rpt - 1, 0, (int)n, rawfull_bar(rs));
                }
            } else if (lane == 2) {
                uint32_t it = 0;
                for (long long pass = p_begin; pass < p_end; ++pass) {
#pragma unroll 1
                    for (int s = 0; s < S; ++s, ++it) {
                        const int slot = (int)(it % NH);
                        mbar_wait(empty_bar(slot), ((it / NH) & 1u) ^ 1u);
                        if (UMMA_DBG & 4) mbar_arrive(full_bar(slot));
                        else {
                            mbar_arrive_expect_tx(full_bar(slot), Sh::B_BYTES);
                            bulk_g2s(base + slot * Sh::SLOT + Sh::BOFF, a.wst + (size_t)s * (Sh::B_BYTES / 4), Sh::B_BYTES, full_bar(slot));
                        }
                    }
                }
            }
        } else {
            // one lane per request of a stage (a single thread sustains about one TMA request per 280 ns): lanes
            // 0 / 1 the two tiles' boxes, lane 2 the weight stages; lane 0 also posts the byte count.  A box row is the 32
            // channels of the stage (128 bytes, 128-byte swizzle): the TMA engine needs ~2.7 clocks per box row whatever its
            // length, so 64-byte rows made the feed twice as expensive
            constexpr int NBOX = kTPP, GS = G / Sh::SUB;
            static_assert(Sh::SUB == 2, "a box row is two K = 16 sub-stages");
            if (lane < NBOX + 1) {
                const int rows_per_tile = 128 / a.Wo;                  // output rows of one M tile
                const int bt = lane;                                   // this lane's box: tile
                uint32_t it = 0;
                for (long long pass = p_begin; pass < p_end; ++pass) {
                    const long long n = pass / half_tiles;
                    const int tile0 = (int)(pass - n * half_tiles) * kTPP;
#pragma unroll 1
                    for (int s = 0; s < S; ++s, ++it) {
                        const int slot = (int)(it % NH);
                        const uint32_t sa = base + slot * Sh::SLOT;
                        const int tap = s / GS, g = s - tap * GS;             // 32-channel group of the stage
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        mbar_wait(empty_bar(slot), ((it / NH) & 1u) ^ 1u);
                        if (lane == 0)
                            mbar_arrive_expect_tx(full_bar(slot), NBOX * Sh::SUB * kTileBytes + ((UMMA_DBG & 4) ? 0u : Sh::SUB * Sh::B_BYTES));
                        // the raw fp32 activations ARE the hi operand (the tensor core reads the upper 19 bits): box = 32 channels
                        // x Wo pixels at stride 2 x (128 / Wo) rows at stride 2, zero-filled outside the image = the padding
                        if (lane < NBOX)
                            tma_load_5d(sa + bt * Sh::SUB * kTileBytes, &tmap, 32 * g, (dx + 1) & 1, (dx + 1) / 2 - 1,
                                        2 * (tile0 + bt) * rows_per_tile + dy - 1, (int)n, full_bar(slot));
                        else if (!(UMMA_DBG & 4))
                            bulk_g2s(sa + Sh::BOFF, a.wst + (size_t)s * (Sh::SUB * Sh::B_BYTES / 4), Sh::SUB * Sh::B_BYTES, full_bar(slot));
                    }
                }
            }
        }
    } else if constexpr (TMA) {
        // ------------------------------------------------------------------------------------------ A-operand builders (layers 2 / 3)
        // A thread owns one row (output pixel) of one M tile: warp % 4 is the TMEM lane quadrant it may write, the first four
        // builder warps take tile 0, the other four tile 1.  Per stage it reads its 128-byte row of the landed (128-byte
        // swizzled) tile -- chunk c at position c ^ (row & 7): conflict free over a quarter-warp -- and writes 16 hi
        // and 16 lo values to tensor memory.  hi = x as it is (the MMA truncates it to TF32); lo = x - trunc(x) is exact in
        // fp32 and is stored as it is: the MMA truncates it too, an error of at most 2^-21 |x|, below the lo * lo term the
        // split drops anyway (rounding it here cost two more integer instructions per value and 3 % of the layer).
        const int q = warp & 3, t = (warp - kBuild0 / 32) >> 2;
        const int r = 32 * q + lane;
        const uint32_t rowoff = (uint32_t)t * Sh::SUB * kTileBytes + (uint32_t)r * 128u, sw = (uint32_t)r & 7u;
        const uint32_t tdst = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(Sh::ABASE + t * 32 * Sh::SUB);
        const long long my_passes = p_end - p_begin;
        const long long total = my_passes * S;
        uint32_t it = 0;
        for (long long qq = 0; qq < total; ++qq, ++it) {
            const int slot = (int)(it % NH), ls = (int)(it % NL);
            const uint32_t sa = base + slot * Sh::SLOT + rowoff;
            mbar_wait_relaxed(full_bar(slot), (it / NH) & 1u);
            mbar_wait_relaxed(loempty_bar(ls), ((it / NL) & 1u) ^ 1u);
            tc_fence_after();
            if (!(UMMA_DBG & 2)) {
#pragma unroll
                for (int h = 0; h < Sh::SUB; ++h) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                     : "=r"(hi[4 * c]), "=r"(hi[4 * c + 1]), "=r"(hi[4 * c + 2]), "=r"(hi[4 * c + 3])
                                     : "r"(sa + (((uint32_t)(4 * h + c) ^ sw) << 4)));
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float l = __uint_as_float(hi[k]) - __uint_as_float(hi[k] & 0xFFFFE000u);
                        lo[k] = __float_as_uint(l);
                    }
                    tmem_st16(tdst + (uint32_t)(ls * Sh::ACOLS + h * 32), hi);
                    tmem_st16(tdst + (uint32_t)(ls * Sh::ACOLS + h * 32) + 16, lo);
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            mbar_arrive(lofull_bar(ls));
        }
    } else {
        // ------------------------------------------------------------------------------------------ A-operand builders (first layer)
        // The 5-band [N, 5, H, W] input cannot be fetched as K-major boxes (a 16-byte chunk of the operand would be four taps
        // of one pixel), so threads gather it -- from the input rows the producer staged in shared memory (a gather from
        // global memory waited a third of its time on load latency even with the loads issued a pass ahead) straight into
        // tensor memory: a thread owns one row (output pixel) of one M tile (warp % 4 = TMEM lane quadrant, first four
        // builder warps tile 0, the others tile 1) and writes 16 hi and 16 lo values per stage with two tcgen05.st.
        // K order: stage s holds the five (band, row) triples T = 5 s + j, j < 5 (band = T / 3, dy = T % 3), k = 3 j + dx,
        // k = 15 is zero.  Per triple a lane reads the aligned pair (2 ox, 2 ox + 1) -- conflict free: a warp reads 256
        // contiguous bytes -- and takes column 2 ox - 1 from its left neighbour's pair by shuffle (lane 0 reads it itself).
        const int q = warp & 3, t = (warp - kBuild0 / 32) >> 2;
        const int r = 32 * q + lane;
        const uint32_t tdst = tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(Sh::ABASE + t * 32);
        const int RB = 4 * (128 / a.Wo) + 1;                            // staged rows per band
        const int Pl = t * 128 + r, oyl = Pl / a.Wo, ox = Pl - oyl * a.Wo;       // this thread's pixel inside any pass
        const uint32_t poff = (uint32_t)((2 * oyl * a.W + 2 * ox) * 4);          // staged row 2 oyl, column 2 ox of band 0
        const bool left_ok = lane == 0 && ox != 0;
        uint32_t it = 0, pc = 0;
        for (long long pass = p_begin; pass < p_end; ++pass, ++pc) {
            const int rs = (int)(pc % Sh::NR);
            const uint32_t rawp = rawbase + rs * Sh::RAW_SLOT + poff;
            mbar_wait_relaxed(rawfull_bar(rs), (pc / Sh::NR) & 1u);
            float2 pr[3][5];
            float ex[3][5];
            if (!(UMMA_DBG & 2)) {
#pragma unroll
                for (int u = 0; u < 3; ++u)
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const int T = 5 * u + j, ch = T / 3, dy = T - 3 * ch;
                        const uint32_t ad = rawp + (uint32_t)(((ch * RB + dy) * a.W) * 4);
                        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pr[u][j].x), "=f"(pr[u][j].y) : "r"(ad));
                        ex[u][j] = 0.0f;
                        if (left_ok) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(ex[u][j]) : "r"(ad - 4));
                    }
            }
            // mbarrier.arrive does not wait for the thread's outstanding LDS (load-completion fence): a CTA-scope
            // fence between the loads and the release of the slot
            __threadfence_block();
            mbar_arrive(rawempty_bar(rs));
#pragma unroll
            for (int u = 0; u < 3; ++u, ++it) {
                const int ls = (int)(it % NL);
                uint32_t hi[16], lo[16];
                if (!(UMMA_DBG & 2)) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const float left = __shfl_up_sync(0xffffffffu, pr[u][j].y, 1);
                        hi[3 * j] = __float_as_uint(lane == 0 ? ex[u][j] : left);
                        hi[3 * j + 1] = __float_as_uint(pr[u][j].x);
                        hi[3 * j + 2] = __float_as_uint(pr[u][j].y);
                    }
                    hi[15] = 0u;
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const float l = __uint_as_float(hi[k]) - __uint_as_float(hi[k] & 0xFFFFE000u);
                        lo[k] = __float_as_uint(l);
                    }
                }
                mbar_wait_relaxed(loempty_bar(ls), ((it / NL) & 1u) ^ 1u);
                tc_fence_after();
                if (!(UMMA_DBG & 2)) {
                    tmem_st16(tdst + (uint32_t)(ls * Sh::ACOLS), hi);
                    tmem_st16(tdst + (uint32_t)(ls * Sh::ACOLS) + 16, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                tc_fence_before();
                mbar_arrive(lofull_bar(ls));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

template <int CIN, int COUT, bool POOL>
int launch_conv_umma(const ConvUArgs& a, long long N, int sms, cudaStream_t st) {
    using Sh = Shape<CIN, COUT, POOL>;
    auto kern = conv_umma_kernel<CIN, COUT, POOL>;
    KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Sh::SMEM));
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof tmap);
    if (CIN == 5) {
        // [N, 5, H, W] seen as (x, y, band, n); one box = the rows of one pass: full width x 4 (128 / Wo) + 1 rows x 5 bands
        EncodeTiledFn enc = get_tensor_map_encoder();
        KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled is not available");
        KMSR_REQUIRE(a.W <= 256 && (size_t)a.W * (4 * (128 / a.Wo) + 1) * 5 * 4 <= Sh::RAW_SLOT, KMSR_E_UNSUPPORTED,
                     "selector (tcgen05): first-layer rows of %d pixels do not fit the staging slot", a.W);
        cuuint64_t gdim[4] = {(cuuint64_t)a.W, (cuuint64_t)a.H, 5, (cuuint64_t)N};
        cuuint64_t gstr[3] = {(cuuint64_t)a.W * 4, (cuuint64_t)a.H * a.W * 4, (cuuint64_t)a.H * a.W * 20};
        cuuint32_t box[4] = {(cuuint32_t)a.W, (cuuint32_t)(4 * (128 / a.Wo) + 1), 5, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)a.in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled (input rows) failed with CUresult %d", (int)cr);
    } else {
        // channel-last activations [N, H, W, CIN] seen as (c, x parity, x / 2, y, n); one box = one 128-pixel M tile of one
        // 32-channel group of one tap: 32 channels x Wo pixels (one column parity) x 128 / Wo rows (every second row), 128-byte swizzle
        EncodeTiledFn enc = get_tensor_map_encoder();
        KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled is not available");
        cuuint64_t gdim[5] = {(cuuint64_t)CIN, 2, (cuuint64_t)(a.W / 2), (cuuint64_t)a.H, (cuuint64_t)N};
        cuuint64_t gstr[4] = {(cuuint64_t)CIN * 4, (cuuint64_t)CIN * 8, (cuuint64_t)a.W * CIN * 4, (cuuint64_t)a.H * a.W * CIN * 4};
        cuuint32_t box[5] = {32, 1, (cuuint32_t)a.Wo, (cuuint32_t)(2 * (128 / a.Wo)), 1};
        cuuint32_t estr[5] = {1, 1, 1, 2, 1};
        CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)a.in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    }
    CUtensorMap omap;
    memset(&omap, 0, sizeof omap);
    if (!POOL) {
        // channel-last output [N Ho Wo, COUT] seen as (c, pixel); one box = 32 channels of one 128-pixel tile, 128-byte swizzle
        EncodeTiledFn enc = get_tensor_map_encoder();
        KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled is not available");
        const long long pixels = N * a.Ho * a.Wo;
        KMSR_REQUIRE(pixels < (1ll << 31), KMSR_E_INVALID, "selector (tcgen05): %lld output pixels in one call", pixels);
        cuuint64_t gdim[2] = {(cuuint64_t)COUT, (cuuint64_t)pixels};
        cuuint64_t gstr[1] = {(cuuint64_t)COUT * 4};
        cuuint32_t box[2] = {32, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult cr = enc(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a.out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "selector (tcgen05): cuTensorMapEncodeTiled (output) failed with CUresult %d", (int)cr);
    }
    const unsigned grid = (unsigned)(a.passes < sms ? a.passes : sms);
    kern<<<grid, kThreads, Sh::SMEM, st>>>(tmap, omap, a);
    KMSR_LAUNCH_CHECK("conv_umma_kernel");
    return KMSR_OK;
}

}  // namespace umma
}  // namespace kmsr
