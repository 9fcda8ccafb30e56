// degrade_tma.cu -- TMA row-streaming fused blur + downsample + noise kernel for the headline shape
// (k = 13, factor 8, replicate padding, W = 256, H a multiple of 8; BASELINE configs 1-3).
//
// Same arithmetic as degrade_tiled.cu (C_30apply_kernel_to_landsat.py:68-124 with the box mean
// folded into a 20 x 20 stride-8 composite kernel, E_make_train_data.py:72-74 /
// train_gemini.py:137 noise in the epilogue); what changes is how the bytes move:
//
//  * persistent CTAs (one per SM), each running 4 independent band streams.  A stream walks a
//    256 x 256 band top to bottom in chunks of 8 rows; every HR byte crosses HBM -> SMEM exactly
//    once (no vertical halo re-read) through a ring of TMA tiles (cp.async.bulk.tensor, one
//    producer thread, full/empty mbarriers).  The tensor map describes [N, C, H, W/2] 64-bit
//    elements and the box is 138 x 8 starting at x = -4 (the byte offset of a box start must be a
//    multiple of 16: x = -3 raises an illegal-instruction fault, measured): TMA's out-of-bounds
//    zero fill lays each row down as [8 halo | 256 pixels | 12 pad] with a pitch of 276 floats =
//    69 x 16 B (odd), so a quarter-warp reading the same 16-byte column of 8 consecutive rows is
//    bank-conflict free.
//  * two warps per stream; lane = (ly, gx): ly = row residue mod 8, gx = group of 4 adjacent LR
//    columns.  A lane owns input rows r == ly (mod 8) and keeps the 2-3 composite-kernel rows that
//    can meet such a row (u = ly, ly+8, ly+16) in REGISTERS for the whole band: 60 weights, no
//    shared-memory weight traffic.  Per 8-row chunk a lane loads its 44-float row segment once
//    (12 LDS.128) and feeds 200 useful FMAs from it (3 output rows x 4 output columns x 20 taps).
//  * replicate padding never touches shared memory: rows clamp by address, the six halo columns of
//    the two edge groups are substituted in registers.
//  * the 8 row-residue partial sums of an output are combined with a 4-shuffle reduce-scatter and
//    written as 64-byte segments; the noise value is prefetched a whole chunk earlier.
//  * pixels are accumulated as (x - pivot), pivot = first pixel of the lane's column group
//    (SURVEY.md 7.3.2); pivot * sum(K') is added back once in the epilogue.
#include <cuda.h>

#include "common.cuh"

namespace kmsr {

namespace {

constexpr int kS = 8;                          // output stride (effective downscale factor)
constexpr int kK = 13;                         // blur kernel size
constexpr int kKW = kK + kS - 1;               // 20: composite window
constexpr int kPad = kK / 2;                   // 6
constexpr int kStreams = 4;                    // band streams per CTA
constexpr int kDepth = 5;                      // ring slots per stream
constexpr int kRowF = 276;                     // floats per staged row (138 x 8 B box)
constexpr int kChunkF = 8 * kRowF;             // floats per chunk
constexpr int kChunkBytes = kChunkF * 4;       // 8832 = 69 * 128
constexpr int kConsumerWarps = 2 * kStreams;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kSegF = 3 * kS + kKW;            // 44 floats: the row segment 4 adjacent outputs need
constexpr int kLeftF = 8;                      // staged halo columns left of pixel 0 (16-byte aligned box start)
constexpr int kSkew = kLeftF - kPad;           // 2: the segment starts 2 floats into its first 16-byte chunk
constexpr int kLoadF = (kSkew + kSegF + 3) / 4 * 4;   // 48 floats = 12 LDS.128
constexpr size_t kSmemBytes = (size_t)kStreams * kDepth * kChunkBytes + 2 * kStreams * kDepth * 8 + 128;

struct TmaArgs {
    const float* comp;
    const float* dsum;
    const int* kidx;
    const float* sigma;
    const float* pool;
    const int* nidx;
    float* lr;
    long long nbands;
    int C, H, Ho, Wo;
    int nchunks;      // H/8 + 1
    int noise_mode;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int x, int y, int z, int w,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(z), "r"(w), "r"(bar)
        : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
degrade_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kStreams * kDepth * kChunkBytes);
    // bars[s*kDepth + d] = full, bars[kStreams*kDepth + s*kDepth + d] = empty
    const uint32_t full0 = smem_u32(bars);
    const uint32_t empty0 = smem_u32(bars + kStreams * kDepth);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStreams * kDepth; ++i) {
            mbar_init(full0 + 8 * i, 1);        // the producer's arrive.expect_tx
            mbar_init(empty0 + 8 * i, 2);       // one arrive per consumer warp of the stream
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long G = (long long)gridDim.x * kStreams;

    if (warp == kConsumerWarps) {
        // ===================== TMA producer: one thread feeds the 4 rings =====================
        if (lane != 0) return;
        long long band[kStreams];
        int chunk[kStreams], slot[kStreams];
        uint32_t par[kStreams];
#pragma unroll
        for (int s = 0; s < kStreams; ++s) {
            band[s] = (long long)blockIdx.x * kStreams + s;
            chunk[s] = 0; slot[s] = 0; par[s] = 1;       // fresh barriers: waiting on parity 1 passes
        }
        bool active = true;
        while (active) {
            active = false;
#pragma unroll
            for (int s = 0; s < kStreams; ++s) {
                if (band[s] >= a.nbands) continue;
                active = true;
                const int b = s * kDepth + slot[s];
                mbar_wait(empty0 + 8 * b, par[s]);
                mbar_arrive_expect_tx(full0 + 8 * b, kChunkBytes);
                const long long n = band[s] / a.C;
                const int c = (int)(band[s] - n * a.C);
                tma_load_4d(smem_u32(ring + (size_t)b * kChunkF), &tmap, -(kLeftF / 2), kS * chunk[s] - kPad, c,
                            (int)n, full0 + 8 * b);
                if (++slot[s] == kDepth) { slot[s] = 0; par[s] ^= 1; }
                if (++chunk[s] == a.nchunks) { chunk[s] = 0; band[s] += G; }
            }
        }
        return;
    }

    // ============================== consumers: 2 warps per stream ==============================
    const int s = warp >> 1, half = warp & 1;
    const int ly = lane & 7, gx = lane >> 3;
    const int g = half * 4 + gx;                 // group of 4 output columns: X = 4g .. 4g+3
    const int ngroups = a.Wo >> 2;
    const bool left_edge = g == 0, right_edge = g == ngroups - 1;
    const float* sring = ring + (size_t)s * kDepth * kChunkF;
    const uint32_t sfull = full0 + 8 * s * kDepth, sempty = empty0 + 8 * s * kDepth;
    const int nsteps = a.nchunks + 1;
    const long long ohw = (long long)a.Ho * a.Wo;
    // after the reduce-scatter lane (ly) holds output column 4g + (ly >> 1); even ly writes
    const int ox_mine = ly >> 1;
    const bool writer = (ly & 1) == 0;

    int slot = 0;
    uint32_t par = 0;

    for (long long band = (long long)blockIdx.x * kStreams + s; band < a.nbands; band += G) {
        const long long n = band / a.C;
        const int c = (int)(band - n * a.C);
        const int kid = a.kidx ? __ldg(a.kidx + n) : 0;
        const float* kc = a.comp + ((long long)kid * a.C + c) * (kKW * kKW);

        // composite-kernel rows this lane can ever meet: u = ly, ly + 8, ly + 16 (< 20)
        float w[3][kKW];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const int u = ly + 8 * q;
#pragma unroll
            for (int v4 = 0; v4 < kKW / 4; ++v4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (u < kKW) t = __ldg(reinterpret_cast<const float4*>(kc + u * kKW) + v4);
                w[q][4 * v4 + 0] = t.x; w[q][4 * v4 + 1] = t.y; w[q][4 * v4 + 2] = t.z; w[q][4 * v4 + 3] = t.w;
            }
        }
        const float ds = __ldg(a.dsum + (long long)kid * a.C + c);
        float scale = 1.0f;
        const float* nz = nullptr;
        if (a.noise_mode != KMSR_NOISE_NONE) {
            nz = a.pool + ((long long)__ldg(a.nidx + n) * a.C + c) * ohw + 4 * g + ox_mine;
            if (a.noise_mode == KMSR_NOISE_SIGMA) scale = __ldg(a.sigma + (long long)kid * a.C + c);
        }
        float* out = a.lr + band * ohw + 4 * g + ox_mine;

        float acc[3][4];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int x = 0; x < 4; ++x) acc[q][x] = 0.0f;
        float pv = 0.0f;

#pragma unroll 1
        for (int i = 0; i < nsteps; ++i) {
            // chunk i holds padded rows 8i .. 8i+7 (image rows 8i-6 .. 8i+1); the extra last step
            // (bottom halo) re-reads the last chunk
            const bool fresh = i < a.nchunks;
            if (fresh) mbar_wait(sfull + 8 * slot, par);
            const int ci = fresh ? i : a.nchunks - 1;
            const int r = min(max(kS * i + ly - kPad, 0), a.H - 1);          // replicate: clamp by address
            const float* src = sring + (size_t)slot * kChunkF + (r + kPad - kS * ci) * kRowF + 32 * g;
            if (i == 0) {
                pv = sring[(size_t)slot * kChunkF + kPad * kRowF + 32 * g + kLeftF];   // pixel (0, 32g)
                if (!isfinite(pv)) pv = 0.0f;
            }
            float e[kLoadF];
#pragma unroll
            for (int j = 0; j < kLoadF / 4; ++j) {
                const float4 t = reinterpret_cast<const float4*>(src)[j];
                e[4 * j + 0] = t.x; e[4 * j + 1] = t.y; e[4 * j + 2] = t.z; e[4 * j + 3] = t.w;
            }
            float d[kSegF];                    // d[j] = pixel column 32g - 6 + j
#pragma unroll
            for (int j = 0; j < kSegF; ++j) d[j] = e[j + kSkew];
            // this chunk is no longer needed once its rows sit in registers -- except the last one,
            // which the bottom-halo step reads again
            const bool release = i < a.nchunks - 1;
            __syncwarp();
            if (release && lane == 0) mbar_arrive(sempty + 8 * slot);
            if (release) { if (++slot == kDepth) { slot = 0; par ^= 1; } }

            // noise for the output row that completes in this step (Y = i - 2): issue the load early
            const int Yd = i - 2;
            float nzv = 0.0f;
            if (nz && writer && Yd >= 0) nzv = __ldg(nz + (long long)Yd * a.Wo);

            if (left_edge) {
#pragma unroll
                for (int j = 0; j < kPad; ++j) d[j] = d[kPad];
            }
            if (right_edge) {
#pragma unroll
                for (int j = kSegF - kPad; j < kSegF; ++j) d[j] = d[kSegF - kPad - 1];
            }
#pragma unroll
            for (int j = 0; j < kSegF; ++j) d[j] -= pv;

            // row 8i+ly meets output row Y = i - q with composite row u = ly + 8q
            if (i < a.Ho) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int v = 0; v < kKW; ++v) acc[0][x] = fmaf(w[0][v], d[kS * x + v], acc[0][x]);
            }
            if (i >= 1 && i - 1 < a.Ho) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int v = 0; v < kKW; ++v) acc[1][x] = fmaf(w[1][v], d[kS * x + v], acc[1][x]);
            }
            if (i >= 2 && ly < kKW - 16) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int v = 0; v < kKW; ++v) acc[2][x] = fmaf(w[2][v], d[kS * x + v], acc[2][x]);
            }

            if (Yd >= 0) {
                // reduce-scatter over the 8 row residues (lanes differing in bits 0-2)
                const bool hi = (ly & 4) != 0;
                float k0 = hi ? acc[2][2] : acc[2][0], k1 = hi ? acc[2][3] : acc[2][1];
                const float s0 = hi ? acc[2][0] : acc[2][2], s1 = hi ? acc[2][1] : acc[2][3];
                k0 += __shfl_xor_sync(0xffffffffu, s0, 4);
                k1 += __shfl_xor_sync(0xffffffffu, s1, 4);
                const bool mid = (ly & 2) != 0;
                float k = mid ? k1 : k0;
                const float sx = mid ? k0 : k1;
                k += __shfl_xor_sync(0xffffffffu, sx, 2);
                k += __shfl_xor_sync(0xffffffffu, k, 1);
                if (writer) {
                    float res = pv + fmaf(pv, ds, k);
                    if (nz) res = fmaf(scale, nzv, res);
                    out[(long long)Yd * a.Wo] = res;
                }
            }
#pragma unroll
            for (int x = 0; x < 4; ++x) { acc[2][x] = acc[1][x]; acc[1][x] = acc[0][x]; acc[0][x] = 0.0f; }
        }
        // the last chunk of the band: both the step that loaded it and the bottom-halo step are done
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty + 8 * slot);
        if (++slot == kDepth) { slot = 0; par ^= 1; }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

bool tma_shape_ok(const DegradeArgs& a, const char** why) {
    const Geometry& g = a.g;
    *why = "";
    if (g.kh != kK || g.kw != kK || g.stride != kS || g.KH != kKW) { *why = "needs k=13 and factor 8 (box mean)"; return false; }
    if (a.pad_mode != KMSR_PAD_REPLICATE) { *why = "needs replicate padding"; return false; }
    if (a.W != 256) { *why = "needs W == 256"; return false; }
    if (a.H < 8 || a.H % 8 != 0) { *why = "needs H % 8 == 0"; return false; }
    if (a.patch_offsets) { *why = "patch_offsets (scene windows) not covered"; return false; }
    if (((uintptr_t)a.hr & 15) || (a.sH & 3) || (a.sC & 3) || (a.N > 1 && (a.sN & 3))) {
        *why = "HR base / strides not 16-byte aligned"; return false;
    }
    if (a.sH < a.W || a.sC < 1 || (a.N > 1 && a.sN < 1)) { *why = "non-positive strides"; return false; }
    if (a.N >= (1ll << 31) || a.N * a.C >= (1ll << 40)) { *why = "too many patches"; return false; }
    return true;
}

int launch_degrade_tma(const DegradeArgs& a, cudaStream_t st) {
    EncodeTiledFn enc = get_encode();
    KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "degrade (tma): cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    // [N, C, H, W/2] of 64-bit elements; box = 138 x 8 x 1 x 1 (x starts at -4: zero-filled halo)
    cuuint64_t gdim[4] = {(cuuint64_t)(a.W / 2), (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.N};
    const long long sN = a.N > 1 ? a.sN : (long long)a.C * a.sC;
    cuuint64_t gstr[3] = {(cuuint64_t)a.sH * 4, (cuuint64_t)a.sC * 4, (cuuint64_t)sN * 4};
    cuuint32_t box[4] = {(cuuint32_t)(kRowF / 2), 8, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)a.hr, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "degrade (tma): cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);

    TmaArgs t;
    t.comp = a.comp; t.dsum = a.dsum; t.kidx = a.kidx; t.sigma = a.sigma; t.pool = a.pool; t.nidx = a.nidx;
    t.lr = a.lr; t.nbands = a.N * a.C; t.C = a.C; t.H = a.H; t.Ho = a.g.Ho; t.Wo = a.g.Wo;
    t.nchunks = a.H / 8 + 1; t.noise_mode = a.noise_mode;

    int dev = 0, sms = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long grid = (t.nbands + kStreams - 1) / kStreams;
    if (grid > sms) grid = sms;
    KMSR_CUDA_OK(cudaFuncSetAttribute(degrade_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    set_algo("tma");
    degrade_tma_kernel<<<(unsigned)grid, kThreads, kSmemBytes, st>>>(tmap, t);
    KMSR_LAUNCH_CHECK("degrade_tma_kernel");
    return KMSR_OK;
}

}  // namespace kmsr
