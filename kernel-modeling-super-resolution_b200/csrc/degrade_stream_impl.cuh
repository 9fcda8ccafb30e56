// degrade_stream_impl.cuh -- generic TMA row-streaming fused blur + box-mean downsample + noise kernel.
//
// The sweep shapes of BASELINE config 5 (kernel 11..31, patch 64..512, factor 2/4/8) and every other
// box-mean call with an odd kernel that the headline kernel (degrade_tma.cu: k = 13, factor 8, W = 256)
// does not take.  Same arithmetic (C_30apply_kernel_to_landsat.py:68-124 with the box mean folded
// into a stride-S composite kernel, E_make_train_data.py:72-74 / train_gemini.py:137 noise in the
// epilogue), same data movement idea, parameterised by <K, S>:
//
//  * a band (or a 256-column block of a wider band) is a *stream*: its rows cross HBM -> SMEM exactly
//    once, S rows per TMA tile (cp.async.bulk.tensor, producer thread per stream or per two streams,
//    full/empty mbarriers, ring of D tiles).  Tiles hold image rows only; the replicate halo is made by
//    clamping row addresses (top / bottom) and substituting registers (left / right).
//  * lane = (ly, gx): ly = input row within the tile (S rows), gx = group of 4 adjacent LR columns
//    (32/S groups per warp, 128 input columns per warp whatever S is).  At step i a lane holds padded
//    row S*i + ly, which meets output rows i - q through composite rows u = ly + S*q, q < Q =
//    ceil(KW/S): Q accumulator sets per lane, rotated by value; the oldest one completes each step, is
//    reduce-scattered over the S row lanes by shuffles and written with the noise term.
//  * composite-kernel rows come from shared memory (a per-warp copy with an odd 16-byte pitch: the S
//    distinct rows a warp reads at once never collide); pixels are pivot-shifted and multiplied as
//    packed FFMA2 pairs exactly like the headline kernel.
// Included by degrade_stream.cu (replicate padding) and degrade_stream_zero.cu (zero padding): one translation unit per
// padding mode, compiled in parallel.
#pragma once
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "tma_util.cuh"

namespace kmsr {

namespace {

constexpr int kConsumers = 8;                  // consumer warps per CTA
constexpr int kProducers = 4;                  // producer warps (one elected thread each)
constexpr int kThreadsS = (kConsumers + kProducers) * 32;
constexpr bool kWide4 = true;                  // factor 4, 256-wide blocks: 8 LR columns per lane (one warp per block row)
constexpr int kRenameQ = 5;                    // accumulator sets are renamed (steps unrolled by Q) up to this many; 6..8 measured 4-9 % slower (code size)

constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

template <int K, int S, int TX>
struct Cfg {
    static constexpr int kTX = TX;                               // LR columns per lane
    static constexpr int KW = K + S - 1;                         // composite taps per row / column
    static constexpr int PAD = K / 2;
    static constexpr int Q = (KW + S - 1) / S;                   // live output rows per lane
    static constexpr int NV = TX >= S ? TX / S : 1;               // outputs a lane writes per step after the reduce-scatter
    static constexpr int GW = kTX * S;                           // input columns per lane group
    static constexpr int GX = 32 / S;                            // groups per warp
    static constexpr int AL = (PAD + 3) / 4 * 4;                 // staged columns left of the block (16-byte aligned)
    static constexpr int SKEW = AL - PAD;                        // first needed float of a lane's aligned segment
    static constexpr int SEG = KW + (kTX - 1) * S;               // floats a lane needs per row
    static constexpr int LOADF = (SKEW + SEG + 3) / 4 * 4;       // floats it loads (LDS.128 granules)
    static constexpr int NP = SEG / 2;                           // pixel pairs (SEG is even: K odd, S even)
    static constexpr int TP = KW / 2;                            // tap pairs
    static constexpr int WP = ((KW + 3) / 4 * 4 / 4) % 2 ? (KW + 3) / 4 * 4 : (KW + 3) / 4 * 4 + 4;   // weight row pitch, /4 odd
    static constexpr int WROWS = Q * S;                          // rows >= KW are zero
    // steps unrolled with renamed accumulator sets: all Q when Q <= kRenameQ, else none (one in-place rotation per
    // step): unrolling 2 or 4 steps of the larger shapes was measured 3-9 % slower (code size), r49
    static constexpr int U = Q <= kRenameQ ? Q : 1;
    static constexpr int GC = cgcd(Q, U);                        // cycles of the rotation by U
    static constexpr bool PERSTEP = S == 4 && K >= 21;           // interior / general fetch chosen per step instead of per block
    static_assert(K % 2 == 1 && (S == 2 || S == 4 || S == 8), "odd kernel, factor 2/4/8");
};

// Q consecutive steps with the accumulator-set index as a compile-time constant (renaming instead of moving)
template <int J, int Q>
struct Unroll {
    template <class F>
    static __device__ __forceinline__ void run(F& step, int i, int nsteps) {
        if (i + J < nsteps) step(std::integral_constant<int, J>{}, i + J);
        Unroll<J + 1, Q>::run(step, i, nsteps);
    }
};
template <int Q>
struct Unroll<Q, Q> {
    template <class F>
    static __device__ __forceinline__ void run(F&, int, int) {}
};

struct StreamArgs {
    const float* comp;      // [nK, C, KW, KWp]
    int compPitch;          // KWp
    const float* dsum;
    const int* kidx;
    const float* sigma;
    const float* pool;
    const int* nidx;
    float* lr;
    long long nitems;       // bands * nblk
    int C, H, W, Ho, Wo;
    int nblk;               // column blocks per band (W / BW)
    int BW;                 // block width in pixels (<= 256)
    int ng;                 // lane groups per block row (BW / GW)
    int nw;                 // consumer warps per stream (1 or 2)
    int ns;                 // streams per CTA (kConsumers / nw)
    int depth;              // ring slots per stream
    int pitchF;             // floats per staged row
    int chunkBytes;         // S * pitchF * 4 (what one TMA tile delivers)
    int slotBytes;          // chunkBytes rounded up to 128 (TMA destinations are 128-byte aligned)
    int nchunks;            // H / S
    int pad_mode, noise_mode;
    unsigned ringOff, barOff, wOff;
    int nofast;             // KMSR_STREAM_NOFAST=1 (measurement aid): every step takes the general path
    int wbuf;               // weight buffers per warp: 2 = the next band's kernel is prefetched, 1 = staged between bands
};

// REPL: replicate padding (C_30:107-109) as a compile-time fact -- the zero-padding variant (train_gemini.py:128) zeroes
// the rows outside the image in registers every step and needs no edge substitution; compiled into its own translation
// unit (degrade_stream_zero.cu) so that the 42..62 predicated moves per step it costs stay out of the replicate kernels.
template <int K, int S, int TX, bool REPL>
__global__ void __launch_bounds__(kThreadsS, 1)
degrade_stream_kernel(const __grid_constant__ CUtensorMap tmap, const StreamArgs a) {
    using G = Cfg<K, S, TX>;
    constexpr int kTX = TX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.depth;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + a.barOff);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + a.ns * D);
    if (threadIdx.x == 0) {
        for (int i = 0; i < a.ns * D; ++i) {
            mbar_init(full0 + 8 * i, 1);
            mbar_init(empty0 + 8 * i, a.nw);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const long long GS = (long long)gridDim.x * a.ns;        // streams in the grid

    if (warp >= kConsumers) {
        // ============ producers: warp kConsumers + p feeds streams p, p + 4, ... of this CTA ============
        if (lane != 0) return;
        const int p = warp - kConsumers;
        const int nmine = (a.ns - p + kProducers - 1) / kProducers;          // 0, 1 or 2 streams
        if (nmine <= 0) return;
        long long item[2], pn[2];
        int chunk[2], slot[2], pc[2], px[2];
        uint32_t par[2];
        auto locate = [&](int j) {               // item -> (patch, band, column block)
            if (item[j] >= a.nitems) return;
            const long long band = item[j] / a.nblk;
            px[j] = ((int)(item[j] - band * a.nblk) * a.BW - G::AL) / 2;
            pn[j] = band / a.C;
            pc[j] = (int)(band - pn[j] * a.C);
        };
        for (int j = 0; j < 2; ++j) {
            item[j] = (long long)blockIdx.x * a.ns + p + j * kProducers;
            if (j >= nmine) item[j] = a.nitems;
            chunk[j] = 0; slot[j] = 0; par[j] = 1; pn[j] = 0; pc[j] = 0; px[j] = 0;
            locate(j);
        }
        const uint64_t policy = l2_evict_first_policy();
        bool active = true;
        while (active) {
            active = false;
            for (int j = 0; j < 2; ++j) {
                if (item[j] >= a.nitems) continue;
                active = true;
                const int s = p + j * kProducers;
                const uint32_t sfull = full0 + 8 * (s * D + slot[j]), sempty = empty0 + 8 * (s * D + slot[j]);
                mbar_wait_relaxed(sempty, par[j]);
                mbar_arrive_expect_tx(sfull, a.chunkBytes);
                tma_load_4d_hint(smem_u32(smem_raw + a.ringOff + (size_t)(s * D + slot[j]) * a.slotBytes), &tmap, px[j],
                                 S * chunk[j], pc[j], (int)pn[j], sfull, policy);
                if (++slot[j] == D) { slot[j] = 0; par[j] ^= 1; }
                if (++chunk[j] == a.nchunks) { chunk[j] = 0; item[j] += GS; locate(j); }
            }
        }
        return;
    }

    // ======================================= consumers =======================================
    const int s = warp / a.nw, wq = warp - s * a.nw;          // stream, warp within the stream
    const int ly = lane & (S - 1), gx = lane / S;
    const int g_raw = wq * G::GX + gx;                        // lane group within the block row
    const bool lane_on = g_raw < a.ng;                        // W = 64: half of the groups have no columns
    const int g = lane_on ? g_raw : a.ng - 1;
    const unsigned char* sring = smem_raw + a.ringOff + (size_t)s * D * a.slotBytes;
    const uint32_t sfull = full0 + 8 * s * D, sempty = empty0 + 8 * s * D;
    // per-warp weight copies: buffer 0, and buffer 1 when the next band's kernel is prefetched (a.wbuf == 2)
    float* wbase = reinterpret_cast<float*>(smem_raw + a.wOff) + (size_t)warp * a.wbuf * (G::WROWS * G::WP);
    constexpr bool replicate = REPL;
    const bool noisy = a.noise_mode != KMSR_NOISE_NONE;
    const long long ohw = (long long)a.Ho * a.Wo;
    const int nsteps = a.Ho + G::Q - 1;

    // ring bookkeeping, per item: chunks [rel, wai) of the item are resident, chunk `rel` sits in slot rslot
    int wslot = 0, rslot = 0;
    uint32_t wpar = 0;

    for (int r = lane; r < a.wbuf * G::WROWS * G::WP; r += 32) wbase[r] = 0.0f;    // rows >= KW and pitch padding stay zero
    __syncwarp();

    // composite kernel of (kid, c) -> weight buffer `buf` with cp.async (4-byte copies: the bank rows are not 16-byte
    // multiples for every K); the caller commits and waits
    auto stage_weights = [&](int buf, int kid, int c) {
        const float* kc = a.comp + ((long long)kid * a.C + c) * (G::KW * a.compPitch);
        const uint32_t dst = smem_u32(wbase + (size_t)buf * (G::WROWS * G::WP));
        for (int e = lane; e < G::KW * G::KW; e += 32) {
            const int u = e / G::KW, v = e - u * G::KW;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4 * (u * G::WP + v)), "l"(kc + u * a.compPitch + v)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    struct ItemParams { long long band; int blk, c, kid, nid; float ds, scale; };
    auto load_params = [&](long long item) {
        ItemParams q;
        q.band = item / a.nblk;
        q.blk = (int)(item - q.band * a.nblk);
        const long long n = q.band / a.C;
        q.c = (int)(q.band - n * a.C);
        q.kid = a.kidx ? __ldg(a.kidx + n) : 0;
        q.nid = noisy ? __ldg(a.nidx + n) : 0;
        q.ds = __ldg(a.dsum + (long long)q.kid * a.C + q.c);
        q.scale = a.noise_mode == KMSR_NOISE_SIGMA ? __ldg(a.sigma + (long long)q.kid * a.C + q.c) : 1.0f;
        return q;
    };

    long long item = (long long)blockIdx.x * a.ns + s;
    ItemParams cur;
    int wcur = 0;
    if (item < a.nitems) {
        cur = load_params(item);
        stage_weights(0, cur.kid, cur.c);
    }
    for (; item < a.nitems; item += GS) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        const float* wsm = wbase + (size_t)wcur * (G::WROWS * G::WP);
        const ItemParams it = cur;
        const bool has_next = item + GS < a.nitems;
        if (has_next) {
            cur = load_params(item + GS);                      // consumed at the next band: the latency hides behind this one
            if (a.wbuf == 2) { stage_weights(wcur ^ 1, cur.kid, cur.c); }
        }
        const bool edge_l = replicate && it.blk == 0, edge_r = replicate && it.blk == a.nblk - 1;
        const float ds = it.ds, scale = it.scale;
        const float* nz = a.pool + ((long long)it.nid * a.C + it.c) * ohw;

        u64 A[G::Q][kTX];
#pragma unroll
        for (int q = 0; q < G::Q; ++q)
#pragma unroll
            for (int x = 0; x < kTX; ++x) A[q][x] = 0ull;
        float pv = 0.0f;
        u64 npv2 = 0ull;
        int rel = 0, wai = 0;                                  // chunks of this item released / waited for

        // output columns after the reduce-scatter over the S row lanes
        const int col0 = it.blk * (a.BW / S) + kTX * g;       // first LR column of this lane group
        const int ox = kTX >= S ? (kTX / S) * ly : ly / (S / kTX);
        const bool writer = lane_on && (kTX >= S || (ly & (S / kTX - 1)) == 0);
        float* outp = a.lr + it.band * ohw + col0 + ox;
        const float* nzp = nz + col0 + ox;
        const int lane_off = G::GW * g;

        // Hand `count` chunks (from slot rslot on) back to the producer once this step's rows sit in registers.
        // The shared-memory loads just ISSUED must have been PERFORMED before the arrive becomes visible: an
        // mbarrier.arrive does not wait for the warp's outstanding LDS, and with a nearly empty ring the producer's
        // refill can land in the slot before a queued LDS has read it (seen as run-to-run differences in the last
        // columns of a lane group once the interior fast path shortened the distance between the loads and the
        // arrive, r77).  The CTA-scope fence completes the loads of every lane; __syncwarp orders the lanes.
        auto release_after_loads = [&](int count) {
            __threadfence_block();
            __syncwarp();
            for (; count > 0; --count) {
                if (lane == 0) mbar_arrive(sempty + 8 * rslot);
                if (++rslot == D) rslot = 0;
                ++rel;
            }
        };

        // One step.  SH = position of the step inside its unrolled block: the set of output row i - q is (SH - q) mod Q.
        auto body = [&](auto sh_tag, const int i, const float (&e)[G::LOADF], const bool row_ok) {
            constexpr int SH = decltype(sh_tag)::value;
            // noise of the row that completes in this step: issue the load before the arithmetic
            const int Y = i - (G::Q - 1);
            float nzv[G::NV];
#pragma unroll
            for (int j = 0; j < G::NV; ++j) nzv[j] = (noisy && writer && Y >= 0) ? __ldg(nzp + (long long)Y * a.Wo + j) : 0.0f;

            // d[j] = pixel column GW*g - PAD + j of the block, j < SEG
            float d[G::SEG];
#pragma unroll
            for (int j = 0; j < G::SEG; ++j) d[j] = e[j + G::SKEW];
            if (!replicate && !row_ok) {
#pragma unroll
                for (int j = 0; j < G::SEG; ++j) d[j] = 0.0f;
            }
            if (edge_l) {
                // columns < 0 of the image take pixel 0: group gg has PAD - GW*gg of them
#pragma unroll
                for (int gg = 0; gg * G::GW < G::PAD; ++gg) {
                    if (g == gg) {
                        const int hl = G::PAD - G::GW * gg;
#pragma unroll
                        for (int j = 0; j < G::SEG; ++j)
                            if (j < hl) d[j] = d[hl < G::SEG ? hl : G::SEG - 1];
                    }
                }
            }
            if (edge_r) {
                // columns >= W take pixel W-1: the group gg from the right end sees column W at d[PAD + GW*(gg+1)]
#pragma unroll
                for (int gg = 0; G::PAD + G::GW * (gg + 1) < G::SEG; ++gg) {
                    if (g == a.ng - 1 - gg) {
                        const int hr = G::PAD + G::GW * (gg + 1);
#pragma unroll
                        for (int j = 0; j < G::SEG; ++j)
                            if (j >= hr) d[j] = d[hr - 1];
                    }
                }
            }
            // pivot shift: packed FADD2 when the segment starts on a register pair (even SKEW); with an odd SKEW every pair
            // would first need two moves to be formed, so two scalar FADDs write the halves of the pair directly
            u64 P[G::NP];
#pragma unroll
            for (int m = 0; m < G::NP; ++m)
                P[m] = (G::SKEW % 2 == 0) ? add2(pack2(d[2 * m], d[2 * m + 1]), npv2) : pack2(d[2 * m] - pv, d[2 * m + 1] - pv);

            // ---- multiply-accumulate: composite row u = ly + S*q meets output row i - q ----
#pragma unroll
            for (int q = 0; q < G::Q; ++q) {
                constexpr int dummy = 0; (void)dummy;
                const int j = (SH - q + G::Q) % G::Q;                       // accumulator set of output row i - q
                const int u = ly + S * q;
                const float* wrow = wsm + u * G::WP;
                u64 T[kTX];
#pragma unroll
                for (int x = 0; x < kTX; ++x) T[x] = A[j][x];
#pragma unroll
                for (int t4 = 0; t4 < (G::TP + 1) / 2; ++t4) {
                    const ulonglong2 w2 = reinterpret_cast<const ulonglong2*>(wrow)[t4];
#pragma unroll
                    for (int x = 0; x < kTX; ++x) {
                        T[x] = fma2(w2.x, P[(S * x) / 2 + 2 * t4], T[x]);
                        if (2 * t4 + 1 < G::TP) T[x] = fma2(w2.y, P[(S * x) / 2 + 2 * t4 + 1], T[x]);
                    }
                }
                // a row below the window (u >= KW) must not even add 0 * pixel: a NaN there would poison an
                // output the reference keeps
                const bool live = (q + 1) * S <= G::KW || u < G::KW;
#pragma unroll
                for (int x = 0; x < kTX; ++x) if (live) A[j][x] = T[x];
            }

            // ---- the oldest set completes: reduce over the S row lanes, epilogue, store ----
            constexpr int JO = (SH + 1) % G::Q;                             // == (SH - (Q-1)) mod Q
            if (Y >= 0) {
                float v[kTX];
#pragma unroll
                for (int x = 0; x < kTX; ++x) v[x] = lo2(A[JO][x]) + hi2(A[JO][x]);
                // reduce-scatter over the S row lanes: each stage halves the values a lane keeps (upper half for the
                // lanes with the stage bit set) until one is left, the remaining stages are plain pair sums
                int nv = kTX;
#pragma unroll
                for (int b = S / 2; b >= 1; b >>= 1) {
                    if (nv > 1) {
                        const bool up = (ly & b) != 0;
                        const int half = nv / 2;
#pragma unroll
                        for (int j = 0; j < kTX / 2; ++j) {
                            if (j < half) {
                                const float keep = up ? v[half + j] : v[j];
                                const float send = up ? v[j] : v[half + j];
                                v[j] = keep + __shfl_xor_sync(0xffffffffu, send, b);
                            }
                        }
                        nv = half;
                    } else {
                        v[0] += __shfl_xor_sync(0xffffffffu, v[0], b);
                    }
                }
                if (writer) {
                    float* out = outp + (long long)Y * a.Wo;
#pragma unroll
                    for (int j = 0; j < G::NV; ++j) {
                        float res = pv + fmaf(pv, ds, v[j]);
                        if (noisy) res = fmaf(scale, nzv[j], res);
                        out[j] = res;
                    }
                }
            }
#pragma unroll
            for (int x = 0; x < kTX; ++x) A[JO][x] = 0ull;                 // becomes the fresh set of step i + 1
        };

        // A step = fetch this lane's row + `body` (edges, pivot, FFMA2, reduce, store).  Two ways to fetch:
        // general -- clamped rows, any number of chunks to wait for / release (band top and bottom);
        // interior -- every row of the step lies inside the image and the output row exists.  Chunk accounting is then
        // fixed: rows S*i - PAD .. S*i - PAD + S - 1 start `o_rows` rows into chunk `rel` (slot rslot) and spill into the
        // next one, which is the single chunk this step waits for; chunk `rel` is released once the rows are in registers.
        auto fetch_general = [&](const int i, float (&e)[G::LOADF]) -> bool {
            // ---- rows this step needs: padded rows S*i .. S*i+S-1 = image rows S*i - PAD + ly ----
            const int rr_raw = S * i + ly - G::PAD;
            const int rr = min(max(rr_raw, 0), a.H - 1);
            const bool row_ok = rr_raw >= 0 && rr_raw < a.H;               // zero padding: rows outside are zeros
            const int hi_chunk = min(max(S * i + S - 1 - G::PAD, 0), a.H - 1) / S;
            while (wai <= hi_chunk) {
                mbar_wait(sfull + 8 * wslot, wpar);
                if (++wslot == D) { wslot = 0; wpar ^= 1; }
                ++wai;
            }
            int slot = rslot + (rr / S - rel);                              // chunk rel sits in slot rslot
            if (slot >= D) slot -= D;
            const float* src = reinterpret_cast<const float*>(sring + (size_t)slot * a.slotBytes) + (rr % S) * a.pitchF + lane_off;
            if (i == 0) {
                // pivot: pixel (0, first column of the lane group) -- row 0 is in the item's first chunk (slot rslot)
                pv = reinterpret_cast<const float*>(sring + (size_t)rslot * a.slotBytes)[G::AL + lane_off];
                if (!isfinite(pv)) pv = 0.0f;
                npv2 = pack2(-pv, -pv);
            }
#pragma unroll
            for (int j = 0; j < G::LOADF / 4; ++j) {
                const float4 t = reinterpret_cast<const float4*>(src)[j];
                e[4 * j + 0] = t.x; e[4 * j + 1] = t.y; e[4 * j + 2] = t.z; e[4 * j + 3] = t.w;
            }
            // release the chunks no later step needs (all of them after the item's last step)
            const int lo_next = i + 1 < nsteps ? min(max(S * (i + 1) - G::PAD, 0), a.H - 1) / S : a.nchunks;
            release_after_loads(lo_next - rel);
            return row_ok;
        };
        constexpr int o_rows = (S - G::PAD % S) % S;
        const bool lane_hi = o_rows + ly >= S;
        const int rowin = o_rows + ly - (lane_hi ? S : 0);
        auto fetch_interior = [&](float (&e)[G::LOADF]) {
            mbar_wait(sfull + 8 * wslot, wpar);
            if (++wslot == D) { wslot = 0; wpar ^= 1; }
            ++wai;
            int slot = rslot + (lane_hi ? 1 : 0);
            if (slot >= D) slot -= D;
            const float* src = reinterpret_cast<const float*>(sring + (size_t)slot * a.slotBytes) + rowin * a.pitchF + lane_off;
#pragma unroll
            for (int j = 0; j < G::LOADF / 4; ++j) {
                const float4 t = reinterpret_cast<const float4*>(src)[j];
                e[4 * j + 0] = t.x; e[4 * j + 1] = t.y; e[4 * j + 2] = t.z; e[4 * j + 3] = t.w;
            }
            release_after_loads(1);
        };
        const int i_lo = a.nofast ? (1 << 30) : max((G::PAD + S - 1) / S, G::Q - 1);      // first step with all rows >= 0 and an output row
        // last step whose rows are all <= H - 1 AND whose successor no longer needs the step's first chunk (the next
        // step's first row, S (i + 1) - PAD, must not be clamped back into it: matters when PAD is a multiple of S)
        const int i_hi = (a.H - 1 + G::PAD - S) / S;
        // per block of U steps (two copies of the block: fastest where the code fits) or per step (one copy, both fetches
        // inline: +9-15 % for k = 31 / 21 at factor 4, whose blocks are large and spill; 3-10 % slower elsewhere, r79)
        auto step_general = [&](auto sh_tag, const int i) {
            float e[G::LOADF];
            const bool row_ok = fetch_general(i, e);
            body(sh_tag, i, e, row_ok);
        };
        auto step_interior = [&](auto sh_tag, const int i) {
            float e[G::LOADF];
            fetch_interior(e);
            body(sh_tag, i, e, true);
        };
        auto step_any = [&](auto sh_tag, const int i) {
            float e[G::LOADF];
            bool row_ok = true;
            if (i >= i_lo && i <= i_hi) fetch_interior(e);
            else row_ok = fetch_general(i, e);
            body(sh_tag, i, e, row_ok);
        };

        // U consecutive steps run with compile-time set indices (renaming); when U < Q the sets are then rotated by U
        // positions by value, once per U steps instead of once per step
#pragma unroll 1
        for (int i = 0; i < nsteps; i += G::U) {
            if constexpr (G::PERSTEP) {
                Unroll<0, G::U>::run(step_any, i, nsteps);
            } else {
                if (i >= i_lo && i + G::U - 1 <= i_hi) Unroll<0, G::U>::run(step_interior, i, nsteps);
                else Unroll<0, G::U>::run(step_general, i, nsteps);
            }
            if constexpr (G::U != G::Q) {
                // in-place rotation A[m] <- A[(m + U) mod Q]: gcd(Q, U) cycles, one spare set
                constexpr int GC = G::GC;
#pragma unroll
                for (int c0 = 0; c0 < GC; ++c0) {
                    u64 T[kTX];
#pragma unroll
                    for (int x = 0; x < kTX; ++x) T[x] = A[c0][x];
                    int cur = c0;
#pragma unroll
                    for (int t = 0; t < G::Q / GC - 1; ++t) {
                        const int nxt = (cur + G::U) % G::Q;
#pragma unroll
                        for (int x = 0; x < kTX; ++x) A[cur][x] = A[nxt][x];
                        cur = nxt;
                    }
#pragma unroll
                    for (int x = 0; x < kTX; ++x) A[cur][x] = T[x];
                }
            }
        }
        if (a.wbuf == 2) {
            wcur ^= 1;
        } else if (has_next) {
            __syncwarp();                                      // every lane is done reading the single buffer
            stage_weights(0, cur.kid, cur.c);
        }
    }
}

template <int K, int S, int TX, bool REPL>
int launch_kS(const DegradeArgs& a, StreamArgs& t, int sms, cudaStream_t st) {
    using G = Cfg<K, S, TX>;
    t.nw = t.BW / (32 * TX) > 0 ? t.BW / (32 * TX) : 1;
    t.ns = kConsumers / t.nw;
    // row pitch: AL | BW | right extent of the last group, 16-byte granules, odd count
    const int right = G::LOADF - G::GW - G::AL;
    int pitch = G::AL + t.BW + (right > 0 ? right : 0);
    pitch = (pitch + 3) / 4 * 4;
    if ((pitch / 4) % 2 == 0) pitch += 4;
    t.pitchF = pitch;
    t.chunkBytes = S * pitch * 4;
    t.slotBytes = (t.chunkBytes + 127) / 128 * 128;
    t.nchunks = t.H / S;
    t.ng = t.BW / G::GW;
    int dev = 0, max_smem = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // two weight buffers per warp (prefetch of the next band's kernel) when the ring keeps at least 4 slots
    t.wbuf = 2;
    size_t wbytes = (size_t)kConsumers * 2 * G::WROWS * G::WP * 4;
    if (((size_t)max_smem - (wbytes + 2 * 8 * 8 * 16 + 1024)) / ((size_t)t.ns * t.slotBytes) < 4) {
        t.wbuf = 1;
        wbytes /= 2;
    }
    const size_t fixed = wbytes + 2 * 8 * 8 * 16 + 1024;
    int depth = (int)(((size_t)max_smem - fixed) / ((size_t)t.ns * t.slotBytes));
    // a ring must at least hold the rows one step touches plus one tile in flight
    const int need = 3;
    if (depth > 16) depth = 16;
    KMSR_REQUIRE(depth >= need, KMSR_E_UNSUPPORTED, "degrade (stream): k=%d factor=%d W=%d does not fit shared memory", K, S, t.W);
    t.depth = depth;
    // chunk size must be a multiple of 128 B for the tile base alignment
    t.ringOff = 0;
    size_t ring = (size_t)t.ns * depth * t.slotBytes;
    ring = (ring + 127) / 128 * 128;
    t.barOff = (unsigned)ring;
    t.wOff = (unsigned)(ring + ((2 * t.ns * depth * 8 + 127) / 128) * 128);
    const size_t smem = t.wOff + wbytes;
    KMSR_REQUIRE(pitch / 2 <= 256, KMSR_E_UNSUPPORTED, "degrade (stream): staged row of %d floats exceeds the TMA box limit", pitch);
    EncodeTiledFn enc = get_tensor_map_encoder();
    KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "degrade (stream): cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap tmap;
    cuuint64_t gdim[4] = {(cuuint64_t)(a.W / 2), (cuuint64_t)a.H, (cuuint64_t)a.C, (cuuint64_t)a.N};
    const long long sN = a.N > 1 ? a.sN : (long long)a.C * a.sC;
    cuuint64_t gstr[3] = {(cuuint64_t)a.sH * 4, (cuuint64_t)a.sC * 4, (cuuint64_t)sN * 4};
    cuuint32_t box[4] = {(cuuint32_t)(pitch / 2), (cuuint32_t)S, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void*)a.hr, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "degrade (stream): cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
    auto kern = degrade_stream_kernel<K, S, TX, REPL>;
    KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long grid = (t.nitems + t.ns - 1) / t.ns;
    if (grid > sms) grid = sms;
    kern<<<(unsigned)grid, kThreadsS, smem, st>>>(tmap, t);
    KMSR_LAUNCH_CHECK("degrade_stream_kernel");
    return KMSR_OK;
}

template <int K, bool REPL>
int launch_k(const DegradeArgs& a, StreamArgs& t, int S, int sms, cudaStream_t st) {
    // 64-wide patches: two LR columns per lane keep all 32 lanes of the (single) warp of a stream busy
    if (t.BW == 64) {
        switch (S) {
            case 2: return launch_kS<K, 2, 2, REPL>(a, t, sms, st);
            case 4: return launch_kS<K, 4, 2, REPL>(a, t, sms, st);
            default: return launch_kS<K, 8, 2, REPL>(a, t, sms, st);
        }
    }
    switch (S) {
        case 2: return launch_kS<K, 2, 4, REPL>(a, t, sms, st);
        case 4:
            // 8 LR columns per lane when the accumulator sets still fit (Q <= 4, i.e. k <= 13: +6..16 %; k = 15
            // spills and loses 20 %, r52)
            if constexpr (kWide4 && (K + 3 + 3) / 4 <= 4) { if (t.BW == 256) return launch_kS<K, 4, 8, REPL>(a, t, sms, st); }
            return launch_kS<K, 4, 4, REPL>(a, t, sms, st);
        default: return launch_kS<K, 8, 4, REPL>(a, t, sms, st);     // 8 columns per lane at factor 8: 8 streams no longer fit, spills
    }
}

template <bool REPL>
int launch_stream_k(const DegradeArgs& a, StreamArgs& t, int K, int S, int sms, cudaStream_t st) {
    switch (K) {
        case 11: return launch_k<11, REPL>(a, t, S, sms, st);
        case 13: return launch_k<13, REPL>(a, t, S, sms, st);
        case 15: return launch_k<15, REPL>(a, t, S, sms, st);
        case 21: return launch_k<21, REPL>(a, t, S, sms, st);
        default: return launch_k<31, REPL>(a, t, S, sms, st);
    }
}

// host-side arguments common to both padding modes
inline int fill_stream_args(const DegradeArgs& a, StreamArgs& t, int* sms) {
    const Geometry& g = a.g;
    t.comp = a.comp; t.compPitch = g.KWp; t.dsum = a.dsum; t.kidx = a.kidx; t.sigma = a.sigma; t.pool = a.pool;
    t.nidx = a.nidx; t.lr = a.lr;
    t.C = a.C; t.H = a.H; t.W = a.W; t.Ho = g.Ho; t.Wo = g.Wo;
    t.BW = a.W >= 256 ? 256 : a.W;
    t.nblk = a.W / t.BW;
    t.nw = 1; t.ns = kConsumers;             // set per <K, S, TX> in launch_kS
    t.nitems = a.N * a.C * t.nblk;
    t.pad_mode = a.pad_mode; t.noise_mode = a.noise_mode;
    const char* nf = getenv("KMSR_STREAM_NOFAST");          // read per call: tests toggle it to compare the two step paths
    t.nofast = nf ? atoi(nf) : 0;
    int dev = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, dev));
    return KMSR_OK;
}

}  // namespace

}  // namespace kmsr
