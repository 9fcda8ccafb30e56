// degrade_box.cu -- TMA box-tile fused blur + box-mean downsample + noise kernel for the sweep shapes of BASELINE
// config 5 that the row-streaming kernels serve badly: every factor-2 and factor-4 shape (FP32-bound or close to the
// ridge) and the 64-wide patches at any factor.
//
// Same arithmetic as the other degrade kernels (C_30apply_kernel_to_landsat.py:68-124 with the box mean folded into a
// stride-S composite kernel K' of (k + S - 1)^2 taps, E_make_train_data.py:72-74 / train_gemini.py:137 noise in the
// epilogue).  It is the register-tile stencil of degrade_reg.cu -- a thread owns an 8 x 16 block of HR pixels' worth of
// outputs (8/S x 16/S LR pixels) in registers, walks the k + 7 input rows of its window once, composite-kernel rows
// are broadcast LDS.128, pixels and taps meet as packed FFMA2 pairs -- with what held that kernel at 0.34-0.67 of the
// FP32 peak removed:
//
//  * staging is TMA, not 160 four-byte cp.async per thread: a 128 x 128 tile of one band plus its halo (or a whole
//    band of at most 64 x 64 pixels per warp) arrives as 8 tensor boxes, one per row residue j = row mod 8, issued by
//    8 lanes.  Out-of-bounds zero fill is the zero-padding halo.  Nothing rewrites shared memory afterwards (a first
//    version copied the replicate halo and applied the pivot shift in place: two latency-bound passes and two barriers
//    per tile cost as many issue slots as the arithmetic, ncu r2b): replicate rows clamp by address, the halo columns
//    of the two edge threads are substituted in registers, the pivot shift (x - pivot) is one FADD2
//    per loaded pair;
//  * the layout is [j][q = row div 8][x] with a row pitch that is an odd multiple of 16 bytes.  The 8 lanes of a
//    quarter-warp own thread rows ly .. ly+7 (8 HR rows apart, same columns): they read the same j, consecutive q
//    -- 8 distinct 16-byte bank groups, so every LDS.128 of the inner loop is conflict free (the block-column layout
//    of degrade_reg.cu measured 25 % conflicting wavefronts);
//  * input rows that meet every output row of the thread (most of them) run as one straight-line block, tap-major
//    over all 8/S x 16/S accumulators, so the weight loads of the next taps sit under the FFMA2 of the current ones.
// Kernels with an odd halo (k = 11, 15, 31) put their first tap on an odd column: the first and the last tap of a row
// are issued as scalar FFMA on one half of the accumulator pair and the k + S - 3 taps between them as aligned FFMA2
// pairs (same FMA-pipe time as an even halo, no zero tap multiplied with a real pixel).
// Tile mode: one CTA of 128 threads per tile, two CTAs per SM (shared memory), the hardware schedules; a CTA's TMA wait
// overlaps the other CTA's arithmetic.  Band mode: persistent CTAs of 1-4 warps, every warp draws whole bands from a
// ticket counter and issues the next band's TMA before it writes the current one out.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tma_util.cuh"

namespace kmsr {

namespace {

constexpr int kBoxThreads = 128;

template <int K, int S>
struct BCfg {
    static constexpr int KW = K + S - 1;                 // composite taps per row / column (even)
    static constexpr int TP = KW / 2;                    // tap pairs
    static constexpr int PAD = K / 2;
    static constexpr int PADL = (PAD + 3) / 4 * 4;       // staged columns left of the tile (a box starts on 16 bytes)
    static constexpr int PADU = (PAD + 7) / 8 * 8;       // staged rows above a tile (whole row-residue groups)
    static constexpr int SKEW = PADL - PAD;              // first needed float of a thread's aligned row segment
    static constexpr int ODD = SKEW & 1;                 // first tap on an odd column
    static constexpr int P0 = SKEW >> 1;                 // pair that holds (ODD: in its high half) the first tap of output 0
    static constexpr int TYO = 8 / S;                    // LR rows per thread (8 HR rows)
    static constexpr int TXO = 16 / S;                   // LR columns per thread (16 HR columns)
    static constexpr int SEGF = SKEW + 16 - S + KW;      // floats of the aligned segment a thread needs per row
    static constexpr int NL4 = (SEGF + 3) / 4;           // LDS.128 per row
    static constexpr int NPL = 2 * NL4;                  // pixel pairs held per row
    static constexpr int NR = K + 7;                     // input rows a thread's outputs touch (even)
    static constexpr int NWP = TP + ODD;                 // weight pairs per composite row ([0, K'0] ... [K'last, 0] when ODD)
    static constexpr int NW4 = (NWP + 1) / 2;            // LDS.128 per composite row
    static constexpr int WP = 4 * NW4;                   // floats per staged composite row
    static constexpr int NSPLIT = TYO * TXO < 4 ? 2 : 1; // accumulator chains per output (factor 8: 2 outputs per thread)
    static constexpr int FULL0 = S * (TYO - 1);          // input rows [FULL0, KW) meet every output row of the thread
    static constexpr int RPAIR = 8 + PADL / 2;           // first pixel pair right of the band for its last thread
    static_assert(K % 2 == 1 && (S == 2 || S == 4 || S == 8), "odd kernel, factor 2 / 4 / 8");
    static_assert(NR % 2 == 0 && FULL0 % 2 == 0 && KW % 2 == 0, "row loops are unrolled by two");
    static_assert(RPAIR <= NPL, "a thread's segment ends at most one halo past its 16 columns");
};

struct BoxArgs {
    const float* hr;        // band mode: rows 0 and H-1 are also fetched with plain bulk copies
    long long sN, sC, sH;
    const float* comp;      // [nK, C, KW, KWp]
    int KWp;
    const float* dsum;
    const int* kidx;
    const float* sigma;
    const float* pool;
    const int* nidx;
    float* lr;
    long long nbands;
    int C, H, W, Ho, Wo;
    int tiles_x, tiles;     // 128 x 128 tiles per band (tile mode)
    int groups;             // band mode: bands (= warps) per CTA
    int NQ;                 // staged row groups of 8
    int WB;                 // staged row pitch in floats = box width (odd multiple of 4)
    int planeF;             // floats per row-residue plane: NQ WB rounded up to 128 bytes (a TMA destination)
    int regionF;            // floats per staged region (8 planes; band mode: + the two clamp rows, see clamp_floats)
    int clamp_rows;         // band mode: the two clamp rows are staged
    int replicate, noise_mode;
    int slot;               // ticket slot of this launch
    unsigned int ngroups_total;   // groups (CTAs, or warps in band mode) that draw tickets
};

// Band mode, replicate padding: rows above / below the band clamp to row 0 / H-1.  A lane that clamps would read another
// row-residue plane at the row group q its neighbour reads -- planes are 128-byte aligned, so that is a bank conflict on
// every load of 12 of a thread's 20 rows (measured: 35 % excess wavefronts on the 64-wide factor-8 cells, ncu r2d).
// Two extra staged rows behind the eight planes hold a copy of row H-1 at the bank offset of row group 0 and a copy of
// row 0 at the bank offset of row group 7: exactly the two bank groups the unclamped lanes of a quarter-warp leave free.
// clampF floats = [row H-1 | gap | row 0], the gap chosen so that row 0 starts at 7 WB floats modulo 32.
__host__ __device__ inline int clamp_top_offset(int WB) { return WB + ((7 * WB - WB) % 32 + 32) % 32; }
__host__ __device__ inline int clamp_floats(int WB) { return (clamp_top_offset(WB) + WB + 31) / 32 * 32; }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// Work distribution in band mode: the warps of persistent CTAs draw their bands from a ticket counter, so the hardware's
// dynamic balance between SMs is kept (a static stride lost 10 % to SM-to-SM speed differences, r2o) while a warp can
// fetch its next band's pixels during the epilogue of the current one.  One counter pair per launch slot (the host
// hands out slots round robin, so launches in flight on different streams do not share one); the last group to finish
// resets its slot.
constexpr int kTicketSlots = 64;
__device__ unsigned int g_box_next[kTicketSlots];
__device__ unsigned int g_box_done[kTicketSlots];

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
        "r"(bar)
        : "memory");
}

// BAND: every warp owns one whole band of at most 64 x 64 pixels (its own region, barrier and weights; only the
// image rows are staged, halo rows clamp or are zeroed in registers); otherwise the four warps share one 128 x 128
// tile (2 x 2 warps of 64 x 64) staged with its halo rows.
template <int K, int S, bool BAND>
__global__ void __launch_bounds__(kBoxThreads, (BAND && S != 2) ? 3 : 2)      // factor-2 bands spill at 168 registers (r2n)
degrade_box_kernel(const __grid_constant__ CUtensorMap tmap, const BoxArgs a) {
    using G = BCfg<K, S>;
    extern __shared__ __align__(128) unsigned char bsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ngroups = BAND ? a.groups : 1;                      // regions per CTA
    const int grp = BAND ? warp : 0;
    const int gthreads = BAND ? 32 : kBoxThreads;
    const int gtid = BAND ? lane : (int)threadIdx.x;
    uint64_t* bars = reinterpret_cast<uint64_t*>(bsm);
    float* region = reinterpret_cast<float*>(bsm + 128) + (size_t)grp * a.regionF;
    float* wsm = reinterpret_cast<float*>(bsm + 128) + (size_t)ngroups * a.regionF + (size_t)grp * (G::KW * G::WP);
    const uint32_t bar = smem_u32(bars + grp);

    if (threadIdx.x == 0) {
        for (int i = 0; i < ngroups; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- work items: tickets (see g_box_next); item -> (band, tile) ----
    const unsigned int nitems = (unsigned int)(BAND ? a.nbands : a.nbands * a.tiles);
    const bool clamp_rows = BAND && a.clamp_rows;
    const bool replicate = a.replicate != 0;
    const bool noisy = a.noise_mode != KMSR_NOISE_NONE;
    // thread (ly, lx): outputs of HR rows tileY0 + 8 ly .. +7, HR columns tileX0 + 16 lx .. +15
    const int lx = BAND ? (lane >> 3) : ((warp & 1) * 4 + (lane >> 3));
    const int ly = BAND ? (lane & 7) : ((warp >> 1) * 8 + (lane & 7));
    const float* tbase = region + 16 * lx;
    unsigned int* ticket_s = reinterpret_cast<unsigned int*>(bsm + 64);      // tile mode: the drawn ticket, broadcast
    auto group_sync = [&]() {
        if (BAND) __syncwarp();
        else __syncthreads();
    };
    // one thread of the group draws, everybody learns the result (the latency of the atomic is what the caller hides)
    auto draw_begin = [&]() -> unsigned int {
        unsigned int t = 0;
        if (gtid == 0) t = atomicAdd(&g_box_next[a.slot], 1u);
        return t;
    };
    auto draw_end = [&](unsigned int t) -> unsigned int {
        if (BAND) return __shfl_sync(0xffffffffu, t, 0);
        if (gtid == 0) *ticket_s = t;
        __syncthreads();
        const unsigned int r = *ticket_s;
        __syncthreads();
        return r;
    };
    auto finish = [&]() {                                  // the last group resets the slot for a later launch
        if (gtid == 0) {
            __threadfence();
            if (atomicAdd(&g_box_done[a.slot], 1u) == a.ngroups_total - 1) {
                g_box_next[a.slot] = 0;
                g_box_done[a.slot] = 0;
                __threadfence();
            }
        }
    };
    struct Item { long long band, n; int c, tileX0, tileY0; };
    auto decode = [&](unsigned int it) {
        Item w;
        w.tileX0 = 0; w.tileY0 = 0;
        unsigned int bnd = it;
        if (!BAND) {
            bnd = it / (unsigned int)a.tiles;
            const int t = (int)(it - bnd * (unsigned int)a.tiles);
            const int ty = t / a.tiles_x;
            w.tileY0 = 128 * ty; w.tileX0 = 128 * (t - ty * a.tiles_x);
        }
        const unsigned int nn = bnd / (unsigned int)a.C;
        w.band = bnd; w.n = nn;
        w.c = (int)(bnd - nn * (unsigned int)a.C);
        return w;
    };
    // TMA: one box per row residue j (rows j, j+8, ... of the staged window), 8 lanes issue in parallel; then the
    // composite kernel of the band (shifted by one float when the first tap is odd) by cp.async, so one global-memory
    // latency covers the whole copy.  Called for the first item up front and for every later one as soon as the row
    // loop of the previous item has read its last staged row: the fetch overlaps the epilogue.
    auto issue = [&](const Item& w) {
        const int rx0 = w.tileX0 - G::PADL;                        // HR column of staged column 0
        const int ry0 = BAND ? 0 : w.tileY0 - G::PADU;             // HR row of staged row 0 (a multiple of 8)
        if (BAND || warp == 0) {
            if (lane == 0) mbar_arrive_expect_tx(bar, 32u * (uint32_t)(a.NQ * a.WB) + (clamp_rows ? 8u * (uint32_t)a.W : 0u));
            __syncwarp();
            if (lane < 8)
                tma_load_5d(smem_u32(region + (size_t)lane * a.planeF), &tmap, rx0, lane, ry0 / 8, w.c, (int)w.n, bar);
            else if (clamp_rows && lane < 10) {
                const float* band0 = a.hr + w.n * a.sN + (long long)w.c * a.sC;
                const int top = lane == 8;               // row 0 -> bank offset of row group 7, row H-1 -> that of row group 0
                bulk_load(smem_u32(region + (size_t)8 * a.planeF + (top ? clamp_top_offset(a.WB) : 0) + G::PADL),
                          band0 + (top ? 0 : (long long)(a.H - 1) * a.sH), 4u * (uint32_t)a.W, bar);
            }
        }
        const int kid = a.kidx ? __ldg(a.kidx + w.n) : 0;
        const float* kc = a.comp + ((long long)kid * a.C + w.c) * (G::KW * a.KWp);
        const uint32_t wdst = smem_u32(wsm);
        if (!G::ODD && a.KWp == G::WP) {
            for (int e = gtid; e < G::KW * G::WP / 4; e += gthreads)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wdst + 16 * e), "l"(kc + 4 * e) : "memory");
        } else {
            for (int d = gtid; d < G::KW * G::WP; d += gthreads) {
                const int u = d / G::WP, v = d - u * G::WP - G::ODD;
                if (v >= 0 && v < G::KW)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(wdst + 4 * d), "l"(kc + u * a.KWp + v) : "memory");
                else
                    wsm[d] = 0.0f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Band mode is persistent (every warp draws bands until none is left: +2-7 %, r2t); tile mode is not -- one tile per
    // CTA, the hardware schedules: with tickets the four warps of a tile rendezvous before the epilogue and the kernel
    // measured 3-7 % slower at factor 2 / 4 (r2t; 9-12 % with a static stride, r2o).
    constexpr bool PERSIST = BAND;
    unsigned int item = PERSIST ? draw_end(draw_begin()) : blockIdx.x;
    if (PERSIST && item >= nitems) { finish(); return; }         // uniform over the group
    Item cur = decode(item);
    issue(cur);
    uint32_t phase = 0;
#pragma unroll 1
    for (;;) {
    const unsigned int next_raw = PERSIST ? draw_begin() : 0u;    // the ticket of the item after this one: in flight during the rows
    const long long band = cur.band, n = cur.n;
    const int c = cur.c, tileX0 = cur.tileX0, tileY0 = cur.tileY0;
    const int rx0 = tileX0 - G::PADL;
    const int ry0 = BAND ? 0 : tileY0 - G::PADU;
    const int kid = a.kidx ? __ldg(a.kidx + n) : 0;
    const float ds = __ldg(a.dsum + (long long)kid * a.C + c);
    const float scale = a.noise_mode == KMSR_NOISE_SIGMA ? __ldg(a.sigma + (long long)kid * a.C + c) : 1.0f;
    const int nid = noisy ? __ldg(a.nidx + n) : 0;
    const int hy0 = tileY0 + 8 * ly - G::PAD;                     // HR row of the thread's input row 0
    const int colsLeft = a.W - (tileX0 + 16 * lx);                // image columns from the thread's first column on
    const bool ledge = replicate && tileX0 + 16 * lx == 0;        // halo columns left of the band in this thread's segment
    const bool redge = replicate && colsLeft == 16;               // ... right of the band (W % 16 == 0: box_shape_ok)

    asm volatile("cp.async.wait_all;" ::: "memory");
    group_sync();                                                 // composite kernel staged
    mbar_wait(bar, phase);
    phase ^= 1;

    // pivot: the thread's own first pixel, clamped into the image
    float pv;
    {
        const int py = min(tileY0 + 8 * ly, a.H - 1) - ry0, px = min(tileX0 + 16 * lx, a.W - 1) - rx0;
        pv = region[(size_t)(py & 7) * a.planeF + (py >> 3) * a.WB + px];
        if (!isfinite(pv)) pv = 0.0f;
    }
    const u64 npv2 = pack2(-pv, -pv);

    u64 acc[G::NSPLIT][G::TYO][G::TXO];
#pragma unroll
    for (int sp = 0; sp < G::NSPLIT; ++sp)
#pragma unroll
        for (int yo = 0; yo < G::TYO; ++yo)
#pragma unroll
            for (int x = 0; x < G::TXO; ++x) acc[sp][yo][x] = 0ull;

    // input row rr of the thread's window -> pivot-shifted pixel pairs.  Replicate padding (C_30:107-109): rows clamp
    // by address, halo columns are substituted in the two edge threads.  Zero padding: the TMA fill already is the halo
    // (band mode stages image rows only: rows outside are zeroed here).
    auto load_row = [&](const int rr, u64 (&P)[G::NPL]) {
        const int hr_row = hy0 + rr;
        const int cl_row = min(max(hr_row, 0), a.H - 1);
        const int t = ((replicate || BAND) ? cl_row : hr_row) - ry0;
        KMSR_DASSERT(t >= 0 && t < 8 * a.NQ && 16 * lx + 4 * G::NL4 <= a.WB);     // the segment lies inside the staged region
        const float* pf = tbase + (size_t)(t & 7) * a.planeF + (t >> 3) * a.WB;
        if (clamp_rows) {                                 // band mode: clamped rows come from the conflict-free copies
            if (hr_row < 0) pf = tbase + (size_t)8 * a.planeF + clamp_top_offset(a.WB);
            if (hr_row >= a.H) pf = tbase + (size_t)8 * a.planeF;
        }
        const ulonglong2* prow = reinterpret_cast<const ulonglong2*>(pf);
#pragma unroll
        for (int i = 0; i < G::NL4; ++i) {
            const ulonglong2 v = prow[i];
            P[2 * i] = v.x; P[2 * i + 1] = v.y;
        }
        if (BAND && !replicate && hr_row != cl_row) {
#pragma unroll
            for (int i = 0; i < G::NPL; ++i) P[i] = 0ull;
        }
        // (a warp-uniform branch around these predicated moves measured 5-10 % slower on the large kernels, r2h)
        if (ledge) {
            const float v = lo2(P[G::PADL / 2]);
            const u64 vv = pack2(v, v);
#pragma unroll
            for (int i = 0; i < G::PADL / 2; ++i) P[i] = vv;
        }
        if (redge) {
            const float v = hi2(P[G::RPAIR - 1]);
            const u64 vv = pack2(v, v);
#pragma unroll
            for (int i = G::RPAIR; i < G::NPL; ++i) P[i] = vv;
        }
#pragma unroll
        for (int i = 0; i < G::NPL; ++i) P[i] = add2(P[i], npv2);
    };
    // one weight pair against the pixel pairs of all TXO outputs of row yo; pair index i of the staged composite row
    auto mac_pair = [&](const int yo, const int i, const u64 w, const u64 (&P)[G::NPL]) {
#pragma unroll
        for (int x = 0; x < G::TXO; ++x) {
            const int pb = G::P0 + (S * x) / 2 + i;
            u64& A = acc[G::NSPLIT == 2 ? (i & 1) : 0][yo][x];
            if (G::ODD && i == 0) A = pack2(lo2(A), fmaf(hi2(w), hi2(P[pb]), hi2(A)));          // first tap alone
            else if (G::ODD && i == G::NWP - 1) A = pack2(fmaf(lo2(w), lo2(P[pb]), lo2(A)), hi2(A));   // last tap alone
            else A = fma2(w, P[pb], A);
        }
    };
    // an input row that meets only some of the thread's output rows (top and bottom of the window)
    auto mac_partial = [&](const int rr, const u64 (&P)[G::NPL]) {
#pragma unroll
        for (int yo = 0; yo < G::TYO; ++yo) {
            const int u = rr - S * yo;                               // composite row this input row meets for output row yo
            if (u >= 0 && u < G::KW) {                               // uniform over the CTA
                const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(wsm + u * G::WP);
#pragma unroll
                for (int t4 = 0; t4 < G::NW4; ++t4) {
                    const ulonglong2 w2 = wrow[t4];
                    mac_pair(yo, 2 * t4, w2.x, P);
                    if (2 * t4 + 1 < G::NWP) mac_pair(yo, 2 * t4 + 1, w2.y, P);
                }
            }
        }
    };
    // an input row that meets all TYO output rows: straight line, tap-major over every accumulator
    auto mac_full = [&](const int rr, const u64 (&P)[G::NPL]) {
        const ulonglong2* wrow = reinterpret_cast<const ulonglong2*>(wsm + rr * G::WP);
#pragma unroll
        for (int t4 = 0; t4 < G::NW4; ++t4) {
            ulonglong2 w2[G::TYO];
#pragma unroll
            for (int yo = 0; yo < G::TYO; ++yo) w2[yo] = (wrow - yo * (S * G::WP / 4))[t4];
#pragma unroll
            for (int yo = 0; yo < G::TYO; ++yo) mac_pair(yo, 2 * t4, w2[yo].x, P);
            if (2 * t4 + 1 < G::NWP) {
#pragma unroll
                for (int yo = 0; yo < G::TYO; ++yo) mac_pair(yo, 2 * t4 + 1, w2[yo].y, P);
            }
        }
    };
    // Two rows per iteration, both loaded at the top: the second row's LDS.128 are in flight while the first row is
    // multiplied.  Nothing is carried across the back-edge -- a software pipeline that loaded row rr + 2 inside iteration
    // rr made ptxas copy ~70 registers per iteration at the back-edge (11 % of the issued instructions, ncu r2z).
    {
        u64 PA[G::NPL], PB[G::NPL];
        int rr = 0;
#pragma unroll 1
        for (; rr < G::FULL0; rr += 2) {
            load_row(rr, PA);
            load_row(rr + 1, PB);
            mac_partial(rr, PA);
            mac_partial(rr + 1, PB);
        }
#pragma unroll 1
        for (; rr < G::KW; rr += 2) {
            load_row(rr, PA);
            load_row(rr + 1, PB);
            mac_full(rr, PA);
            mac_full(rr + 1, PB);
        }
#pragma unroll 1
        for (; rr < G::NR; rr += 2) {
            load_row(rr, PA);
            load_row(rr + 1, PB);
            mac_partial(rr, PA);
            mac_partial(rr + 1, PB);
        }
    }

    // ---- every warp of the group has read its last staged row: fetch the next item while this one is written out ----
    const unsigned int next_item = PERSIST ? draw_end(next_raw) : nitems;
    if (PERSIST) group_sync();
    const bool has_next = PERSIST && next_item < nitems;
    Item nxt = cur;
    if (has_next) {
        nxt = decode(next_item);
        issue(nxt);
    }

    // ---- epilogue: even + odd taps, pivot back, noise, store ----
    const long long ohw = (long long)a.Ho * a.Wo;
    float* outb = a.lr + band * ohw;
    const float* nz = a.pool + ((long long)nid * a.C + c) * ohw;
    const int Xb = tileX0 / S + G::TXO * lx;
    constexpr int VEC = G::TXO >= 4 ? 4 : 2;
    const bool vec = (a.Wo % VEC == 0) && Xb + G::TXO <= a.Wo && (((uintptr_t)a.lr & 15) == 0) &&
                     (!noisy || ((uintptr_t)a.pool & 15) == 0);
#pragma unroll
    for (int yo = 0; yo < G::TYO; ++yo) {
        const int Y = tileY0 / S + G::TYO * ly + yo;
        if (Y >= a.Ho) continue;
        float res[G::TXO];
#pragma unroll
        for (int x = 0; x < G::TXO; ++x) {
            float sum = lo2(acc[0][yo][x]) + hi2(acc[0][yo][x]);
            if (G::NSPLIT == 2) sum += lo2(acc[G::NSPLIT - 1][yo][x]) + hi2(acc[G::NSPLIT - 1][yo][x]);
            res[x] = pv + fmaf(pv, ds, sum);
        }
        float* o = outb + (long long)Y * a.Wo + Xb;
        const float* z = nz + (long long)Y * a.Wo + Xb;
        if (vec && VEC == 4) {
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
        } else if (vec) {
            float2 r2 = make_float2(res[0], res[G::TXO > 1 ? 1 : 0]);
            if (noisy) {
                const float2 q = __ldg(reinterpret_cast<const float2*>(z));
                r2.x = fmaf(scale, q.x, r2.x); r2.y = fmaf(scale, q.y, r2.y);
            }
            *reinterpret_cast<float2*>(o) = r2;
        } else {
#pragma unroll
            for (int x = 0; x < G::TXO; ++x)
                if (Xb + x < a.Wo) o[x] = noisy ? fmaf(scale, __ldg(z + x), res[x]) : res[x];
        }
    }
    if (!has_next) break;
    cur = nxt;
    item = next_item;
    }   // ticket loop
    if (PERSIST) finish();
}

template <int K, int S>
int launch_box(const DegradeArgs& a, cudaStream_t st) {
    using G = BCfg<K, S>;
    EncodeTiledFn enc = get_tensor_map_encoder();
    KMSR_REQUIRE(enc != nullptr, KMSR_E_CUDA, "degrade (box): cuTensorMapEncodeTiled is not available from the driver");
    int dev = 0, max_smem = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const bool band_mode = a.H <= 64 && a.W <= 64;
    const int TW = band_mode ? 64 : 128;
    BoxArgs t;
    t.comp = a.comp; t.KWp = a.g.KWp; t.dsum = a.dsum; t.kidx = a.kidx; t.sigma = a.sigma; t.pool = a.pool; t.nidx = a.nidx;
    t.lr = a.lr; t.nbands = a.N * a.C; t.C = a.C; t.H = a.H; t.W = a.W; t.Ho = a.g.Ho; t.Wo = a.g.Wo;
    t.tiles_x = (a.W + 127) / 128;
    t.tiles = t.tiles_x * ((a.H + 127) / 128);
    t.NQ = band_mode ? 8 : (128 + 2 * G::PADU) / 8;
    // row pitch: the halo on both sides and the last thread's aligned LDS.128 run, rounded up to an odd multiple of 16 bytes
    int wb = G::PADL + TW + G::PAD;
    if (wb < TW - 16 + 4 * G::NL4) wb = TW - 16 + 4 * G::NL4;
    wb = (wb + 3) / 4 * 4;
    if ((wb / 4) % 2 == 0) wb += 4;
    t.WB = wb;
    t.planeF = (t.NQ * t.WB + 31) / 32 * 32;
    t.hr = a.hr; t.sN = a.N > 1 ? a.sN : 0; t.sC = a.sC; t.sH = a.sH;
    t.replicate = a.pad_mode == KMSR_PAD_REPLICATE ? 1 : 0;
    t.noise_mode = a.noise_mode;
    // band mode: the bands per CTA (1-4 warps) that keep the most warps resident (ties: larger CTAs).  The two clamp rows
    // (conflict-free replicate rows, see clamp_floats) are worth 10-27 % on the factor-4 / factor-8 bands; at factor 2
    // (FP32-bound) they are staged only when they do not cost a resident warp.
    t.groups = 1;
    t.clamp_rows = 0;
    size_t gbytes = (size_t)(8 * t.planeF + G::KW * G::WP) * 4;
    if (band_mode) {
        // resident warps per SM for g bands per CTA, as the runtime computes it (registers are allocated per scheduler:
        // at 202 registers a scheduler holds two warps, so 3-warp CTAs leave a quarter of the slots empty -- a plain
        // "register file / registers per warp" estimate picked exactly that for k = 11, ncu r2m)
        auto kern_b = degrade_box_kernel<K, S, true>;
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern_b, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern_b, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        auto plan = [&](size_t per_band, int* groups) {
            long long best = -1;
            for (int g = 4; g >= 1; --g) {
                const size_t bytes = 128 + g * per_band;
                if (bytes > (size_t)max_smem) continue;
                int ctas = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern_b, 32 * g, bytes) != cudaSuccess) continue;
                if ((long long)ctas * g > best) { best = (long long)ctas * g; *groups = g; }
            }
            return best;
        };
        int g0 = 1, g1 = 1;
        const long long plain = plan(gbytes, &g0);
        const size_t with_clamp = gbytes + (size_t)clamp_floats(t.WB) * 4;
        const long long clamped = t.replicate ? plan(with_clamp, &g1) : -1;
        KMSR_REQUIRE(plain > 0, KMSR_E_UNSUPPORTED, "degrade (box): k=%d factor=%d does not fit shared memory", K, S);
        if (clamped > 0 && (S > 2 || clamped >= plain)) { t.clamp_rows = 1; t.groups = g1; gbytes = with_clamp; }
        else t.groups = g0;
    }
    t.regionF = (int)(gbytes / 4) - G::KW * G::WP;
    const size_t smem = 128 + (size_t)t.groups * gbytes;
    KMSR_REQUIRE(smem <= (size_t)max_smem, KMSR_E_UNSUPPORTED, "degrade (box): k=%d factor=%d needs %zu B of shared memory", K, S, smem);

    // [N, C, H/8, 8, W] seen as dims (x, j = row mod 8, q = row div 8, c, n); box = WB x 1 x NQ: the rows of one residue
    CUtensorMap tmap;
    cuuint64_t gdim[5] = {(cuuint64_t)a.W, 8, (cuuint64_t)(a.H / 8), (cuuint64_t)a.C, (cuuint64_t)a.N};
    const long long sN = a.N > 1 ? a.sN : (long long)a.C * a.sC;
    cuuint64_t gstr[4] = {(cuuint64_t)a.sH * 4, (cuuint64_t)a.sH * 32, (cuuint64_t)a.sC * 4, (cuuint64_t)sN * 4};
    cuuint32_t box[5] = {(cuuint32_t)t.WB, 1, (cuuint32_t)t.NQ, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)a.hr, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    KMSR_REQUIRE(cr == CUDA_SUCCESS, KMSR_E_CUDA, "degrade (box): cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);

    // band mode: persistent CTAs, as many as are resident at once; bands come from the ticket counter of this launch's slot
    int sms = 0;
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long nitems = band_mode ? t.nbands : t.nbands * t.tiles;
    KMSR_REQUIRE(nitems < (1ll << 31), KMSR_E_INVALID, "degrade (box): too many tiles");
    const long long work_ctas = band_mode ? (t.nbands + t.groups - 1) / t.groups : nitems;
    static std::atomic<unsigned int> launch_seq{0};
    t.slot = (int)(launch_seq.fetch_add(1, std::memory_order_relaxed) % kTicketSlots);
    set_algo("box");
    int per_sm = 0;
    long long grid = 0;
    if (band_mode) {
        auto kern = degrade_box_kernel<K, S, true>;
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        KMSR_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * t.groups, smem));
        const long long resident = (long long)sms * (per_sm > 0 ? per_sm : 1);
        grid = work_ctas < resident ? work_ctas : resident;
        t.ngroups_total = (unsigned int)(grid * t.groups);
        kern<<<(unsigned)grid, 32 * t.groups, smem, st>>>(tmap, t);
    } else {
        auto kern = degrade_box_kernel<K, S, false>;
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KMSR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        grid = work_ctas;                                            // one tile per CTA
        t.ngroups_total = (unsigned int)grid;
        kern<<<(unsigned)grid, kBoxThreads, smem, st>>>(tmap, t);
    }
    KMSR_LAUNCH_CHECK("degrade_box_kernel");
    return KMSR_OK;
}

template <int K>
int launch_box_k(const DegradeArgs& a, cudaStream_t st) {
    switch (a.g.stride) {
        case 2: return launch_box<K, 2>(a, st);
        case 4: return launch_box<K, 4>(a, st);
        default: return launch_box<K, 8>(a, st);
    }
}

}  // namespace

bool box_shape_ok(const DegradeArgs& a, int down_mode, const char** why) {
    const Geometry& g = a.g;
    *why = "";
    if (down_mode != KMSR_DOWN_BOXMEAN) { *why = "box-mean downsampling only"; return false; }
    const int k = g.kh;
    if (g.kh != g.kw || !(k == 11 || k == 13 || k == 15 || k == 21 || k == 31)) { *why = "square kernels of size 11, 13, 15, 21, 31"; return false; }
    if (g.stride != 2 && g.stride != 4 && g.stride != 8) { *why = "factor 2, 4 or 8"; return false; }
    if (a.patch_offsets) { *why = "scene windows not covered"; return false; }
    if (a.stat_part) { *why = "fused statistics not covered"; return false; }
    if (a.H < 8 || a.H % 8 != 0) { *why = "needs H % 8 == 0"; return false; }
    if (a.W < 16 || a.W % 16 != 0) { *why = "needs W % 16 == 0"; return false; }
    if (((uintptr_t)a.hr & 15) || (a.sH & 3) || (a.sC & 3) || (a.N > 1 && (a.sN & 3))) { *why = "HR base / strides not 16-byte aligned"; return false; }
    if (a.sH < a.W || a.sC < 1 || (a.N > 1 && a.sN < 1)) { *why = "non-positive strides"; return false; }
    if (a.N >= (1ll << 31) || a.N * a.C >= (1ll << 31)) { *why = "too many patches"; return false; }
    return true;
}

int launch_degrade_box(const DegradeArgs& a, cudaStream_t st) {
    switch (a.g.kh) {
        case 11: return launch_box_k<11>(a, st);
        case 13: return launch_box_k<13>(a, st);
        case 15: return launch_box_k<15>(a, st);
        case 21: return launch_box_k<21>(a, st);
        default: return launch_box_k<31>(a, st);
    }
}

}  // namespace kmsr
