// common.cuh -- shared host/device helpers of libkmsr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/kmsr.h"

// Debug build (make DEBUG=1 -> libkmsr_debug.so, -DKMSR_DEBUG_ASSERTS): device-side asserts on ring slots, staged rows
// and noise-tile indices of the TMA kernels.  compute-sanitizer is closed on this pool; the GPU parity suite is run once
// per round against this library instead (KMSR_LIB=.../libkmsr_debug.so).  The release build compiles them away.
#ifdef KMSR_DEBUG_ASSERTS
#include <assert.h>
#define KMSR_DASSERT(cond) assert(cond)
#else
#define KMSR_DASSERT(cond) ((void)0)
#endif

namespace kmsr {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
void set_algo(const char* name);

inline int ilog2_floor(int v) {
    int s = 0;
    while ((1 << (s + 1)) <= v) ++s;
    return s;
}

struct Geometry {
    int kh, kw;          // blur kernel
    int KH, KW, KWp;     // composite window, row pitch (multiple of 4 floats)
    int stride;          // output stride in HR pixels
    int pt, pl;          // top / left halo (kh/2, kw/2)
    int Hb, Wb;          // blurred size
    int Ho, Wo;          // output size
};

// Host-side shape algebra shared by every entry point (C_30:107-122, train_gemini.py:128-134).
inline int make_geometry(int H, int W, int kh, int kw, int factor, int down_mode, Geometry* g) {
    if (kh < 1 || kw < 1 || factor < 1 || H < 0 || W < 0) return KMSR_E_INVALID;
    g->kh = kh; g->kw = kw;
    g->pt = kh / 2; g->pl = kw / 2;
    g->Hb = H + 2 * g->pt - kh + 1;
    g->Wb = W + 2 * g->pl - kw + 1;
    if (down_mode == KMSR_DOWN_BOXMEAN) {
        int f = 1 << ilog2_floor(factor);
        g->stride = f;
        g->KH = kh + f - 1; g->KW = kw + f - 1;
        g->Ho = g->Hb / f; g->Wo = g->Wb / f;
    } else if (down_mode == KMSR_DOWN_DECIMATE) {
        g->stride = factor;
        g->KH = kh; g->KW = kw;
        g->Ho = (g->Hb + factor - 1) / factor; g->Wo = (g->Wb + factor - 1) / factor;
    } else {
        return KMSR_E_INVALID;
    }
    g->KWp = (g->KW + 3) & ~3;
    if (g->Hb < 0) g->Hb = 0;
    if (g->Wb < 0) g->Wb = 0;
    if (g->Ho < 0) g->Ho = 0;
    if (g->Wo < 0) g->Wo = 0;
    return KMSR_OK;
}

#define KMSR_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::kmsr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                              __FILE__, __LINE__);                                      \
            return KMSR_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define KMSR_LAUNCH_CHECK(name)                                                         \
    do {                                                                                \
        ::kmsr::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            ::kmsr::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e)); \
            return KMSR_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define KMSR_REQUIRE(cond, code, ...)          \
    do {                                       \
        if (!(cond)) {                         \
            ::kmsr::set_error(__VA_ARGS__);    \
            return (code);                     \
        }                                      \
    } while (0)

// kernels implemented in the sibling .cu files
struct DegradeArgs {
    const float* hr;
    long long N;
    int C, H, W;
    long long sN, sC, sH;
    const long long* patch_offsets;
    int scene_h, scene_w, x_multiple;   // extents of the tensor the windows live in (0 = unknown)
    const float* comp;   // [nK, C, KH, KWp]
    const float* dsum;   // [nK, C]
    long long nK;
    const int* kidx;
    const float* sigma;
    const float* pool;
    long long nPool;
    const int* nidx;
    int pad_mode, noise_mode;
    float* lr;
    double* stat_part;   // fused statistics partials [N*C][2][2] (TMA kernel only), or nullptr
    Geometry g;
};

int launch_degrade_tiled(const DegradeArgs& a, cudaStream_t st);
bool tma_shape_ok(const DegradeArgs& a, const char** why);
int launch_degrade_tma(const DegradeArgs& a, cudaStream_t st);
bool stream_shape_ok(const DegradeArgs& a, int down_mode, const char** why);
int launch_degrade_stream(const DegradeArgs& a, cudaStream_t st);
bool reg_shape_ok(const DegradeArgs& a, int down_mode, const char** why);
int launch_degrade_reg(const DegradeArgs& a, cudaStream_t st);
bool box_shape_ok(const DegradeArgs& a, int down_mode, const char** why);
int launch_degrade_box(const DegradeArgs& a, cudaStream_t st);

}  // namespace kmsr
