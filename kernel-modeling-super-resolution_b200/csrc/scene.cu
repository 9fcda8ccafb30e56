// scene.cu -- scene-scale helpers of A_00_patch_cutter_universal.py: water mask (:89-123) and the
// zero-NaN patch filter of create_patches_nc (:152-183).  Integer / compare work, HBM-bound.
#include "common.cuh"

namespace kmsr {

// data[c,i] == invalid -> NaN in place (CUT:102); masked = NaN where NIR not in [tmin,tmax] (CUT:108-113)
__global__ void __launch_bounds__(256)
water_mask_kernel(float* __restrict__ data, int C, long long hw, int nir, float invalid, float tmin,
                  float tmax, float* __restrict__ masked) {
    const float qnan = __int_as_float(0x7fc00000);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw;
         i += (long long)gridDim.x * blockDim.x) {
        float v = data[(long long)nir * hw + i];
        if (v == invalid) v = qnan;
        const bool water = (v >= tmin) && (v <= tmax);     // NaN compares false
        for (int c = 0; c < C; ++c) {
            float d = data[(long long)c * hw + i];
            if (d == invalid) d = qnan;
            data[(long long)c * hw + i] = d;
            masked[(long long)c * hw + i] = water ? d : qnan;
        }
    }
}

// stage 1: NaN count of every stride x stride cell over all bands -> cells [ch, cw] int32
__global__ void __launch_bounds__(256)
nan_cells_kernel(const float* __restrict__ masked, int C, int H, int W, int cell, int ch, int cw,
                 int* __restrict__ cells) {
    __shared__ int red[8];
    const int cj = blockIdx.x % cw, ci = blockIdx.x / cw;
    const int y0 = ci * cell, x0 = cj * cell;
    const int hh = min(cell, H - y0), ww = min(cell, W - x0);
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
        const float* base = masked + ((long long)c * H + y0) * W + x0;
        for (int i = threadIdx.x; i < hh * ww; i += blockDim.x) {
            const float v = base[(long long)(i / ww) * W + (i % ww)];
            cnt += (v != v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        cells[blockIdx.x] = t;
    }
}

// stage 2: window (i,j) covers (P/cell)^2 cells
__global__ void keep_from_cells_kernel(const int* __restrict__ cells, int cw, int per, int hp, int wp,
                                       long long limit, unsigned char* __restrict__ keep,
                                       int* __restrict__ nan_count) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hp * wp) return;
    const int i = idx / wp, j = idx % wp;
    long long t = 0;
    for (int a = 0; a < per; ++a)
        for (int b = 0; b < per; ++b) t += cells[(i + a) * cw + (j + b)];
    keep[idx] = t <= limit ? 1 : 0;
    if (nan_count) nan_count[idx] = (int)t;
}

// general fallback (P not a multiple of stride): one CTA per window
__global__ void __launch_bounds__(256)
keep_direct_kernel(const float* __restrict__ masked, int C, int H, int W, int P, int stride, int wp,
                   long long limit, unsigned char* __restrict__ keep, int* __restrict__ nan_count) {
    __shared__ int red[8];
    const int i = blockIdx.x / wp, j = blockIdx.x % wp;
    int cnt = 0;
    for (int c = 0; c < C; ++c) {
        const float* base = masked + ((long long)c * H + (long long)i * stride) * W + (long long)j * stride;
        for (int e = threadIdx.x; e < P * P; e += blockDim.x) {
            const float v = base[(long long)(e / P) * W + (e % P)];
            cnt += (v != v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < 8; ++w) t += red[w];
        keep[blockIdx.x] = t <= limit ? 1 : 0;
        if (nan_count) nan_count[blockIdx.x] = (int)t;
    }
}

int launch_water_mask(float* data, int C, long long hw, int nir, float invalid, float tmin, float tmax,
                      float* masked, cudaStream_t st) {
    if (hw == 0) return KMSR_OK;
    long long blocks = (hw + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    water_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(data, C, hw, nir, invalid, tmin, tmax, masked);
    KMSR_LAUNCH_CHECK("water_mask_kernel");
    return KMSR_OK;
}

long long keep_mask_workspace(int H, int W, int P, int stride) {
    if (stride <= 0 || P <= 0 || P % stride != 0) return 0;
    const long long ch = (H + stride - 1) / stride, cw = (W + stride - 1) / stride;
    return ch * cw * (long long)sizeof(int);
}

int launch_keep_mask(const float* masked, int C, int H, int W, int P, int stride, double thr,
                     unsigned char* keep, int* nan_count, void* ws, long long ws_bytes,
                     cudaStream_t st) {
    if (H < P || W < P) return KMSR_OK;
    const int hp = (H - P) / stride + 1, wp = (W - P) / stride + 1;
    // CUT:179-183: drop iff nan_ratio > thr  <=>  keep iff count <= floor(thr * total) (exact in fp64 here)
    const double total = (double)C * P * P;
    long long limit = thr >= 1.0 ? (long long)total : (long long)floor(thr * total);
    while (limit < (long long)total && (double)(limit + 1) / total <= thr) ++limit;
    while (limit >= 0 && (double)limit / total > thr) --limit;
    if (P % stride == 0) {
        const int ch = (H + stride - 1) / stride, cw = (W + stride - 1) / stride;
        KMSR_REQUIRE(ws && ws_bytes >= (long long)ch * cw * (long long)sizeof(int), KMSR_E_INVALID,
                     "keep_mask: workspace of %lld B needed", (long long)ch * cw * (long long)sizeof(int));
        nan_cells_kernel<<<ch * cw, 256, 0, st>>>(masked, C, H, W, stride, ch, cw, (int*)ws);
        KMSR_LAUNCH_CHECK("nan_cells_kernel");
        keep_from_cells_kernel<<<(hp * wp + 127) / 128, 128, 0, st>>>((const int*)ws, cw, P / stride, hp, wp,
                                                                      limit, keep, nan_count);
        KMSR_LAUNCH_CHECK("keep_from_cells_kernel");
    } else {
        keep_direct_kernel<<<hp * wp, 256, 0, st>>>(masked, C, H, W, P, stride, wp, limit, keep, nan_count);
        KMSR_LAUNCH_CHECK("keep_direct_kernel");
    }
    return KMSR_OK;
}

}  // namespace kmsr
