// scene.cu -- scene-scale helpers of A_00_patch_cutter_universal.py: water mask (:89-123) and the
// zero-NaN patch filter of create_patches_nc (:152-183).  Integer / compare work, HBM-bound.
#include "common.cuh"

namespace kmsr {

// data[c,i] == invalid -> NaN in place (CUT:102); masked = NaN where NIR not in [tmin,tmax] (CUT:108-113)
__global__ void __launch_bounds__(256)
water_mask_kernel(float* __restrict__ data, int C, long long hw, int nir, float invalid, float tmin, float tmax,
                  float* __restrict__ masked) {
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

// 128-bit form (hw % 4 == 0, 16-byte aligned bases, C <= 8): all band loads of a pixel quad are in flight
// together; the in-place write-back happens only where a fill value was actually replaced.
template <int C>
__global__ void __launch_bounds__(256)
water_mask_vec_kernel(float4* __restrict__ data, long long hw4, int nir, float invalid, float tmin,
                      float tmax, float4* __restrict__ masked) {
    const float qnan = __int_as_float(0x7fc00000);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw4;
         i += (long long)gridDim.x * blockDim.x) {
        float4 d[C];
#pragma unroll
        for (int c = 0; c < C; ++c) d[c] = data[(long long)c * hw4 + i];
        float4 n = d[0];
#pragma unroll
        for (int c = 1; c < C; ++c)
            if (c == nir) n = d[c];
        const bool w0 = (n.x != invalid) && (n.x >= tmin) && (n.x <= tmax);
        const bool w1 = (n.y != invalid) && (n.y >= tmin) && (n.y <= tmax);
        const bool w2 = (n.z != invalid) && (n.z >= tmin) && (n.z <= tmax);
        const bool w3 = (n.w != invalid) && (n.w >= tmin) && (n.w <= tmax);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            {
                float4 v = d[c];
                const bool ch = v.x == invalid || v.y == invalid || v.z == invalid || v.w == invalid;
                if (v.x == invalid) v.x = qnan;
                if (v.y == invalid) v.y = qnan;
                if (v.z == invalid) v.z = qnan;
                if (v.w == invalid) v.w = qnan;
                if (ch) data[(long long)c * hw4 + i] = v;
                masked[(long long)c * hw4 + i] = make_float4(w0 ? v.x : qnan, w1 ? v.y : qnan, w2 ? v.z : qnan, w3 ? v.w : qnan);
            }
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
    if ((W & 3) == 0 && (x0 & 3) == 0 && (ww & 3) == 0 && (((uintptr_t)masked) & 15) == 0) {
        // rows of the cell as float4: a warp covers 128 columns per load, four row loads in flight per thread
        const int w4 = ww >> 2, rows = C * hh, per_it = 256 / min(w4, 256);
        const int lx = threadIdx.x % w4, lr = threadIdx.x / w4;
        if (w4 <= 256 && lr < per_it) {
            for (int r0 = lr; r0 < rows; r0 += 4 * per_it) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u * per_it;
                    if (r < rows) {
                        const int c = r / hh, y = r - c * hh;
                        v[u] = *reinterpret_cast<const float4*>(masked + ((long long)c * H + y0 + y) * W + x0 + 4 * lx);
                    } else {
                        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    cnt += (v[u].x != v[u].x) + (v[u].y != v[u].y) + (v[u].z != v[u].z) + (v[u].w != v[u].w);
            }
        } else if (w4 > 256) {
            for (int r = 0; r < rows; ++r) {
                const int c = r / hh, y = r - c * hh;
                for (int x = threadIdx.x; x < w4; x += 256) {
                    const float4 t = *reinterpret_cast<const float4*>(masked + ((long long)c * H + y0 + y) * W + x0 + 4 * x);
                    cnt += (t.x != t.x) + (t.y != t.y) + (t.z != t.z) + (t.w != t.w);
                }
            }
        }
    } else {
        for (int c = 0; c < C; ++c) {
            const float* base = masked + ((long long)c * H + y0) * W + x0;
            for (int i = threadIdx.x; i < hh * ww; i += blockDim.x) {
                const float v = base[(long long)(i / ww) * W + (i % ww)];
                cnt += (v != v);
            }
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

// stage 1 without a materialised mask: the NaN count the masked scene WOULD have in every cell, straight from the
// raw scene -- a pixel of band c is NaN iff NIR is fill / NaN / outside [tmin, tmax] (CUT:108-113) or the band
// value itself is fill (CUT:102) or NaN.  With the reference's nan_threshold = 0 a kept window holds no masked
// pixel, so its pixels equal the raw scene's and the degrade kernel can read the raw scene directly.
__global__ void __launch_bounds__(256)
scene_cells_kernel(const float* __restrict__ data, int C, int H, int W, int nir, float invalid, float tmin,
                   float tmax, int cell, int ch, int cw, int* __restrict__ cells) {
    __shared__ int red[8];
    const int cj = blockIdx.x % cw, ci = blockIdx.x / cw;
    const int y0 = ci * cell, x0 = cj * cell;
    const int hh = min(cell, H - y0), ww = min(cell, W - x0);
    const long long plane = (long long)H * W;
    int cnt = 0;
    auto bad = [&](float n) { return !((n != invalid) && (n >= tmin) && (n <= tmax)); };
    if ((W & 3) == 0 && (x0 & 3) == 0 && (ww & 3) == 0 && (((uintptr_t)data) & 15) == 0 && (ww >> 2) <= 256) {
        const int w4 = ww >> 2, per_it = 256 / w4;
        const int lx = threadIdx.x % w4, lr = threadIdx.x / w4;
        if (lr < per_it) {
            for (int y = lr; y < hh; y += per_it) {
                const float* px = data + (long long)(y0 + y) * W + x0 + 4 * lx;
                const float4 n = *reinterpret_cast<const float4*>(px + (long long)nir * plane);
                const bool b0 = bad(n.x), b1 = bad(n.y), b2 = bad(n.z), b3 = bad(n.w);
                for (int c = 0; c < C; ++c) {
                    const float4 v = c == nir ? n : *reinterpret_cast<const float4*>(px + (long long)c * plane);
                    cnt += (b0 || v.x == invalid || v.x != v.x) + (b1 || v.y == invalid || v.y != v.y) +
                           (b2 || v.z == invalid || v.z != v.z) + (b3 || v.w == invalid || v.w != v.w);
                }
            }
        }
    } else {
        for (int i = threadIdx.x; i < hh * ww; i += blockDim.x) {
            const float* px = data + (long long)(y0 + i / ww) * W + x0 + (i % ww);
            const bool b = bad(px[(long long)nir * plane]);
            for (int c = 0; c < C; ++c) {
                const float v = px[(long long)c * plane];
                cnt += (b || v == invalid || v != v);
            }
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
    if ((hw & 3) == 0 && C <= 6 && (((uintptr_t)data | (uintptr_t)masked) & 15) == 0) {
        long long blocks = (hw / 4 + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        float4* d4 = (float4*)data;
        float4* m4 = (float4*)masked;
        const unsigned gb = (unsigned)blocks;
        switch (C) {
            case 1: water_mask_vec_kernel<1><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
            case 2: water_mask_vec_kernel<2><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
            case 3: water_mask_vec_kernel<3><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
            case 4: water_mask_vec_kernel<4><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
            case 5: water_mask_vec_kernel<5><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
            default: water_mask_vec_kernel<6><<<gb, 256, 0, st>>>(d4, hw / 4, nir, invalid, tmin, tmax, m4); break;
        }
        KMSR_LAUNCH_CHECK("water_mask_vec_kernel");
        return KMSR_OK;
    }
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

// CUT:179-183: drop iff nan_ratio > thr  <=>  keep iff count <= floor(thr * total) (exact in fp64 here)
static long long keep_limit(int C, int P, double thr) {
    const double total = (double)C * P * P;
    long long limit = thr >= 1.0 ? (long long)total : (long long)floor(thr * total);
    while (limit < (long long)total && (double)(limit + 1) / total <= thr) ++limit;
    while (limit >= 0 && (double)limit / total > thr) --limit;
    return limit;
}

int launch_scene_keep_mask(const float* data, int C, int H, int W, int nir, float invalid, float tmin, float tmax,
                           int P, int stride, double thr, unsigned char* keep, int* nan_count, void* ws,
                           long long ws_bytes, cudaStream_t st) {
    if (H < P || W < P) return KMSR_OK;
    KMSR_REQUIRE(P % stride == 0, KMSR_E_UNSUPPORTED,
                 "scene_keep_mask: patch size %d is not a multiple of the stride %d (use water_mask + keep_mask)", P, stride);
    const int hp = (H - P) / stride + 1, wp = (W - P) / stride + 1;
    const int ch = (H + stride - 1) / stride, cw = (W + stride - 1) / stride;
    KMSR_REQUIRE(ws && ws_bytes >= (long long)ch * cw * (long long)sizeof(int), KMSR_E_INVALID,
                 "scene_keep_mask: workspace of %lld B needed", (long long)ch * cw * (long long)sizeof(int));
    scene_cells_kernel<<<ch * cw, 256, 0, st>>>(data, C, H, W, nir, invalid, tmin, tmax, stride, ch, cw, (int*)ws);
    KMSR_LAUNCH_CHECK("scene_cells_kernel");
    keep_from_cells_kernel<<<(hp * wp + 127) / 128, 128, 0, st>>>((const int*)ws, cw, P / stride, hp, wp,
                                                                  keep_limit(C, P, thr), keep, nan_count);
    KMSR_LAUNCH_CHECK("keep_from_cells_kernel");
    return KMSR_OK;
}

int launch_keep_mask(const float* masked, int C, int H, int W, int P, int stride, double thr,
                     unsigned char* keep, int* nan_count, void* ws, long long ws_bytes,
                     cudaStream_t st) {
    if (H < P || W < P) return KMSR_OK;
    const int hp = (H - P) / stride + 1, wp = (W - P) / stride + 1;
    const long long limit = keep_limit(C, P, thr);
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
