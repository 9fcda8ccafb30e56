// noise.cu -- noise-pool gather + add (E_make_train_data.py:65-74) and noise-pool construction
// (D_build_noise_pool.py:88 `noise = geo - den`, :41-53 random_crop).  All indices / offsets are
// drawn on the host so that the MT19937 streams are the reference's bit for bit.
#include "common.cuh"

namespace kmsr {

// out[n,c,i] = blurred[n,c,i] + scale(n,c) * pool[nidx[n],c,i]; one CTA strip per (n, c)
template <bool VEC4>
__global__ void __launch_bounds__(256)
add_noise_kernel(const float* __restrict__ blurred, const float* __restrict__ pool,
                 const int* __restrict__ nidx, const float* __restrict__ sigma,
                 const int* __restrict__ kidx, float* __restrict__ out, int C, long long hw) {
    const long long band = blockIdx.x;
    const long long n = band / C;
    const int c = (int)(band % C);
    const float scale = sigma ? sigma[(long long)(kidx ? kidx[n] : 0) * C + c] : 1.0f;
    const float* b = blurred + band * hw;
    const float* z = pool + ((long long)nidx[n] * C + c) * hw;
    float* o = out + band * hw;
    if (VEC4) {
        const long long q = hw >> 2;
        for (long long i = threadIdx.x + (long long)blockIdx.y * blockDim.x; i < q;
             i += (long long)blockDim.x * gridDim.y) {
            const float4 bv = reinterpret_cast<const float4*>(b)[i];
            const float4 zv = __ldg(reinterpret_cast<const float4*>(z) + i);
            float4 r;
            r.x = fmaf(scale, zv.x, bv.x); r.y = fmaf(scale, zv.y, bv.y);
            r.z = fmaf(scale, zv.z, bv.z); r.w = fmaf(scale, zv.w, bv.w);
            reinterpret_cast<float4*>(o)[i] = r;
        }
    } else {
        for (long long i = threadIdx.x + (long long)blockIdx.y * blockDim.x; i < hw;
             i += (long long)blockDim.x * gridDim.y)
            o[i] = fmaf(scale, __ldg(z + i), b[i]);     // scale == 1: exactly b + z
    }
}

// pool[m,c,y,x] = geo[c,top[m]+y,left[m]+x] - den[...]; one CTA per (sample, band)
__global__ void __launch_bounds__(256)
crop_sub_kernel(const float* __restrict__ geo, const float* __restrict__ den, int C, int H, int W,
                const int* __restrict__ top, const int* __restrict__ left, int crop,
                float* __restrict__ pool) {
    const long long m = blockIdx.x / C;
    const int c = (int)(blockIdx.x % C);
    const int t = top[m], l = left[m];
    const long long src0 = ((long long)c * H + t) * W + l;
    float* o = pool + ((long long)m * C + c) * crop * crop;
    for (int i = threadIdx.x; i < crop * crop; i += blockDim.x) {
        const int y = i / crop, x = i % crop;
        const long long s = src0 + (long long)y * W + x;
        o[i] = __ldg(geo + s) - __ldg(den + s);
    }
}

// counts the entries of idx outside [0, upper): the opt-in range check for device-resident indices
__global__ void __launch_bounds__(256)
count_out_of_range_kernel(const int* __restrict__ idx, long long n, long long upper, int* __restrict__ bad) {
    int mine = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = idx[i];
        mine += (v < 0 || v >= upper) ? 1 : 0;
    }
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(bad, mine);
}

int launch_validate_indices(const int* idx, long long n, long long upper, int* bad, int* bad_host, cudaStream_t st) {
    *bad_host = 0;
    if (n == 0) return KMSR_OK;
    KMSR_CUDA_OK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    long long blocks = (n + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    count_out_of_range_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, n, upper, bad);
    KMSR_LAUNCH_CHECK("count_out_of_range_kernel");
    KMSR_CUDA_OK(cudaMemcpyAsync(bad_host, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    KMSR_CUDA_OK(cudaStreamSynchronize(st));
    return KMSR_OK;
}

int launch_add_noise(const float* blurred, long long N, int C, long long hw, const float* pool,
                     const int* nidx, const float* sigma, const int* kidx, float* out,
                     cudaStream_t st) {
    const long long bands = N * C;
    if (bands == 0 || hw == 0) return KMSR_OK;
    KMSR_REQUIRE(bands < (1ll << 31), KMSR_E_INVALID, "add_noise: N*C too large");
    const bool vec = hw % 4 == 0 && ((uintptr_t)blurred % 16 == 0) && ((uintptr_t)pool % 16 == 0) &&
                     ((uintptr_t)out % 16 == 0);
    const long long per = vec ? hw / 4 : hw;
    const unsigned gy = (unsigned)((per + 256 * 4 - 1) / (256 * 4) > 64 ? 64 : (per + 256 * 4 - 1) / (256 * 4));
    dim3 grid((unsigned)bands, gy ? gy : 1);
    if (vec) add_noise_kernel<true><<<grid, 256, 0, st>>>(blurred, pool, nidx, sigma, kidx, out, C, hw);
    else add_noise_kernel<false><<<grid, 256, 0, st>>>(blurred, pool, nidx, sigma, kidx, out, C, hw);
    KMSR_LAUNCH_CHECK("add_noise_kernel");
    return KMSR_OK;
}

int launch_crop_sub(const float* geo, const float* den, int C, int H, int W, const int* top,
                    const int* left, long long n, int crop, float* pool, cudaStream_t st) {
    if (n == 0) return KMSR_OK;
    KMSR_REQUIRE(n * C < (1ll << 31), KMSR_E_INVALID, "crop_sub: too many samples");
    crop_sub_kernel<<<(unsigned)(n * C), 256, 0, st>>>(geo, den, C, H, W, top, left, crop, pool);
    KMSR_LAUNCH_CHECK("crop_sub_kernel");
    return KMSR_OK;
}

}  // namespace kmsr
