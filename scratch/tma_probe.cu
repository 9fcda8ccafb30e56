// tma_probe: one TMA tile load into smem, dumped to global, to bisect descriptor constraints.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, int w, int bytes, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*)(smem + 65536);
    uint32_t b = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(dst), "l"(&tmap), "r"(x), "r"(y), "r"(z), "r"(w), "r"(b) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(b) : "memory");
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((float*)smem)[i];
}

int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int N = 3, C = 5, H = 256, W = 256;
    size_t n = (size_t)N * C * H * W;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 65536);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap tm;
    CUresult cr;
    int x = 0, y = 0, bytes = 0, rowf = 0;
    if (mode == 0) {          // u64, box 138 x 8, x=-3, y=-6
        cuuint64_t gd[4] = {W / 2, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {138, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = -3; y = -6; rowf = 276; bytes = 276 * 8 * 4;
    } else if (mode == 1) {   // f32, box 256 x 8, x=0
        cuuint64_t gd[4] = {W, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {256, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = 0; y = 8; rowf = 256; bytes = 256 * 8 * 4;
    } else if (mode == 2) {   // u64, box 128 x 8, x=0
        cuuint64_t gd[4] = {W / 2, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {128, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = 0; y = 8; rowf = 256; bytes = 256 * 8 * 4;
    } else if (mode == 3) {   // f32, box 140 x 8, x=-6 (two such boxes would cover a 268-wide row)
        cuuint64_t gd[4] = {W, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {140, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = -6; y = -6; rowf = 140; bytes = 140 * 8 * 4;
    } else if (mode == 4) {   // u64, box 138 x 8, x=0 y=8 (no OOB start)
        cuuint64_t gd[4] = {W / 2, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {138, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = 0; y = 8; rowf = 276; bytes = 276 * 8 * 4;
    } else if (mode == 6 || mode == 7) {   // u64, box 138 x 8, x=-4 (16-byte aligned start), y=-6 / 8
        cuuint64_t gd[4] = {W / 2, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {138, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = -4; y = mode == 6 ? -6 : 8; rowf = 276; bytes = 276 * 8 * 4;
    } else if (mode == 8) {   // u64, box 138 x 8, x=-4, y=250 (bottom OOB)
        cuuint64_t gd[4] = {W / 2, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {138, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = -4; y = 250; rowf = 276; bytes = 276 * 8 * 4;
    } else {                  // int32 (typeless 4-byte), box 256, x = -6: box max inner 256 elements
        cuuint64_t gd[4] = {W, H, C, N}; cuuint64_t gs[3] = {W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
        cuuint32_t bx[4] = {256, 8, 1, 1}, es[4] = {1, 1, 1, 1};
        cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 4, d, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        x = -6; y = -6; rowf = 256; bytes = 256 * 8 * 4;
    }
    printf("mode %d encode rc=%d\n", mode, (int)cr);
    if (cr != CUDA_SUCCESS) return 2;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
    probe<<<1, 128, 65536 + 64>>>(tm, x, y, 1, 2, bytes, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %d kernel: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 3;
    std::vector<float> r(bytes / 4);
    cudaMemcpy(r.data(), o, bytes, cudaMemcpyDeviceToHost);
    // check against expectation: smem[row][col] = img[n=2][c=1][y+row][xf + col] or 0 when OOB
    int xf = (mode == 0 || mode == 2 || mode == 4 || mode == 6 || mode == 7 || mode == 8) ? 2 * x : x;
    int bad = 0;
    for (int row = 0; row < 8; ++row)
        for (int col = 0; col < rowf; ++col) {
            int gy = y + row, gx = xf + col;
            float want = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) want = h[(((size_t)2 * C + 1) * H + gy) * W + gx];
            if (r[row * rowf + col] != want) { if (bad < 5) printf("  mismatch row %d col %d got %f want %f\n", row, col, r[row * rowf + col], want); ++bad; }
        }
    printf("mode %d mismatches=%d\n", mode, bad);
    return bad ? 4 : 0;
}
