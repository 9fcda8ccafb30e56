// umma_ts_probe.cu -- A operand from tensor memory (written by tcgen05.st), B from shared memory; derived from umma_probe.cu: smallest tcgen05 program: D[128 x N] (TMEM) = A[128 x 16] * B[N x 16]^T, kind::tf32, both operands in
// shared memory in the K-major no-swizzle canonical layout (core matrix = 8 rows x 16 bytes, contiguous), written by
// threads; checks the descriptor encoding, the fences and the tcgen05.ld lane / column mapping against the host.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_probe scratch/umma_probe.cu && /tmp/umma_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 16;      // two k-steps of 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                  // descriptor version (Blackwell)
    return d;                                // base offset 0, layout type 0 = no swizzle
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, int* err) {
    __shared__ __align__(128) float sB[2 * 2 * (N / 8) * 8 * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < N)
        for (int k = 0; k < K; ++k)
            sB[(((k / 8) * 2 + (k / 4) % 2) * (N / 8) + tid / 8) * 32 + (tid % 8) * 4 + k % 4] = B[tid * K + k];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the async proxy
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    {   // A row r = tid -> TMEM lane tid, columns 64 .. 79 (one 32-bit column per K element)
        uint32_t v[16];
        for (int k = 0; k < 16; ++k) v[k] = __float_as_uint(A[tid * K + k]);
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + 64;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ta = tm + 64 + ks * 8;
            const uint64_t db = make_desc(smem_u32(sB) + ks * (2 * (N / 8) * 128), (N / 8) * 128, 128);
            const uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tm), "r"(ta), "l"(db), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    bool ok = false;
    for (int spin = 0; spin < 2000000 && !ok; ++spin) ok = mbar_try(smem_u32(&bar), 0);
    if (!ok) { if (tid == 0) *err = 1; }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (ok) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128));
}

int main() {
    float hA[M * K], hB[N * K], hD[M * N], ref[M * N];
    for (int i = 0; i < M * K; ++i) hA[i] = (float)((i * 7 + 3) % 13 - 6);
    for (int i = 0; i < N * K; ++i) hB[i] = (float)((i * 5 + 1) % 11 - 5) * 0.5f;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += hA[m * K + k] * hB[n * K + k]; ref[m * N + n] = s; }
    float *dA, *dB, *dD; int* dE; int hE = 0;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD); cudaMalloc(&dE, 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, sizeof hD); cudaMemset(dE, 0, 4);
    probe<<<1, 128>>>(dA, dB, dD, dE);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost); cudaMemcpy(&hE, dE, 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int i = 0; i < M * N; ++i) if (hD[i] != ref[i]) { if (bad < 8) printf("mismatch at (%d, %d): got %g want %g\n", i / N, i % N, hD[i], ref[i]); ++bad; }
    printf("cuda: %s, barrier timeout flag %d, mismatches %d of %d\n", cudaGetErrorString(e), hE, bad, M * N);
    return bad != 0 || hE != 0 || e != cudaSuccess;
}
