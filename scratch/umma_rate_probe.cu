// umma_rate_probe.cu -- time per tcgen05.mma.kind::tf32 (M = 128, K = 8) as a function of N and of where A comes from
// (shared memory descriptor / tensor memory), issued back to back by one thread of one CTA per SM; and with two issuing threads.
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/urate scratch/umma_rate_probe.cu && /tmp/urate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | (uint64_t)((lbo >> 4) & 0x3FFF) << 16 | (uint64_t)((sbo >> 4) & 0x3FFF) << 32 | (uint64_t)1 << 46;
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

template <int N, bool TS, int ISSUERS>
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (4096 + N * 32) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 1.0f;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;
    const int me = tid >> 5;                                   // issuer = lane 0 of warps 0 .. ISSUERS-1
    if ((tid & 31) == 0 && me < ISSUERS) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(sm), 2048, 128);
        const uint64_t db = make_desc(smem_u32(sm) + 4096, (N / 8) * 128, 128);
        const uint32_t d = tm + me * 256, ta = tm + 480;
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if (TS) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db), "r"(idesc), "r"(1) : "memory");
            else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(1) : "memory");
        }
        const long long t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[me])) : "memory");
        while (!mbar_try(smem_u32(&bar[me]), 0)) {}
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[2 * me] = t1 - t0; out[2 * me + 1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

template <int N, bool TS, int ISSUERS>
void run(long long* d_out) {
    const int iters = 4096;
    auto k = rate<N, TS, ISSUERS>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<148, 128, 65536>>>(d_out, iters);
    cudaDeviceSynchronize();
    k<<<148, 128, 65536>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[4] = {0, 0, 0, 0};
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("N = %3d  A from %s  issuers %d : issue %.1f clk / MMA, complete %.1f clk / MMA (per issuer; %d MMAs each)  %s\n", N, TS ? "TMEM" : "smem",
           ISSUERS, (double)h[0] / iters, (double)h[1] / iters, iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d;
    cudaMalloc(&d, 64);
    run<32, false, 1>(d); run<64, false, 1>(d); run<128, false, 1>(d); run<256, false, 1>(d);
    run<32, true, 1>(d); run<64, true, 1>(d); run<128, true, 1>(d); run<256, true, 1>(d);
    run<64, true, 2>(d); run<128, true, 2>(d); run<64, false, 2>(d); run<128, false, 2>(d);
    return 0;
}
