// Microbenchmark: tcgen05.mma.cta_group::2 rate (M=256 over a CTA pair, A from TMEM, B halves in both CTAs' smem)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) rate(int iters, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    constexpr int NHALF = N / 2;
    for (int i = tid; i < (NHALF * 128 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (16u << 24);
    if (warp == 0) {
        long long t0 = clock64();
        if (rank == 0) {
            const uint32_t b_base = smem_u32(sm);
            for (int it = 0; it < iters; ++it) {
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        uint64_t bdesc = (uint64_t)(((b_base + ks * 2 * (NHALF * 16)) >> 4) & 0x3FFF) | ((uint64_t)((NHALF * 16) >> 4) << 16) |
                                         ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::
                                     "r"(tmem + 256), "r"(tmem + ks * 8), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                    }
                }
                __syncwarp();
            }
            if (elect_one())
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                             ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
            __syncwarp();
        }
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        long long t1 = clock64();
        if (tid == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int N>
void run() {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int iters = 2000;
    size_t smem = (N / 2) * 128 * 2;
    cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {2, 148}) {
        rate<N><<<grid, 128, smem>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        double cyc = (double)h[0] / (iters * 8.0);
        printf("cta_group::2 M=256 N=%d grid=%d: %.1f cycles/MMA -> %.0f MAC/clk/SM (%s)\n", N, grid, cyc, 256.0 * N * 16 / cyc / 2, cudaGetErrorString(e));
    }
    cudaFree(d);
}
int main() { run<64>(); run<128>(); run<256>(); return 0; }
