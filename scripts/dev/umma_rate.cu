// Microbenchmark: issue rate of tcgen05.mma kind::f16 (A from TMEM or smem, B from smem, no swizzle)
// for N = 64 / 128 / 256, M = 128, K = 16.  Prints cycles per MMA and MAC/clk/SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
template <int N, bool A_TMEM>
__global__ void __launch_bounds__(128) rate(int iters, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (N * 128 * 2 + 128 * 128 * 2) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_s;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t b_base = smem_u32(sm), a_base = smem_u32(sm) + N * 128 * 2;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {   // K slab of 128 = 8 MMAs
                    uint64_t bdesc = (uint64_t)(((b_base + ks * 2 * (N * 16)) >> 4) & 0x3FFF) | ((uint64_t)((N * 16) >> 4) << 16) |
                                     ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
                    if (A_TMEM) {
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::
                                     "r"(tmem + 256), "r"(tmem + ks * 8), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        uint64_t adesc = (uint64_t)(((a_base + ks * 2 * (128 * 16)) >> 4) & 0x3FFF) | ((uint64_t)((128 * 16) >> 4) << 16) |
                                         ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::
                                     "r"(tmem + 256), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u) : "memory");
                    }
                }
            }
            __syncwarp();
        }
        if (elect_one())
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        __syncwarp();
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        t1 = clock64();
        if (tid == 0) out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int N, bool A_TMEM>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 148 * 8);
    const int iters = 2000;
    size_t smem = N * 128 * 2 + 128 * 128 * 2;
    cudaFuncSetAttribute(rate<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {1, 148}) {
        rate<N, A_TMEM><<<grid, 128, smem>>>(iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        double cyc = (double)h[0] / (iters * 8.0);
        printf("%s N=%d grid=%d: %.1f cycles/MMA -> %.0f MAC/clk/SM (%s)\n", name, N, grid, cyc, 128.0 * N * 16 / cyc, cudaGetErrorString(e));
    }
    cudaFree(d);
}
int main() {
    run<64, true>("A=TMEM"); run<128, true>("A=TMEM"); run<256, true>("A=TMEM");
    run<64, false>("A=SMEM"); run<128, false>("A=SMEM"); run<256, false>("A=SMEM");
    return 0;
}
