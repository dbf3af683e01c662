// Microbenchmark: latency of small tcgen05.mma.cta_group::2 groups (M=256 over a CTA pair, N=128, K=16)
// from issue to the multicast commit becoming visible, A from shared memory (SS) or TMEM (TS),
// one commit per MMA or one per group.  nvcc -gencode arch=compute_100a,code=sm_100a -o umma2_lat umma2_lat.cu
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
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t rows) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((rows * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// MODE 0: SS, commit per MMA; 1: SS, one commit; 2: TS, commit per MMA; 3: TS one commit
// NM MMAs per group (different D columns), K16 each; KS = MMAs (K steps) per D
template <int MODE, int NM, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) lat(int reps, long long* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[8];
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < (64 * 1024) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0;
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
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
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (16u << 24);
    long long acc_issue = 0, acc_first = 0, acc_last = 0;
    if (warp == 0) {
        const uint32_t a_base = smem_u32(sm), b_base = smem_u32(sm) + 32768;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64(), t1 = 0;
            if (rank == 0) {
                if (elect_one()) {
#pragma unroll
                    for (int m = 0; m < NM; ++m) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) {
                            const uint64_t bd = desc(b_base + m * 4096 + ks * 2 * (64 * 16), 64);
                            if (MODE < 2) {
                                const uint64_t ad = desc(a_base + ks * 2 * (128 * 16), 128);
                                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::
                                             "r"(tmem + 128 * (m & 3)), "l"(ad), "l"(bd), "r"(idesc), "r"((uint32_t)ks) : "memory");
                            } else {
                                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::
                                             "r"(tmem + 256 + 128 * (m & 1)), "r"(tmem + ks * 8), "l"(bd), "r"(idesc), "r"((uint32_t)ks) : "memory");
                            }
                        }
                        if ((MODE & 1) == 0 || m == NM - 1)
                            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                         ::"r"(smem_u32(&bar[(MODE & 1) ? NM - 1 : m])), "h"((uint16_t)3) : "memory");
                    }
                }
                __syncwarp();
                t1 = clock64();
            }
            long long tf = 0, tl = 0;
            if ((MODE & 1) == 0) {
                wait_bar(&bar[0], r & 1);
                tf = clock64();
                for (int m = 1; m < NM; ++m) wait_bar(&bar[m], r & 1);
                tl = clock64();
            } else {
                wait_bar(&bar[NM - 1], r & 1);
                tf = tl = clock64();
            }
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (r > 0) { acc_issue += t1 - t0; acc_first += tf - t0; acc_last += tl - t0; }
            // keep the pair in lock-step between repetitions
            __syncwarp();
            cluster_sync_all();
        }
        if (tid == 0) { out[blockIdx.x * 3] = acc_issue; out[blockIdx.x * 3 + 1] = acc_first; out[blockIdx.x * 3 + 2] = acc_last; }
    } else {
        for (int r = 0; r < reps; ++r) cluster_sync_all();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int MODE, int NM, int KS>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 148 * 3 * 8);
    const int reps = 201;
    size_t smem = 64 * 1024;
    cudaFuncSetAttribute(lat<MODE, NM, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int grid : {2, 148}) {
        lat<MODE, NM, KS><<<grid, 128, smem>>>(reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[6]; cudaMemcpy(h, d, 6 * 8, cudaMemcpyDeviceToHost);
        printf("%-34s NM=%d KS=%d grid=%3d: leader issue %.0f first-commit %.0f last-commit %.0f | peer first %.0f last %.0f (%s)\n", name, NM, KS, grid,
               h[0] / 200.0, h[1] / 200.0, h[2] / 200.0, h[4] / 200.0, h[5] / 200.0, cudaGetErrorString(e));
    }
    cudaFree(d);
}
int main() {
    run<0, 1, 1>("SS commit/MMA");
    run<0, 4, 1>("SS commit/MMA");
    run<1, 4, 1>("SS one commit");
    run<0, 4, 2>("SS commit/MMA");
    run<2, 1, 1>("TS commit/MMA");
    run<2, 4, 1>("TS commit/MMA");
    run<3, 4, 1>("TS one commit");
    run<2, 4, 16>("TS commit/MMA");
    run<3, 4, 16>("TS one commit");
    return 0;
}
