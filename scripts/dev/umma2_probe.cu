// Probe: tcgen05.mma.cta_group::2 (M=256 over a CTA pair, N=128, K=32), A from TMEM, B split by N over the
// two CTAs' shared memory, commit multicast to both CTAs.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <vector>
#include <cstdint>
#include <cmath>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int M = 256, N = 128, K = 32, NH = N / 2;
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe(const float* __restrict__ A, const __nv_bfloat16* __restrict__ Bimg, float* __restrict__ D) {
    __shared__ __align__(128) __nv_bfloat16 sB[NH * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = tid; i < NH * K; i += 128) sB[i] = Bimg[rank * NH * K + i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // A rows of this CTA: global row = rank*128 + tid
    uint32_t a[16];
    for (int c = 0; c < 16; ++c) {
        float lo = A[(rank * 128 + tid) * K + 2 * c], hi = A[(rank * 128 + tid) * K + 2 * c + 1];
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a[c]) : "f"(hi), "f"(lo));
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(tmem + lane_base + 128), "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]),
        "r"(a[7]), "r"(a[8]), "r"(a[9]), "r"(a[10]), "r"(a[11]), "r"(a[12]), "r"(a[13]), "r"(a[14]), "r"(a[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (rank == 0 && tid == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint32_t addr = smem_u32(sB) + ks * 2 * (NH * 16);
            uint64_t desc = (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((NH * 16) >> 4) << 16) |
                            ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
            const uint32_t acc = ks > 0;
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
                "r"(tmem + 128 + ks * 8), "l"(desc), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    }
    {
        uint32_t done = 0; int spins = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
            if (++spins > 4000000) { if (tid == 0) printf("timeout rank %u\n", rank); asm volatile("trap;"); }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t r[32];
    for (int part = 0; part < 4; ++part) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem + lane_base + part * 32));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(rank * 128 + tid) * N + part * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}
int main() {
    std::vector<float> A(M * K), B(N * K), Dref(M * N, 0.f), D(M * N);
    for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k) A[r * K + k] = (float)((r * 3 + k * 5) % 7 - 3);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = (float)((n * 2 + k * 3) % 5 - 2);
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[r * K + k] * B[n * K + k]; Dref[r * N + n] = s; }
    // per-CTA half images: half h holds units [64h, 64h+64): [k/8][64][8]
    std::vector<__nv_bfloat16> img(N * K);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k)
        img[(n / NH) * NH * K + (k / 8) * (NH * 8) + (n % NH) * 8 + (k % 8)] = __float2bfloat16(B[n * K + k]);
    float *dA, *dD; __nv_bfloat16* dB;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dB, img.size() * 2);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    probe<<<2, 128>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_top = 0, bad_left = 0;
    for (int i = 0; i < M * N; ++i) if (fabs(D[i] - Dref[i]) > 1e-3) { ++bad; if (i / N < 128) ++bad_top; if (i % N < 64) ++bad_left; }
    printf("cta_group::2 mismatches=%d/%d (rows<128: %d, cols<64: %d)  D[0][0..3]=%g %g %g %g ref=%g %g %g %g | D[200][100]=%g ref=%g\n",
           bad, M * N, bad_top, bad_left, D[0], D[1], D[2], D[3], Dref[0], Dref[1], Dref[2], Dref[3], D[200 * N + 100], Dref[200 * N + 100]);
    return 0;
}
