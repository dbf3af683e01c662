// Development micro-benchmark: how long after a kernel STARTS does the host see a result that the kernel
// writes to mapped pinned memory, for different store / fence sequences.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o host_flag_latency host_flag_latency.cu
#include <cstdio>
#include <cstring>
#include <chrono>
#include <cuda_runtime.h>

struct Out { unsigned long long flag; double v; long long j; int status, bits; double extra[60]; };

__global__ void k_fenced(Out* o, unsigned long long seq, int n_extra) {
    if (threadIdx.x == 0) {
        o->v = 1.5; o->j = 7; o->status = 0; o->bits = 3;
        for (int i = 0; i < n_extra; ++i) o->extra[i] = i;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&o->flag), "l"(seq) : "memory");
        __threadfence_system();
    }
}
__global__ void k_release_only(Out* o, unsigned long long seq, int n_extra) {
    if (threadIdx.x == 0) {
        o->v = 1.5; o->j = 7; o->status = 0; o->bits = 3;
        for (int i = 0; i < n_extra; ++i) o->extra[i] = i;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&o->flag), "l"(seq) : "memory");
    }
}
__global__ void k_vec32(Out* o, unsigned long long seq, int n_extra) {
    if (threadIdx.x == 0) {
        // one 32-byte store: flag + payload in a single transaction
        unsigned long long a = seq, b = __double_as_longlong(1.5), c = 7ull, d = 3ull << 32;
        asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(o), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
    }
}
__global__ void k_parallel(Out* o, unsigned long long seq, int n_extra) {
    // payload by many threads, one barrier, fence + flag by thread 0
    if (threadIdx.x < n_extra) o->extra[threadIdx.x] = threadIdx.x;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        o->v = 1.5; o->j = 7;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&o->flag), "l"(seq) : "memory");
    }
}
__global__ void k_nothing(Out* o, unsigned long long seq, int n_extra) {}

template <typename K>
void run(const char* name, K kernel, Out* h, Out* d, int n_extra, bool wait_flag) {
    cudaStream_t s; cudaStreamCreate(&s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double t_flag = 0, t_sync = 0; float t_kernel = 0;
    const int reps = 200;
    for (int it = 0; it < reps + 20; ++it) {
        unsigned long long seq = 1000 + it;
        cudaStreamSynchronize(s);
        auto t0 = std::chrono::steady_clock::now();
        cudaEventRecord(e0, s);
        kernel<<<1, 128, 0, s>>>(d, seq, n_extra);
        cudaEventRecord(e1, s);
        if (wait_flag) while (*(volatile unsigned long long*)&h->flag != seq) {}
        auto t1 = std::chrono::steady_clock::now();
        cudaStreamSynchronize(s);
        auto t2 = std::chrono::steady_clock::now();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 20) {
            t_flag += std::chrono::duration<double, std::micro>(t1 - t0).count();
            t_sync += std::chrono::duration<double, std::micro>(t2 - t0).count();
            t_kernel += ms * 1e3f;
        }
    }
    printf("%-28s extra=%2d: host sees flag after %6.2f us, stream idle after %6.2f us, events around kernel %6.2f us\n",
           name, n_extra, t_flag / reps, t_sync / reps, t_kernel / reps);
}

int main() {
    Out* h; Out* d;
    cudaHostAlloc(&h, 4096, cudaHostAllocMapped);
    memset(h, 0, 4096);
    cudaHostGetDevicePointer(&d, h, 0);
    run("empty kernel (no flag)", k_nothing, h, d, 0, false);
    for (int n_extra : {0, 56}) {
        run("fence + release + fence", k_fenced, h, d, n_extra, true);
        run("release store only", k_release_only, h, d, n_extra, true);
        run("parallel payload, release", k_parallel, h, d, n_extra, true);
    }
    run("one 32-byte store", k_vec32, h, d, 0, true);
    return 0;
}
