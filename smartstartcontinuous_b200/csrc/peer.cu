// peer.cu -- the two small exchanges of a sharded MPC decision over NVLink peer memory, fused with
// the kernels that produce their payload (no NCCL call, no host round trip):
//
//   1. projection sums (reference penalty, numerical.py:89-93 sums over ALL K sequences): the kernel
//      that reduces this GPU's per-block partials also stores the 2(H+1) doubles into every peer's
//      receive slot, raises a flag, waits for the peers' flags and adds the slots in rank order
//      -> every rank holds bit-identical global sums;
//   2. winner packages (score, k, sequence, path): the kernel that packs this GPU's winner also
//      stores the package into every peer's slot, waits for theirs and picks the global winner with
//      np.argmax ordering -> every rank ends with the same winner package in its own memory.
//
// Exchange buffers are plain cudaMalloc allocations shared through CUDA IPC handles (one process per
// GPU); slots and flags are double-buffered by the epoch parity of their channel, so a fast rank
// can never overwrite data a slow rank is still reading.  Stores to peer
// memory are followed by __threadfence_system() and a release-scope flag store; readers poll the
// flag with acquire scope and read the payload past L1 (ld.global.cg).  A missing peer traps after
// PEER_TIMEOUT_NS instead of hanging the GPU.
#include <atomic>
#include <cstring>

#include "mpc_kernels.cuh"

namespace {

constexpr int PEER_CH_SUMS = 0, PEER_CH_PKG = 1;
constexpr unsigned long long PEER_TIMEOUT_NS = 10ull * 1000 * 1000 * 1000;   // 10 s

__host__ __device__ inline size_t peer_slot_index(int ch, int par, int src) {
    return ((size_t)(ch * 2 + par) * SS_PEER_MAX_WORLD + src);
}
__host__ __device__ inline size_t peer_flags_offset_doubles() {
    return (size_t)4 * SS_PEER_MAX_WORLD * SS_PEER_SLOT_DOUBLES;
}
size_t peer_buffer_bytes() { return (peer_flags_offset_doubles() + 4 * SS_PEER_MAX_WORLD) * 8; }

__device__ __forceinline__ double* peer_slot(double* base, int ch, int par, int src) {
    return base + peer_slot_index(ch, par, src) * SS_PEER_SLOT_DOUBLES;
}
__device__ __forceinline__ unsigned long long* peer_flag(double* base, int ch, int par, int src) {
    return reinterpret_cast<unsigned long long*>(base + peer_flags_offset_doubles()) + peer_slot_index(ch, par, src);
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// thread `src` (< world) waits until rank `src` has delivered epoch `epoch` on the channel
__device__ __forceinline__ void peer_wait(double* my_base, int ch, int par, int src, unsigned long long epoch) {
    const unsigned long long* f = peer_flag(my_base, ch, par, src);
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(f) < epoch)
        if (global_ns() - t0 > PEER_TIMEOUT_NS) asm volatile("trap;");
}

// ---- 1. per-block partials -> local sums -> all-reduce over the peers ----------------------
__global__ void __launch_bounds__(1024)
mpc_reduce_allreduce_kernel(const double* __restrict__ partial, int blocks, int T, double* __restrict__ sums,
                            PeerView pv, unsigned long long epoch) {
    __shared__ double s_my[SS_PEER_SLOT_DOUBLES > 2048 ? 2048 : SS_PEER_SLOT_DOUBLES];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = 2 * T;
    for (int o = warp; o < n; o += 32) {                 // same fixed order as mpc_reduce_sums_kernel
        const int t = o >> 1, w = o & 1;
        double s = 0.0;
        for (int b = lane; b < blocks; b += 32) s += partial[(size_t)o * blocks + b];      // [t][2][blocks]
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) s_my[o] = s;
    }
    __syncthreads();
    const int par = (int)(epoch & 1);
    for (int r = 0; r < pv.world; ++r) {
        double* dst = peer_slot(pv.base[r], PEER_CH_SUMS, par, pv.rank);
        for (int o = tid; o < n; o += blockDim.x) dst[o] = s_my[o];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pv.world) {
        st_release_sys(peer_flag(pv.base[tid], PEER_CH_SUMS, par, pv.rank), epoch);
        peer_wait(pv.base[pv.rank], PEER_CH_SUMS, par, tid, epoch);
    }
    __syncthreads();
    double* mine = pv.base[pv.rank];
    for (int o = tid; o < n; o += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pv.world; ++r) s += __ldcg(peer_slot(mine, PEER_CH_SUMS, par, r) + o);   // rank order
        sums[o] = s;
    }
}

// ---- 2. local winner package -> all peers -> global pick ----------------------------------
// pkg_local: [score, k, sequence, path] of this rank (n doubles); pkg_out: the global winner's package
__global__ void __launch_bounds__(256)
mpc_package_exchange_kernel(const double* __restrict__ pkg_local, int n, double* __restrict__ pkg_out, PeerView pv,
                            unsigned long long epoch) {
    __shared__ int s_win;
    const int tid = threadIdx.x;
    const int par = (int)(epoch & 1);
    for (int r = 0; r < pv.world; ++r) {
        double* dst = peer_slot(pv.base[r], PEER_CH_PKG, par, pv.rank);
        for (int o = tid; o < n; o += blockDim.x) dst[o] = pkg_local[o];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pv.world) {
        st_release_sys(peer_flag(pv.base[tid], PEER_CH_PKG, par, pv.rank), epoch);
        peer_wait(pv.base[pv.rank], PEER_CH_PKG, par, tid, epoch);
    }
    __syncthreads();
    double* mine = pv.base[pv.rank];
    if (tid == 0) {
        // np.argmax ordering over (score, global k): NaN first, then larger score, then lower k
        int best = -1;
        double bv = 0.0;
        long long bk = -1;
        for (int r = 0; r < pv.world; ++r) {
            const double v = __ldcg(peer_slot(mine, PEER_CH_PKG, par, r));
            const long long k = (long long)__ldcg(peer_slot(mine, PEER_CH_PKG, par, r) + 1);
            if (argmax_better(v, k, bv, bk)) { best = r; bv = v; bk = k; }
        }
        s_win = best < 0 ? pv.rank : best;
    }
    __syncthreads();
    const double* src = peer_slot(mine, PEER_CH_PKG, par, s_win);
    for (int o = tid; o < n; o += blockDim.x) pkg_out[o] = __ldcg(src + o);
}

// ---- 3. (value, global index) pairs -> all peers -> np.argmax-ordered pick (KDE query shards) ---------
// the pair travels as kernel arguments; the result lands in mapped pinned host memory with a flag
struct PeerPickOut {
    unsigned long long flag;
    double value;
    long long index;
};
__global__ void __launch_bounds__(32)
peer_argmax_exchange_kernel(double value, long long index, PeerView pv, unsigned long long epoch,
                            PeerPickOut* __restrict__ host_out, unsigned long long seq) {
    const int tid = threadIdx.x;
    const int par = (int)(epoch & 1);
    if (tid < pv.world) {
        double* dst = peer_slot(pv.base[tid], PEER_CH_PKG, par, pv.rank);
        dst[0] = value;
        dst[1] = (double)index;
        __threadfence_system();
        st_release_sys(peer_flag(pv.base[tid], PEER_CH_PKG, par, pv.rank), epoch);
        peer_wait(pv.base[pv.rank], PEER_CH_PKG, par, tid, epoch);
    }
    __syncwarp();
    if (tid == 0) {
        double* mine = pv.base[pv.rank];
        double bv = 0.0;
        long long bk = -1;
        for (int r = 0; r < pv.world; ++r) {
            const double v = __ldcg(peer_slot(mine, PEER_CH_PKG, par, r));
            const long long k = (long long)__ldcg(peer_slot(mine, PEER_CH_PKG, par, r) + 1);
            if (argmax_better(v, k, bv, bk)) { bv = v; bk = k; }
        }
        host_out->value = bv;
        host_out->index = bk;
        __threadfence_system();
        st_release_sys(&host_out->flag, seq);
    }
}

}  // namespace

// every rank contributes (value, global index; index < 0 = nothing); every rank gets the np.argmax-ordered
// winner (NaN first, larger value, lower index).  Shares the package channel (same epoch sequence on every rank).
extern "C" int ss_peer_argmax_merge(ss_ctx* c, double value, int64_t index, double* out_value, int64_t* out_index) {
    if (!c) return SS_EINVAL;
    if (!c->peer_ready) SS_FAIL(c, SS_ESTATE, "peer exchange: not open (ss_peer_init / ss_peer_open)");
    if (!out_value || !out_index) SS_FAIL(c, SS_EINVAL, "peer exchange: null output");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    if (!c->host_kde) {
        SS_CUDA_CHECK(c, cudaHostAlloc(&c->host_kde, 4096, cudaHostAllocMapped));
        std::memset(c->host_kde, 0, 4096);
        SS_CUDA_CHECK(c, cudaHostGetDevicePointer(&c->host_kde_dev, c->host_kde, 0));
    }
    // the pick result uses the second half of the mapped page (the first holds the KDE result)
    PeerPickOut* dev_out = reinterpret_cast<PeerPickOut*>(reinterpret_cast<char*>(c->host_kde_dev) + 2048);
    volatile PeerPickOut* ho = reinterpret_cast<volatile PeerPickOut*>(reinterpret_cast<char*>(c->host_kde) + 2048);
    const unsigned long long seq = ++c->host_kde_seq;
    c->peer_epoch[PEER_CH_PKG]++;
    peer_argmax_exchange_kernel<<<1, 32, 0, c->stream>>>(value, (long long)index, c->peer_view, c->peer_epoch[PEER_CH_PKG],
                                                         dev_out, seq);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    unsigned spins = 0;
    while (ho->flag != seq) {
        if ((++spins & 0x3fff) == 0) {
            cudaError_t qe = cudaStreamQuery(c->stream);
            if (qe != cudaSuccess && qe != cudaErrorNotReady) SS_CUDA_CHECK(c, qe);
            if (qe == cudaSuccess && ho->flag != seq) {
                SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
                if (ho->flag != seq) SS_FAIL(c, SS_ECUDA, "peer exchange: the pick kernel ended without raising its flag");
            }
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    *out_value = ho->value;
    *out_index = (int64_t)ho->index;
    return SS_OK;
}

int peer_allreduce_sums(ss_ctx* c, const double* partial, int blocks, int T, double* sums) {
    if (2 * T > 2048 || 2 * T > SS_PEER_SLOT_DOUBLES)
        SS_FAIL(c, SS_EUNSUPPORTED, "peer exchange: horizon too long for the sums slot");
    c->peer_epoch[PEER_CH_SUMS]++;
    mpc_reduce_allreduce_kernel<<<1, 1024, 0, c->stream>>>(partial, blocks, T, sums, c->peer_view,
                                                           c->peer_epoch[PEER_CH_SUMS]);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

int peer_package_exchange(ss_ctx* c, const double* pkg_local, int n, double* pkg_out) {
    if (n > SS_PEER_SLOT_DOUBLES) SS_FAIL(c, SS_EUNSUPPORTED, "peer exchange: package larger than a slot");
    c->peer_epoch[PEER_CH_PKG]++;
    mpc_package_exchange_kernel<<<1, 256, 0, c->stream>>>(pkg_local, n, pkg_out, c->peer_view,
                                                          c->peer_epoch[PEER_CH_PKG]);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

extern "C" int ss_peer_init(ss_ctx* c, int rank, int world, void* out_handle64) {
    if (!c) return SS_EINVAL;
    if (world < 1 || world > SS_PEER_MAX_WORLD || rank < 0 || rank >= world || !out_handle64)
        SS_FAIL(c, SS_EINVAL, "peer exchange: need 0 <= rank < world <= 8");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    c->peer_ready = false;
    // a re-initialisation (new group / world size) starts from a clean slate: nothing of this context
    // may still be reading the buffer, mappings of the previous peers are closed, and flags AND slots
    // are zeroed so that the epochs (restarted at 0 by ss_peer_open on every rank) can never match
    // a flag left over from the previous exchange.  Callers all-gather the handles after this call,
    // which orders every rank's clear before any peer's first store.
    if (c->stream) SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    for (int r = 0; r < SS_PEER_MAX_WORLD; ++r)
        if (c->peer_opened[r]) { cudaIpcCloseMemHandle(c->peer_opened[r]); c->peer_opened[r] = nullptr; }
    if (!c->peer_buf) SS_CUDA_CHECK(c, cudaMalloc(&c->peer_buf, peer_buffer_bytes()));
    SS_CUDA_CHECK(c, cudaMemset(c->peer_buf, 0, peer_buffer_bytes()));
    SS_CUDA_CHECK(c, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SS_CUDA_CHECK(c, cudaIpcGetMemHandle(&h, c->peer_buf));
    std::memcpy(out_handle64, &h, 64);
    c->peer_view.rank = rank;
    c->peer_view.world = world;
    return SS_OK;
}

extern "C" int ss_peer_open(ss_ctx* c, const void* handles, int world) {
    if (!c) return SS_EINVAL;
    if (!c->peer_buf || !handles || world != c->peer_view.world) SS_FAIL(c, SS_ESTATE, "peer exchange: ss_peer_init first");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    for (int r = 0; r < world; ++r) {
        if (r == c->peer_view.rank) {
            c->peer_view.base[r] = reinterpret_cast<double*>(c->peer_buf);
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)r * 64, 64);
        void* p = nullptr;
        SS_CUDA_CHECK(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_view.base[r] = reinterpret_cast<double*>(p);
        c->peer_opened[r] = p;
    }
    c->peer_epoch[0] = c->peer_epoch[1] = 0;
    c->peer_ready = true;
    return SS_OK;
}

extern "C" int ss_peer_close(ss_ctx* c) {
    if (!c) return SS_EINVAL;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int r = 0; r < SS_PEER_MAX_WORLD; ++r)
        if (c->peer_opened[r]) { cudaIpcCloseMemHandle(c->peer_opened[r]); c->peer_opened[r] = nullptr; }
    if (c->peer_buf) { cudaFree(c->peer_buf); c->peer_buf = nullptr; }
    c->peer_ready = false;
    return SS_OK;
}

extern "C" int ss_peer_ready(ss_ctx* c) { return c && c->peer_ready ? 1 : 0; }
