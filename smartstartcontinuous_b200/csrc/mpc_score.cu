// mpc_score.cu -- the scoring tail of the MPC planner.
//
//   mpc_reduce_sums      deterministic reduction of the per-CTA projection sums
//                        (sum_k a'.b', sum_k b'.b' per time step; numerical.py:89-93)
//   mpc_score_reference  SS_PENALTY_REFERENCE second pass: re-scan the stored trajectories with
//                        the global projection coefficient of every step (NND_MB_agent.py:566-628)
//   mpc_argmax           np.argmax-ordered arg-max over the scores (NND_MB_agent.py:625-626)
#include "mpc_kernels.cuh"

namespace {

__global__ void mpc_reduce_sums_kernel(const double* __restrict__ partial, int blocks, int T,
                                       double* __restrict__ sums) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= 2 * T) return;
    double s = 0.0;
    for (int b = 0; b < blocks; ++b) s += partial[(size_t)b * 2 * T + o];
    sums[o] = s;
}

// first pass over the stored trajectories: waypoint logic per sample, per-time-step block sums of
// a'.b' and b'.b' (the two dot products of numerical.py:89-93) -> partial[block][t][2]
template <int DT>
__global__ void __launch_bounds__(256)
mpc_sums_reference_kernel(const PlanView P, int wp_index, const float* __restrict__ state0,
                          const float* __restrict__ states, long long K, int T,
                          double* __restrict__ partial) {
    __shared__ double s_part[8][2];
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool live = k < K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float x[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) x[j] = j < P.d ? state0[j] : 0.f;
    ScoreAcc sc;
    score_init<DT>(P, wp_index, x, sc);
    for (int t = 0; t < T; ++t) {
        float ab = 0.f, bb = 0.f;
        if (live) {
            const float* row = states + ((size_t)t * K + k) * P.d;
#pragma unroll
            for (int j = 0; j < DT; ++j)
                if (j < P.d) x[j] = row[j];
            score_point<DT>(P, t, x, sc, false, ab, bb);
        }
        double dab = (double)ab, dbb = (double)bb;
        for (int off = 16; off > 0; off >>= 1) {
            dab += __shfl_down_sync(0xffffffffu, dab, off);
            dbb += __shfl_down_sync(0xffffffffu, dbb, off);
        }
        if (lane == 0) { s_part[warp][0] = dab; s_part[warp][1] = dbb; }
        __syncthreads();
        if (threadIdx.x < 2) {
            double tot = 0.0;
            for (int w = 0; w < 8; ++w) tot += s_part[w][threadIdx.x];
            partial[((size_t)blockIdx.x * T + t) * 2 + threadIdx.x] = tot;
        }
        __syncthreads();
    }
}

template <int DT>
__global__ void __launch_bounds__(256)
mpc_score_reference_kernel(const PlanView P, int wp_index, const float* __restrict__ state0,
                           const float* __restrict__ states, long long K, int T,
                           const double* __restrict__ sums, float* __restrict__ scores) {
    extern __shared__ float s_lam[];
    for (int t = threadIdx.x; t < T; t += blockDim.x) s_lam[t] = (float)(sums[2 * t] / sums[2 * t + 1]);
    __syncthreads();
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= K) return;
    float x[DT];
#pragma unroll
    for (int j = 0; j < DT; ++j) x[j] = j < P.d ? state0[j] : 0.f;
    ScoreAcc sc;
    score_init<DT>(P, wp_index, x, sc);
    for (int t = 0; t < T; ++t) {
        const float* row = states + ((size_t)t * K + k) * P.d;
#pragma unroll
        for (int j = 0; j < DT; ++j)
            if (j < P.d) x[j] = row[j];
        float ab, bb;
        score_point<DT>(P, t, x, sc, false, ab, bb);
        sc.score -= penalty_with_lambda<DT>(P, sc.idx, x, s_lam[t]);
    }
    scores[k] = sc.score;
}

__global__ void __launch_bounds__(256)
mpc_argmax_kernel(const float* __restrict__ scores, long long K, long long k_offset,
                  double* __restrict__ block_v, long long* __restrict__ block_i,
                  MpcResult* __restrict__ result) {
    __shared__ double s_v[32];
    __shared__ long long s_i[32];
    __shared__ bool s_last;
    double v = 0.0;
    long long bi = -1;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < K;
         k += (long long)gridDim.x * blockDim.x) {
        double ov = (double)scores[k];
        if (argmax_better(ov, k, v, bi)) { v = ov; bi = k; }
    }
    block_argmax(v, bi, s_v, s_i);
    if (threadIdx.x == 0) {
        block_v[blockIdx.x] = v;
        block_i[blockIdx.x] = bi;
        __threadfence();
        s_last = atomicAdd(&result->blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        v = 0.0;
        bi = -1;
        for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            double ov = block_v[b];
            long long oi = block_i[b];
            if (argmax_better(ov, oi, v, bi)) { v = ov; bi = oi; }
        }
        block_argmax(v, bi, s_v, s_i);
        if (threadIdx.x == 0) {
            result->best_score = v;
            result->best_k = bi < 0 ? -1 : bi + k_offset;
            result->blocks_done = 0;
        }
    }
}

}  // namespace

int mpc_reduce_sums(ss_ctx* c, const double* partial, int blocks, int T, double* sums) {
    mpc_reduce_sums_kernel<<<(2 * T + 127) / 128, 128, 0, c->stream>>>(partial, blocks, T, sums);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

int mpc_sums_reference_blocks(long long K_local) { return (int)((K_local + 255) / 256); }

int mpc_sums_reference(ss_ctx* c, const PlanView& plan, int wp_index, const float* state0, const float* states,
                       long long K_local, int T, double* partial) {
    const unsigned grid = (unsigned)mpc_sums_reference_blocks(K_local);
    if (plan.d <= 4)
        mpc_sums_reference_kernel<4><<<grid, 256, 0, c->stream>>>(plan, wp_index, state0, states, K_local, T, partial);
    else if (plan.d <= 8)
        mpc_sums_reference_kernel<8><<<grid, 256, 0, c->stream>>>(plan, wp_index, state0, states, K_local, T, partial);
    else
        mpc_sums_reference_kernel<SS_MAX_D><<<grid, 256, 0, c->stream>>>(plan, wp_index, state0, states, K_local, T,
                                                                        partial);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

int mpc_score_reference(ss_ctx* c, const PlanView& plan, int wp_index, const float* state0,
                        const float* states, long long K_local, int T, const double* sums,
                        float* scores) {
    const unsigned grid = (unsigned)((K_local + 255) / 256);
    const size_t smem = (size_t)T * sizeof(float);
    if (plan.d <= 4)
        mpc_score_reference_kernel<4><<<grid, 256, smem, c->stream>>>(plan, wp_index, state0, states,
                                                                      K_local, T, sums, scores);
    else if (plan.d <= 8)
        mpc_score_reference_kernel<8><<<grid, 256, smem, c->stream>>>(plan, wp_index, state0, states,
                                                                      K_local, T, sums, scores);
    else
        mpc_score_reference_kernel<SS_MAX_D><<<grid, 256, smem, c->stream>>>(
            plan, wp_index, state0, states, K_local, T, sums, scores);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

int mpc_argmax(ss_ctx* c, const float* scores, long long K_local, long long k_offset, double* block_v,
               long long* block_i, void* result_dev) {
    long long want = (K_local + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
    mpc_argmax_kernel<<<grid, 256, 0, c->stream>>>(scores, K_local, k_offset, block_v, block_i,
                                                   reinterpret_cast<MpcResult*>(result_dev));
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}
