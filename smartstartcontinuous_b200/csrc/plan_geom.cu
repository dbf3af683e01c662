// plan_geom.cu -- plan set-up geometry on the device.
//
// path_shortcutter (numerical.py:226-246) starts from the P x P matrix of elliptical distances
// between the states of the episodic path and keeps the pairs (s, e), e >= s + 2, that lie within
// theta of each other; the weighted-interval-scheduling DP that follows (numerical.py:189-222) is
// sequential and stays on the host.  Here: the pair extraction, in float64 with the reference's
// operation order and no FMA contraction (the <= theta decisions are bit-identical to numpy), in
// np.argwhere order (s ascending, then e ascending).
#include "common.cuh"

namespace {

__device__ __forceinline__ bool close_pair(const double* __restrict__ path, const double* __restrict__ radii, int d,
                                           int s, int e, double theta) {
    // elliptical_euclidean_distance_function_generator (numerical.py:116-124):
    // sqrt(sum(((a - b) / radii) ** 2)), summed left to right like numpy does for short rows
    double acc = 0.0;
    for (int j = 0; j < d; ++j) {
        const double q = __ddiv_rn(__dsub_rn(path[(size_t)s * d + j], path[(size_t)e * d + j]), radii[j]);
        acc = __dadd_rn(acc, __dmul_rn(q, q));
    }
    return __dsqrt_rn(acc) <= theta;
}

// one block per row s: count (FILL = false) or write (FILL = true) the pairs of that row in e order
template <bool FILL>
__global__ void __launch_bounds__(256)
close_pairs_kernel(const double* __restrict__ path, const double* __restrict__ radii, int P, int d, double theta,
                   long long* __restrict__ row_count, const long long* __restrict__ row_offset,
                   int* __restrict__ out_pairs, long long max_pairs) {
    __shared__ int s_warp[8];
    __shared__ long long s_base;
    const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = FILL ? row_offset[s] : 0;
    __syncthreads();
    long long total = 0;
    for (int e0 = s + 2; e0 < P; e0 += blockDim.x) {
        const int e = e0 + threadIdx.x;
        const bool hit = e < P && close_pair(path, radii, d, s, e, theta);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, all = 0;
        for (int w = 0; w < 8; ++w) {
            if (w < warp) before += s_warp[w];
            all += s_warp[w];
        }
        if (FILL && hit) {
            const long long pos = s_base + total + before + __popc(m & ((1u << lane) - 1u));
            if (pos < max_pairs) {
                out_pairs[2 * pos] = s;
                out_pairs[2 * pos + 1] = e;
            }
        }
        total += all;
        __syncthreads();
    }
    if (!FILL && threadIdx.x == 0) row_count[s] = total;
}

// exclusive scan of the row counts (P is a path length: a few thousand at most), total in out[P]
__global__ void scan_rows_kernel(const long long* __restrict__ count, int P, long long* __restrict__ offset) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long run = 0;
        for (int i = 0; i < P; ++i) { offset[i] = run; run += count[i]; }
        offset[P] = run;
    }
}

}  // namespace

extern "C" int ss_path_close_pairs(ss_ctx* c, const double* path, int P, int d, const double* radii, double theta,
                                   int32_t* out_pairs, int64_t max_pairs, int64_t* out_count) {
    if (!c) return SS_EINVAL;
    if (!path || !radii || !out_count || P < 1 || d < 1 || max_pairs < 0 || (max_pairs > 0 && !out_pairs))
        SS_FAIL(c, SS_EINVAL, "path_close_pairs: bad arguments");
    for (int j = 0; j < d; ++j)
        if (!(radii[j] > 0.0)) SS_FAIL(c, SS_EINVAL, "path_close_pairs: radii must be > 0 (AssertionError in the reference)");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    const size_t path_bytes = (size_t)P * d * 8;
    SS_CUDA_CHECK(c, c->geom_in.ensure(path_bytes + (size_t)d * 8));
    SS_CUDA_CHECK(c, c->geom_rows.ensure((size_t)(2 * P + 1) * 8));
    SS_CUDA_CHECK(c, c->geom_pairs.ensure((size_t)(max_pairs > 0 ? max_pairs : 1) * 8));
    double* path_dev = c->geom_in.as<double>();
    double* radii_dev = path_dev + (size_t)P * d;
    long long* count = c->geom_rows.as<long long>();
    long long* offset = count + P;
    SS_CUDA_CHECK(c, cudaMemcpyAsync(path_dev, path, path_bytes, cudaMemcpyHostToDevice, c->stream));
    SS_CUDA_CHECK(c, cudaMemcpyAsync(radii_dev, radii, (size_t)d * 8, cudaMemcpyHostToDevice, c->stream));
    close_pairs_kernel<false><<<P, 256, 0, c->stream>>>(path_dev, radii_dev, P, d, theta, count, nullptr, nullptr, 0);
    scan_rows_kernel<<<1, 32, 0, c->stream>>>(count, P, offset);
    close_pairs_kernel<true><<<P, 256, 0, c->stream>>>(path_dev, radii_dev, P, d, theta, nullptr, offset,
                                                        c->geom_pairs.as<int>(), max_pairs);
    c->launches += 3;
    SS_CUDA_CHECK(c, cudaGetLastError());
    long long total = 0;
    SS_CUDA_CHECK(c, cudaMemcpyAsync(&total, offset + P, 8, cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    *out_count = total;
    const long long n_copy = total < max_pairs ? total : max_pairs;
    if (n_copy > 0) {
        SS_CUDA_CHECK(c, cudaMemcpyAsync(out_pairs, c->geom_pairs.p, (size_t)n_copy * 8, cudaMemcpyDeviceToHost, c->stream));
        SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    }
    return SS_OK;
}
