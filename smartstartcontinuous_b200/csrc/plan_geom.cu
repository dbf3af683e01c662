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

// mask[e][s] = 1 when (s, e) is a shortcut candidate: e >= s + 2 and within theta
__global__ void __launch_bounds__(256)
close_mask_kernel(const double* __restrict__ path, const double* __restrict__ radii, int P, int d, double theta,
                  unsigned char* __restrict__ mask) {
    const int e = blockIdx.x;
    for (int s = threadIdx.x; s < P; s += blockDim.x)
        mask[(size_t)e * P + s] = (s + 2 <= e && close_pair(path, radii, d, s, e, theta)) ? 1 : 0;
}

// length_weighted_activities_solver (numerical.py:189-222) with weight = end - start - 1 over the
// candidate intervals, in the dense form of its column DP: one column per distinct end value e (in
// increasing order); bval[x] = best weight of the last column with end <= x.  For a column,
//   carry   = best of the previous column,
//   with(s) = bval[s] + e - s - 1 for every candidate start s (ascending),
//   the reference replaces the column's choice whenever with >= the current best, so the result is
//   the LARGEST s attaining max(with) if that maximum is >= carry, else "carry, nothing taken";
//   the globally first interval (smallest e, then smallest s) is taken unconditionally with weight
//   e - s (numerical.py:202 ignores sub_extra there).
// One block walks the ends sequentially, the candidates of an end in parallel.  keep[i] = 1 for
// the states that survive (interior states of the chosen intervals are dropped).
__global__ void __launch_bounds__(1024)
shortcut_dp_kernel(const unsigned char* __restrict__ mask, int P, int* __restrict__ bval, int* __restrict__ lastcol,
                   int* __restrict__ taken, int* __restrict__ back, unsigned char* __restrict__ keep) {
    __shared__ int s_with[32], s_smax[32], s_smin[32];
    __shared__ int s_first_done;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_first_done = 0;
    for (int i = tid; i < P; i += blockDim.x) keep[i] = 1;
    __syncthreads();
    for (int e = 0; e < P; ++e) {
        // best candidate of this end: max with, ties -> larger s; and the smallest s (first column)
        int w = -1, smax = -1, smin = 0x7fffffff;
        for (int s = tid; s + 2 <= e; s += blockDim.x)
            if (mask[(size_t)e * P + s]) {
                const int v = bval[s] + e - s - 1;
                if (v > w || (v == w && s > smax)) { w = v; smax = s; }
                if (s < smin) smin = s;
            }
        for (int off = 16; off > 0; off >>= 1) {
            const int ow = __shfl_xor_sync(0xffffffffu, w, off), os = __shfl_xor_sync(0xffffffffu, smax, off);
            const int om = __shfl_xor_sync(0xffffffffu, smin, off);
            if (ow > w || (ow == w && os > smax)) { w = ow; smax = os; }
            if (om < smin) smin = om;
        }
        if (lane == 0) { s_with[warp] = w; s_smax[warp] = smax; s_smin[warp] = smin; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < 32; ++k) {
                if (s_with[k] > w || (s_with[k] == w && s_smax[k] > smax)) { w = s_with[k]; smax = s_smax[k]; }
                if (s_smin[k] < smin) smin = s_smin[k];
            }
            const int prev_best = e > 0 ? bval[e - 1] : 0;
            const int prev_col = e > 0 ? lastcol[e - 1] : -1;
            if (smax < 0) {                       // no interval ends here: not a column
                bval[e] = prev_best;
                lastcol[e] = prev_col;
                taken[e] = -1;
                back[e] = prev_col;
            } else if (!s_first_done) {           // the very first interval of the sorted list
                s_first_done = 1;
                bval[e] = e - smin;
                taken[e] = smin;
                back[e] = -1;
                lastcol[e] = e;
            } else if (w >= prev_best) {          // taking wins ties
                bval[e] = w;
                taken[e] = smax;
                back[e] = lastcol[smax];
                lastcol[e] = e;
            } else {                              // carry the previous column
                bval[e] = prev_best;
                taken[e] = -1;
                back[e] = prev_col;
                lastcol[e] = e;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        int col = P > 0 ? lastcol[P - 1] : -1;
        while (col >= 0) {
            if (taken[col] >= 0)
                for (int i = taken[col] + 1; i < col; ++i) keep[i] = 0;
            col = back[col];
        }
    }
}

}  // namespace

extern "C" int ss_path_shortcut(ss_ctx* c, const double* path, int P, int d, const double* radii, double theta,
                                int32_t* out_keep, int* out_count) {
    if (!c) return SS_EINVAL;
    if (!path || !radii || !out_keep || !out_count || P < 1 || d < 1) SS_FAIL(c, SS_EINVAL, "path_shortcut: bad arguments");
    if (P > 16384) SS_FAIL(c, SS_EUNSUPPORTED, "path_shortcut: path longer than 16384 states");
    for (int j = 0; j < d; ++j)
        if (!(radii[j] > 0.0)) SS_FAIL(c, SS_EINVAL, "path_shortcut: radii must be > 0 (AssertionError in the reference)");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    const size_t path_bytes = (size_t)P * d * 8;
    SS_CUDA_CHECK(c, c->geom_in.ensure(path_bytes + (size_t)d * 8));
    SS_CUDA_CHECK(c, c->geom_rows.ensure((size_t)4 * P * 4 + P));
    SS_CUDA_CHECK(c, c->geom_pairs.ensure((size_t)P * P));
    double* path_dev = c->geom_in.as<double>();
    double* radii_dev = path_dev + (size_t)P * d;
    int* bval = c->geom_rows.as<int>();
    int* lastcol = bval + P;
    int* taken = lastcol + P;
    int* back = taken + P;
    unsigned char* keep = reinterpret_cast<unsigned char*>(back + P);
    unsigned char* mask = c->geom_pairs.as<unsigned char>();
    SS_CUDA_CHECK(c, cudaMemcpyAsync(path_dev, path, path_bytes, cudaMemcpyHostToDevice, c->stream));
    SS_CUDA_CHECK(c, cudaMemcpyAsync(radii_dev, radii, (size_t)d * 8, cudaMemcpyHostToDevice, c->stream));
    close_mask_kernel<<<P, 256, 0, c->stream>>>(path_dev, radii_dev, P, d, theta, mask);
    shortcut_dp_kernel<<<1, 1024, 0, c->stream>>>(mask, P, bval, lastcol, taken, back, keep);
    c->launches += 2;
    SS_CUDA_CHECK(c, cudaGetLastError());
    std::vector<unsigned char> hk(P);
    SS_CUDA_CHECK(c, cudaMemcpyAsync(hk.data(), keep, P, cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    int n = 0;
    for (int i = 0; i < P; ++i)
        if (hk[i]) out_keep[n++] = i;
    *out_count = n;
    return SS_OK;
}

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
