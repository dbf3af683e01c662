// value_net.cu -- SURVEY 8f row f4: the critic value batch V_j = critic(q_j, actor(q_j)) that feeds the
// UCB (smartexplorationcontinuous.py:274 -> DDPG_Baselines_agent.py:197-204 -> ddpg_editted.py:274-279,
// graph :106-131, networks models_editted.py:22-100) evaluated on the device in front of the KDE, so the
// selection call needs no host value vector.
//
//   obs_n = clip((q - obs_mean) / obs_std, obs_range)                       (ddpg_editted.py:106-107; optional)
//   actor:  dense(h1) [LN] relu -> dense(h2) [LN] relu|tanh -> dense(da) tanh          (models_editted.py:38-58)
//   critic: dense(h1) [LN] relu -> concat action -> dense(h2) [LN] relu|tanh -> dense(1)        (:81-99)
//   V = clip(critic, return_range) * ret_std + ret_mean                      (ddpg_editted.py:130-131; optional)
//
// FP32 like the TF graph.  One warp per query at a time, all parameters in shared memory, units spread
// over the lanes, layer norm (tf.contrib.layers.layer_norm: biased variance, eps 1e-12) by warp shuffles.
#include "common.cuh"

namespace {

struct ValueNetDev {
    int d, da, h1a, h2a, h1c, h2c, layer_norm, last_tanh, obs_norm, ret_norm;
    // offsets (floats) into the parameter block
    int aW1, ab1, ag1, abe1, aW2, ab2, ag2, abe2, aW3, ab3;
    int cW1, cb1, cg1, cbe1, cW2, cb2, cg2, cbe2, cW3, cb3;
    int obs_mean, obs_inv_std, total;
    float obs_lo, obs_hi, ret_lo, ret_hi, ret_mean, ret_std;
};

// out[u] = b[u] + sum_k in[k] * W[k][u] for the lane's units u = lane, lane + 32, ...
__device__ __forceinline__ void dense(const float* __restrict__ in, int n_in, const float* __restrict__ W,
                                      const float* __restrict__ b, int n_out, float* __restrict__ out, int lane) {
    for (int u = lane; u < n_out; u += 32) {
        float acc = b[u];
        for (int k = 0; k < n_in; ++k) acc = fmaf(in[k], W[k * n_out + u], acc);
        out[u] = acc;
    }
    __syncwarp();
}

// optional layer norm over the n units, then relu (act = 0) or tanh (act = 1); in place
__device__ __forceinline__ void norm_act(float* __restrict__ v, int n, bool layer_norm, const float* __restrict__ g,
                                         const float* __restrict__ be, int act, int lane) {
    float mean = 0.f, rstd = 1.f;
    if (layer_norm) {
        float s = 0.f;
        for (int u = lane; u < n; u += 32) s += v[u];
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        mean = s / n;
        float q = 0.f;
        for (int u = lane; u < n; u += 32) { const float df = v[u] - mean; q = fmaf(df, df, q); }
        for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
        rstd = rsqrtf(q / n + 1e-12f);
    }
    for (int u = lane; u < n; u += 32) {
        float x = v[u];
        if (layer_norm) x = (x - mean) * rstd * g[u] + be[u];
        v[u] = act ? tanhf(x) : fmaxf(x, 0.f);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(128)
value_net_kernel(const ValueNetDev N, const float* __restrict__ params, const double* __restrict__ queries,
                 long long m, float* __restrict__ values) {
    extern __shared__ float sm[];
    float* P = sm;                                         // parameters
    for (int i = threadIdx.x; i < N.total; i += blockDim.x) P[i] = params[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int hmax = max(max(N.h1a, N.h2a), max(N.h1c, N.h2c)) + N.da;
    float* xs = sm + N.total + warp * (N.d + 2 * hmax + N.da);   // per-warp scratch: x | bufA | bufB | action
    float* bufA = xs + N.d;
    float* bufB = bufA + hmax;
    float* act = bufB + hmax;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long q = (long long)blockIdx.x * (blockDim.x >> 5) + warp; q < m; q += warps) {
        for (int j = lane; j < N.d; j += 32) {
            float x = (float)queries[q * N.d + j];
            // clip(normalize(obs, obs_rms), observation_range): the clip applies with or without statistics
            // (normalize() is the identity for obs_rms = None, ddpg_editted.py:106-109; blob holds mean 0, 1/std 1 then)
            x = fminf(fmaxf((x - P[N.obs_mean + j]) * P[N.obs_inv_std + j], N.obs_lo), N.obs_hi);
            xs[j] = x;
        }
        __syncwarp();
        // actor
        dense(xs, N.d, P + N.aW1, P + N.ab1, N.h1a, bufA, lane);
        norm_act(bufA, N.h1a, N.layer_norm, P + N.ag1, P + N.abe1, 0, lane);
        dense(bufA, N.h1a, P + N.aW2, P + N.ab2, N.h2a, bufB, lane);
        norm_act(bufB, N.h2a, N.layer_norm, P + N.ag2, P + N.abe2, N.last_tanh, lane);
        dense(bufB, N.h2a, P + N.aW3, P + N.ab3, N.da, act, lane);
        for (int u = lane; u < N.da; u += 32) act[u] = tanhf(act[u]);
        __syncwarp();
        // critic
        dense(xs, N.d, P + N.cW1, P + N.cb1, N.h1c, bufA, lane);
        norm_act(bufA, N.h1c, N.layer_norm, P + N.cg1, P + N.cbe1, 0, lane);
        for (int u = lane; u < N.da; u += 32) bufA[N.h1c + u] = act[u];      // concat([h, action])
        __syncwarp();
        dense(bufA, N.h1c + N.da, P + N.cW2, P + N.cb2, N.h2c, bufB, lane);
        norm_act(bufB, N.h2c, N.layer_norm, P + N.cg2, P + N.cbe2, N.last_tanh, lane);
        float v = 0.f;
        for (int k = lane; k < N.h2c; k += 32) v = fmaf(bufB[k], P[N.cW3 + k], v);
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) {
            v += P[N.cb3];
            // denormalize(clip_by_value(critic, return_range), ret_rms): the clip is unconditional, the affine
            // map only exists with return normalisation (ddpg_editted.py:130-131)
            v = fminf(fmaxf(v, N.ret_lo), N.ret_hi);
            if (N.ret_norm) v = v * N.ret_std + N.ret_mean;
            values[q] = v;
        }
        __syncwarp();
    }
}

}  // namespace

struct ValueNetHost {
    ValueNetDev dev;
    size_t smem = 0;
};

int value_net_eval_dev(ss_ctx* c, const double* queries_dev, long long m, float* values_dev) {
    if (!c->value_net_set) SS_FAIL(c, SS_ESTATE, "value net: ss_value_net_set first");
    const ValueNetDev& N = *reinterpret_cast<const ValueNetDev*>(c->value_net_desc.data());
    const int hmax = std::max(std::max(N.h1a, N.h2a), std::max(N.h1c, N.h2c)) + N.da;
    const size_t smem = ((size_t)N.total + 4 * (size_t)(N.d + 2 * hmax + N.da)) * 4;
    SS_CUDA_CHECK(c, cudaFuncSetAttribute(value_net_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long blocks = (m + 3) / 4;
    if (blocks > (long long)c->sm_count * 8) blocks = (long long)c->sm_count * 8;
    value_net_kernel<<<(unsigned)blocks, 128, smem, c->stream>>>(N, c->value_net_params.as<float>(), queries_dev, m,
                                                                   values_dev);
    c->launches++;
    SS_CUDA_CHECK(c, cudaGetLastError());
    return SS_OK;
}

extern "C" int ss_value_net_set(ss_ctx* c, const ss_value_net* n) {
    if (!c) return SS_EINVAL;
    if (!n) { c->value_net_set = false; return SS_OK; }
    if (n->d < 1 || n->d > SS_MAX_D || n->da < 1 || n->da > SS_MAX_DA || n->h1a < 1 || n->h2a < 1 || n->h1c < 1 ||
        n->h2c < 1)
        SS_FAIL(c, SS_EINVAL, "value net: bad sizes");
    const bool ln = n->layer_norm != 0;
    const float* need[] = {n->aW1, n->ab1, n->aW2, n->ab2, n->aW3, n->ab3, n->cW1, n->cb1, n->cW2, n->cb2, n->cW3, n->cb3};
    for (const float* p : need)
        if (!p) SS_FAIL(c, SS_EINVAL, "value net: null weight pointer");
    if (ln && (!n->ag1 || !n->abe1 || !n->ag2 || !n->abe2 || !n->cg1 || !n->cbe1 || !n->cg2 || !n->cbe2))
        SS_FAIL(c, SS_EINVAL, "value net: layer_norm needs gamma / beta");
    if ((n->obs_mean == nullptr) != (n->obs_std == nullptr)) SS_FAIL(c, SS_EINVAL, "value net: obs_mean and obs_std go together");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    ValueNetDev N;
    std::memset(&N, 0, sizeof(N));
    N.d = n->d; N.da = n->da; N.h1a = n->h1a; N.h2a = n->h2a; N.h1c = n->h1c; N.h2c = n->h2c;
    N.layer_norm = ln; N.last_tanh = n->last_layer_tanh != 0;
    N.obs_norm = n->obs_mean != nullptr; N.ret_norm = n->has_ret_norm != 0;
    // an empty range (lo >= hi, e.g. a zero-initialised struct) means "no clip"
    const bool obs_clip = n->obs_clip_lo < n->obs_clip_hi, ret_clip = n->ret_clip_lo < n->ret_clip_hi;
    N.obs_lo = obs_clip ? (float)n->obs_clip_lo : -INFINITY; N.obs_hi = obs_clip ? (float)n->obs_clip_hi : INFINITY;
    N.ret_lo = ret_clip ? (float)n->ret_clip_lo : -INFINITY; N.ret_hi = ret_clip ? (float)n->ret_clip_hi : INFINITY;
    N.ret_mean = (float)n->ret_mean; N.ret_std = (float)n->ret_std;
    std::vector<float> blob;
    auto put = [&](const float* src, size_t count) -> int {
        const int off = (int)blob.size();
        if (src) blob.insert(blob.end(), src, src + count);
        else blob.insert(blob.end(), count, 0.f);
        return off;
    };
    N.aW1 = put(n->aW1, (size_t)N.d * N.h1a); N.ab1 = put(n->ab1, N.h1a);
    N.ag1 = put(ln ? n->ag1 : nullptr, N.h1a); N.abe1 = put(ln ? n->abe1 : nullptr, N.h1a);
    N.aW2 = put(n->aW2, (size_t)N.h1a * N.h2a); N.ab2 = put(n->ab2, N.h2a);
    N.ag2 = put(ln ? n->ag2 : nullptr, N.h2a); N.abe2 = put(ln ? n->abe2 : nullptr, N.h2a);
    N.aW3 = put(n->aW3, (size_t)N.h2a * N.da); N.ab3 = put(n->ab3, N.da);
    N.cW1 = put(n->cW1, (size_t)N.d * N.h1c); N.cb1 = put(n->cb1, N.h1c);
    N.cg1 = put(ln ? n->cg1 : nullptr, N.h1c); N.cbe1 = put(ln ? n->cbe1 : nullptr, N.h1c);
    N.cW2 = put(n->cW2, (size_t)(N.h1c + N.da) * N.h2c); N.cb2 = put(n->cb2, N.h2c);
    N.cg2 = put(ln ? n->cg2 : nullptr, N.h2c); N.cbe2 = put(ln ? n->cbe2 : nullptr, N.h2c);
    N.cW3 = put(n->cW3, N.h2c); N.cb3 = put(n->cb3, 1);
    N.obs_mean = (int)blob.size();
    for (int j = 0; j < N.d; ++j) blob.push_back(N.obs_norm ? (float)n->obs_mean[j] : 0.f);
    N.obs_inv_std = (int)blob.size();
    for (int j = 0; j < N.d; ++j) blob.push_back(N.obs_norm ? (float)(1.0 / n->obs_std[j]) : 1.f);
    N.total = (int)blob.size();
    const int hmax = std::max(std::max(N.h1a, N.h2a), std::max(N.h1c, N.h2c)) + N.da;
    if (((size_t)N.total + 4 * (size_t)(N.d + 2 * hmax + N.da)) * 4 > 200 * 1024)
        SS_FAIL(c, SS_EUNSUPPORTED, "value net: parameters do not fit in shared memory");
    SS_CUDA_CHECK(c, c->value_net_params.ensure(blob.size() * 4));
    SS_CUDA_CHECK(c, cudaMemcpyAsync(c->value_net_params.p, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, c->stream));
    SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    c->value_net_desc.assign(reinterpret_cast<const char*>(&N), reinterpret_cast<const char*>(&N) + sizeof(N));
    c->value_net_set = true;
    return SS_OK;
}

extern "C" int ss_value_net_eval(ss_ctx* c, const double* queries, int64_t m, int d, float* out_values) {
    if (!c) return SS_EINVAL;
    if (!c->value_net_set) SS_FAIL(c, SS_ESTATE, "value net: ss_value_net_set first");
    const ValueNetDev& N = *reinterpret_cast<const ValueNetDev*>(c->value_net_desc.data());
    if (!queries || !out_values || m < 1 || d != N.d) SS_FAIL(c, SS_EINVAL, "value net: bad arguments");
    SS_CUDA_CHECK(c, cudaSetDevice(c->device));
    SS_CUDA_CHECK(c, c->kde_q64.ensure((size_t)m * d * 8));
    SS_CUDA_CHECK(c, c->kde_vals.ensure((size_t)m * 4));
    SS_CUDA_CHECK(c, cudaMemcpyAsync(c->kde_q64.p, queries, (size_t)m * d * 8, cudaMemcpyHostToDevice, c->stream));
    int rc = value_net_eval_dev(c, c->kde_q64.as<double>(), m, c->kde_vals.as<float>());
    if (rc) return rc;
    SS_CUDA_CHECK(c, cudaMemcpyAsync(out_values, c->kde_vals.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    SS_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return SS_OK;
}
