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
// FP32 like the TF graph, all parameters in shared memory (rows padded to 16 floats), layer norm =
// tf.contrib.layers.layer_norm (biased variance, eps 1e-12).  Two kernels:
//   value_net_tile_kernel   hidden widths <= 64 (the reference's nets are 64 / 64 and 64 / 32): four threads per
//                           pair of queries, each thread a 2 x 16 register tile of (query, unit) accumulators fed by
//                           one 8-byte activation load and four 16-byte weight loads per input -- FMA-bound;
//   value_net_kernel        any width: one warp per query, units spread over the lanes (2 loads per FMA).
// Both accumulate bias + sum_k in[k] W[k][u] in k order, so they agree bit for bit without layer norm.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

struct ValueNetDev {
    int d, da, h1a, h2a, h1c, h2c, layer_norm, last_tanh, obs_norm, ret_norm;
    // offsets (floats) into the parameter block
    int aW1, ab1, ag1, abe1, aW2, ab2, ag2, abe2, aW3, ab3;
    int cW1, cb1, cg1, cbe1, cW2, cb2, cg2, cbe2, cW3, cb3;
    int obs_mean, obs_inv_std, total;
    int s_a1, s_a2, s_a3, s_c1, s_c2, s_c3;        // row strides (floats; multiples of 16 for the tiled kernel)
    float obs_lo, obs_hi, ret_lo, ret_hi, ret_mean, ret_std;
};

// out[u] = b[u] + sum_k in[k] * W[k][u] for the lane's units u = lane, lane + 32, ...
__device__ __forceinline__ void dense(const float* __restrict__ in, int n_in, const float* __restrict__ W, int stride,
                                      const float* __restrict__ b, int n_out, float* __restrict__ out, int lane) {
    for (int u = lane; u < n_out; u += 32) {
        float acc = b[u];
        for (int k = 0; k < n_in; ++k) acc = fmaf(in[k], W[k * stride + u], acc);
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
        dense(xs, N.d, P + N.aW1, N.s_a1, P + N.ab1, N.h1a, bufA, lane);
        norm_act(bufA, N.h1a, N.layer_norm, P + N.ag1, P + N.abe1, 0, lane);
        dense(bufA, N.h1a, P + N.aW2, N.s_a2, P + N.ab2, N.h2a, bufB, lane);
        norm_act(bufB, N.h2a, N.layer_norm, P + N.ag2, P + N.abe2, N.last_tanh, lane);
        dense(bufB, N.h2a, P + N.aW3, N.s_a3, P + N.ab3, N.da, act, lane);
        for (int u = lane; u < N.da; u += 32) act[u] = tanhf(act[u]);
        __syncwarp();
        // critic
        dense(xs, N.d, P + N.cW1, N.s_c1, P + N.cb1, N.h1c, bufA, lane);
        norm_act(bufA, N.h1c, N.layer_norm, P + N.cg1, P + N.cbe1, 0, lane);
        for (int u = lane; u < N.da; u += 32) bufA[N.h1c + u] = act[u];      // concat([h, action])
        __syncwarp();
        dense(bufA, N.h1c + N.da, P + N.cW2, N.s_c2, P + N.cb2, N.h2c, bufB, lane);
        norm_act(bufB, N.h2c, N.layer_norm, P + N.cg2, P + N.cbe2, N.last_tanh, lane);
        float v = 0.f;
        for (int k = lane; k < N.h2c; k += 32) v = fmaf(bufB[k], P[N.cW3 + k * N.s_c3], v);
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

// ---- register-tiled kernel for hidden widths <= 64 ------------------------------------------------
constexpr int VT_THREADS = 128;            // 32 groups of 4 threads
constexpr int VT_QPB = VT_THREADS / 2;     // 64 queries per block: a group handles 2 queries
constexpr int VT_HP = 64;                  // widest layer; 16 units per thread
constexpr int VT_ROWS_A = VT_HP + 16;      // critic layer 2 reads [h1c | action]

// acc[q][4 v + i] = unit 16 v + 4 r + i of query q of the group (r = thread of the group): the four threads of a
// group read one contiguous 64-byte piece of a weight row, all groups of a warp the same one (a broadcast)
struct VtAcc { float a[2][16]; };

// one copy of tanhf in the kernel: inlined at its 100+ call sites of the unrolled tiles it makes the kernel larger
// than the instruction cache (measured: 38 us, "no instruction" the top stall)
__device__ __noinline__ float vt_tanh(float x) { return tanhf(x); }

__device__ __forceinline__ void vt_dense(const float* __restrict__ in, int n_in, const float* __restrict__ W,
                                         int stride, const float* __restrict__ b, int g, int r, VtAcc& acc) {
    const int nv = stride >> 4;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (v < nv) b4 = *reinterpret_cast<const float4*>(b + 16 * v + 4 * r);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            acc.a[q][4 * v] = b4.x; acc.a[q][4 * v + 1] = b4.y; acc.a[q][4 * v + 2] = b4.z; acc.a[q][4 * v + 3] = b4.w;
        }
    }
    const float* wr = W + 4 * r;
    const float* xr = in + 2 * g;
#pragma unroll 4
    for (int k = 0; k < n_in; ++k) {
        const float2 x = *reinterpret_cast<const float2*>(xr + k * VT_QPB);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v < nv) {
                const float4 w = *reinterpret_cast<const float4*>(wr + k * stride + 16 * v);
                acc.a[0][4 * v] = fmaf(x.x, w.x, acc.a[0][4 * v]);
                acc.a[0][4 * v + 1] = fmaf(x.x, w.y, acc.a[0][4 * v + 1]);
                acc.a[0][4 * v + 2] = fmaf(x.x, w.z, acc.a[0][4 * v + 2]);
                acc.a[0][4 * v + 3] = fmaf(x.x, w.w, acc.a[0][4 * v + 3]);
                acc.a[1][4 * v] = fmaf(x.y, w.x, acc.a[1][4 * v]);
                acc.a[1][4 * v + 1] = fmaf(x.y, w.y, acc.a[1][4 * v + 1]);
                acc.a[1][4 * v + 2] = fmaf(x.y, w.z, acc.a[1][4 * v + 2]);
                acc.a[1][4 * v + 3] = fmaf(x.y, w.w, acc.a[1][4 * v + 3]);
            }
        }
    }
}

// optional layer norm over the n units of each query (sums over the thread's units, then over the group's four
// threads), relu / tanh, and the activations written as rows of `out` for the next layer
__device__ __forceinline__ void vt_norm_act_store(VtAcc& acc, int n, bool layer_norm, const float* __restrict__ gm,
                                                  const float* __restrict__ be, int act, float* __restrict__ out,
                                                  int g, int r) {
    float mean[2] = {0.f, 0.f}, rstd[2] = {1.f, 1.f};
    if (layer_norm) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float sm_ = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int u = 16 * (i >> 2) + 4 * r + (i & 3);
                if (u < n) sm_ += acc.a[q][i];
            }
            sm_ += __shfl_xor_sync(0xffffffffu, sm_, 1);
            sm_ += __shfl_xor_sync(0xffffffffu, sm_, 2);
            mean[q] = sm_ / n;
            float sq = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int u = 16 * (i >> 2) + 4 * r + (i & 3);
                const float df = acc.a[q][i] - mean[q];
                if (u < n) sq = fmaf(df, df, sq);
            }
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            rstd[q] = rsqrtf(sq / n + 1e-12f);
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int u = 16 * (i >> 2) + 4 * r + (i & 3);
        if (u < n) {
            float y[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float x = acc.a[q][i];
                if (layer_norm) x = (x - mean[q]) * rstd[q] * gm[u] + be[u];
                y[q] = act ? vt_tanh(x) : fmaxf(x, 0.f);
            }
            *reinterpret_cast<float2*>(out + u * VT_QPB + 2 * g) = make_float2(y[0], y[1]);
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(VT_THREADS)
value_net_tile_kernel(const ValueNetDev N, const float* __restrict__ params, const double* __restrict__ queries,
                      long long m, float* __restrict__ values) {
    extern __shared__ __align__(16) float sm[];
    float* P = sm;                                         // parameters (total is a multiple of 16 floats):
    __shared__ __align__(8) uint64_t p_full;               // one bulk copy, a single round trip to L2
    if (threadIdx.x == 0) {
        tc::mbar_init(&p_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tc::mbar_expect_tx(&p_full, (uint32_t)N.total * 4u);
        tc::bulk_g2s(P, params, (uint32_t)N.total * 4u, &p_full);
    }
    float* xs = sm + N.total;                              // [d][64] observations, query-minor
    float* bufA = xs + N.d * VT_QPB;                       // [80][64]
    float* bufB = bufA + VT_ROWS_A * VT_QPB;               // [64][64]
    const int g = threadIdx.x >> 2, r = threadIdx.x & 3;
    const long long q0 = (long long)blockIdx.x * VT_QPB + 2 * g;
    for (int j = r; j < N.d; j += 4) {
        float x[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long qq = q0 + q < m ? q0 + q : m - 1;          // the tail repeats the last query
            x[q] = (float)queries[qq * N.d + j];
        }
        xs[j * VT_QPB + 2 * g] = x[0];
        xs[j * VT_QPB + 2 * g + 1] = x[1];
    }
    __syncthreads();
    tc::mbar_wait<false>(&p_full, 0);
    for (int j = r; j < N.d; j += 4) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float x = xs[j * VT_QPB + 2 * g + q];
            x = fminf(fmaxf((x - P[N.obs_mean + j]) * P[N.obs_inv_std + j], N.obs_lo), N.obs_hi);
            xs[j * VT_QPB + 2 * g + q] = x;
        }
    }
    __syncwarp();
    // the six layers run through ONE copy of the tile code (a loop over layer descriptors): unrolled per layer the
    // kernel is 100 KB of instructions and stalls on instruction fetch
    VtAcc acc;
#pragma unroll 1
    for (int layer = 0; layer < 6; ++layer) {
        const float *in, *W, *b, *gm = P, *be = P;
        float* out;
        int n_in, stride, n, act;                          // act: 0 relu, 1 tanh, 2 none (kept in registers)
        bool ln = N.layer_norm != 0;
        switch (layer) {
        case 0:  in = xs;   n_in = N.d;   W = P + N.aW1; stride = N.s_a1; b = P + N.ab1; n = N.h1a; act = 0;
                 gm = P + N.ag1; be = P + N.abe1; out = bufA; break;
        case 1:  in = bufA; n_in = N.h1a; W = P + N.aW2; stride = N.s_a2; b = P + N.ab2; n = N.h2a; act = N.last_tanh;
                 gm = P + N.ag2; be = P + N.abe2; out = bufB; break;
        case 2:  // the action lands where critic layer 2 reads it: rows h1c .. of bufA (actor h1 is dead by now,
                 // critic layer 1 only writes rows < h1c) = concat([h, action])
                 in = bufB; n_in = N.h2a; W = P + N.aW3; stride = N.s_a3; b = P + N.ab3; n = N.da; act = 1; ln = false;
                 out = bufA + N.h1c * VT_QPB; break;
        case 3:  in = xs;   n_in = N.d;   W = P + N.cW1; stride = N.s_c1; b = P + N.cb1; n = N.h1c; act = 0;
                 gm = P + N.cg1; be = P + N.cbe1; out = bufA; break;
        case 4:  in = bufA; n_in = N.h1c + N.da; W = P + N.cW2; stride = N.s_c2; b = P + N.cb2; n = N.h2c;
                 act = N.last_tanh; gm = P + N.cg2; be = P + N.cbe2; out = bufB; break;
        default: in = bufB; n_in = N.h2c; W = P + N.cW3; stride = N.s_c3; b = P + N.cb3; n = 0; act = 2; ln = false;
                 out = bufA; break;
        }
        vt_dense(in, n_in, W, stride, b, g, r, acc);
        vt_norm_act_store(acc, n, ln, gm, be, act, out, g, r);
    }
    if (r == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            // denormalize(clip_by_value(critic, return_range), ret_rms): the clip is unconditional, the affine
            // map only exists with return normalisation (ddpg_editted.py:130-131)
            float v = fminf(fmaxf(acc.a[q][0], N.ret_lo), N.ret_hi);
            if (N.ret_norm) v = v * N.ret_std + N.ret_mean;
            if (q0 + q < m) values[q0 + q] = v;
        }
    }
}

size_t vt_smem_bytes(const ValueNetDev& N) {
    return ((size_t)N.total + (size_t)(N.d + VT_ROWS_A + VT_HP) * VT_QPB) * 4;
}
bool vt_supported(const ValueNetDev& N) {
    return N.h1a <= VT_HP && N.h2a <= VT_HP && N.h1c <= VT_HP && N.h2c <= VT_HP && N.da <= 16 &&
           N.s_a1 % 16 == 0 && vt_smem_bytes(N) <= 200 * 1024 && !getenv("SS_VALUE_NET_GENERAL");
}

}  // namespace

struct ValueNetHost {
    ValueNetDev dev;
    size_t smem = 0;
};

int value_net_eval_dev(ss_ctx* c, const double* queries_dev, long long m, float* values_dev) {
    if (!c->value_net_set) SS_FAIL(c, SS_ESTATE, "value net: ss_value_net_set first");
    const ValueNetDev& N = *reinterpret_cast<const ValueNetDev*>(c->value_net_desc.data());
    if (vt_supported(N)) {
        const size_t smem = vt_smem_bytes(N);
        SS_CUDA_CHECK(c, cudaFuncSetAttribute(value_net_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        value_net_tile_kernel<<<(unsigned)((m + VT_QPB - 1) / VT_QPB), VT_THREADS, smem, c->stream>>>(
            N, c->value_net_params.as<float>(), queries_dev, m, values_dev);
        c->launches++;
        SS_CUDA_CHECK(c, cudaGetLastError());
        return SS_OK;
    }
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
    // every array starts on a 16-float boundary; matrices are stored with rows padded to 16 floats (zeros), so
    // that the tiled kernel reads whole 16-byte pieces and never needs a column bound
    // (only for nets the tiled kernel takes: the general kernel keeps the dense layout, which is what lets a
    // 200 / 100 net fit in shared memory)
    std::vector<float> blob;
    const bool tiled = N.h1a <= VT_HP && N.h2a <= VT_HP && N.h1c <= VT_HP && N.h2c <= VT_HP && N.da <= 16;
    auto pad16 = [tiled](int n) { return tiled ? (n + 15) / 16 * 16 : n; };
    auto put_mat = [&](const float* src, int rows, int cols, int* stride) -> int {
        const int off = (int)blob.size();
        *stride = pad16(cols);
        blob.resize(blob.size() + (size_t)rows * *stride, 0.f);
        for (int k = 0; k < rows; ++k)
            for (int u = 0; u < cols; ++u) blob[(size_t)off + (size_t)k * *stride + u] = src[(size_t)k * cols + u];
        return off;
    };
    auto put = [&](const float* src, size_t count) -> int {
        const int off = (int)blob.size();
        blob.resize(blob.size() + (size_t)pad16((int)count), 0.f);
        if (src) std::copy(src, src + count, blob.begin() + off);
        return off;
    };
    N.aW1 = put_mat(n->aW1, N.d, N.h1a, &N.s_a1); N.ab1 = put(n->ab1, N.h1a);
    N.ag1 = put(ln ? n->ag1 : nullptr, N.h1a); N.abe1 = put(ln ? n->abe1 : nullptr, N.h1a);
    N.aW2 = put_mat(n->aW2, N.h1a, N.h2a, &N.s_a2); N.ab2 = put(n->ab2, N.h2a);
    N.ag2 = put(ln ? n->ag2 : nullptr, N.h2a); N.abe2 = put(ln ? n->abe2 : nullptr, N.h2a);
    N.aW3 = put_mat(n->aW3, N.h2a, N.da, &N.s_a3); N.ab3 = put(n->ab3, N.da);
    N.cW1 = put_mat(n->cW1, N.d, N.h1c, &N.s_c1); N.cb1 = put(n->cb1, N.h1c);
    N.cg1 = put(ln ? n->cg1 : nullptr, N.h1c); N.cbe1 = put(ln ? n->cbe1 : nullptr, N.h1c);
    N.cW2 = put_mat(n->cW2, N.h1c + N.da, N.h2c, &N.s_c2); N.cb2 = put(n->cb2, N.h2c);
    N.cg2 = put(ln ? n->cg2 : nullptr, N.h2c); N.cbe2 = put(ln ? n->cbe2 : nullptr, N.h2c);
    N.cW3 = put_mat(n->cW3, N.h2c, 1, &N.s_c3); N.cb3 = put(n->cb3, 1);
    N.obs_mean = (int)blob.size();
    for (int j = 0; j < N.d; ++j) blob.push_back(N.obs_norm ? (float)n->obs_mean[j] : 0.f);
    blob.resize((size_t)pad16((int)blob.size()), 0.f);
    N.obs_inv_std = (int)blob.size();
    for (int j = 0; j < N.d; ++j) blob.push_back(N.obs_norm ? (float)(1.0 / n->obs_std[j]) : 1.f);
    blob.resize((size_t)pad16((int)blob.size()), 0.f);
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
