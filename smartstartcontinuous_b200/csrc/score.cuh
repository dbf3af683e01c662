// score.cuh -- per-(sequence, time-step) trajectory scoring shared by every rollout kernel.
//
// Restates NND_MB_agent.generate_scores_add_delta (NND_MB_agent.py:566-628) with
// move_to_next (:491-496), the elliptical distance (numerical.py:116-124) and
// dist_line_seg_to_point / projection_of_a_onto_b (numerical.py:74-98) for ONE sample.
// The reference's projection coefficient is global over the K samples of a time step
// (np.sum without axis, numerical.py:89-93): in SS_PENALTY_REFERENCE mode score_point()
// only returns the two dot products a'.b' and b'.b' (to be summed over all samples and
// GPUs, from the spilled rows) and the penalty is applied afterwards by penalty_with_lambda(); in
// SS_PENALTY_PER_SAMPLE mode the coefficient is local and everything fuses.
#pragma once

#include "common.cuh"

struct PlanView {
    const float* ds;        // desired_states [W][d]
    const float* dl;        // distances_left [W]
    const float* gpow;      // gamma^t, t = 0..H   (rounded from float64 on the host)
    int W, d;
    float inv_r[SS_MAX_D];  // 1 / radii
    float pen_scale;        // horizontal_penalty_factor * gamma   (NND_MB_agent.py:622)
};

struct ScoreAcc {
    int idx;       // samples_desired_state_indices[k]
    float prev;    // prev_distances_to_end[k]
    float score;   // scores[k]
};

template <int DT>
__device__ __forceinline__ float ell_dist(const PlanView& P, const float* __restrict__ w,
                                          const float (&x)[DT]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < DT; ++j)
        if (j < P.d) {
            float df = (w[j] - x[j]) * P.inv_r[j];
            s = fmaf(df, df, s);
        }
    return sqrtf(s);
}

// NND_MB_agent.py:571-577
template <int DT>
__device__ __forceinline__ void score_init(const PlanView& P, int wp_index, const float (&x0)[DT],
                                           ScoreAcc& a) {
    a.idx = wp_index;
    a.prev = P.dl[wp_index] + ell_dist<DT>(P, P.ds + (size_t)wp_index * P.d, x0);
    a.score = 0.f;
}

// Spilled trajectories (reference penalty mode): row (t, k) = d state floats followed by the
// sample's waypoint index after the move of that step, so the projection sums and the penalties
// can be recomputed for every (t, k) independently (no serial waypoint scan in the tail passes).
__host__ __device__ __forceinline__ int traj_row_stride(int d) { return d + 1; }

template <int DT>
__device__ __forceinline__ void traj_store(float* __restrict__ rows, size_t row, int d, const float (&x)[DT],
                                           int idx) {
    if (DT >= 3 && d == 3) {
        *reinterpret_cast<float4*>(rows + row * 4) = make_float4(x[0], x[1], x[2], __int_as_float(idx));
    } else {
        float* o = rows + row * (size_t)(d + 1);
#pragma unroll
        for (int j = 0; j < DT; ++j)
            if (j < d) o[j] = x[j];
        o[d] = __int_as_float(idx);
    }
}
template <int DT>
__device__ __forceinline__ void traj_load(const float* __restrict__ rows, size_t row, int d, float (&x)[DT],
                                          int& idx) {
    if (DT >= 3 && d == 3) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(rows + row * 4));
        x[0] = v.x; x[1] = v.y; x[2] = v.z;
        idx = __float_as_int(v.w);
    } else {
        const float* o = rows + row * (size_t)(d + 1);
#pragma unroll
        for (int j = 0; j < DT; ++j)
            if (j < d) x[j] = __ldg(o + j);
        idx = __float_as_int(__ldg(o + d));
    }
}

// a'.b' and b'.b' of the line-segment projection (NND_MB_agent.py:616-621, numerical.py:89-93)
// for a point whose waypoint index after the move is idx
template <int DT>
__device__ __forceinline__ void proj_terms(const PlanView& P, int idx, const float (&x)[DT], float& ab, float& bb) {
    const int b0 = idx - 1 > 0 ? idx - 1 : 0;
    const float* w0 = P.ds + (size_t)b0 * P.d;
    const float* w1 = w0 + P.d;
    ab = 0.f;
    bb = 0.f;
#pragma unroll
    for (int j = 0; j < DT; ++j)
        if (j < P.d) {
            const float av = (x[j] - w0[j]) * P.inv_r[j];
            const float bv = (w1[j] - w0[j]) * P.inv_r[j];
            ab = fmaf(av, bv, ab);
            bb = fmaf(bv, bv, bb);
        }
}

// One trajectory point (NND_MB_agent.py:582-622): waypoint move + progress term; when per_sample
// is true the penalty (per-sample projection coefficient) is applied here as well, otherwise it
// is left to the reference-mode passes (which recompute a'.b', b'.b' from the spilled row).
template <int DT>
__device__ __forceinline__ void score_point(const PlanView& P, int t, const float (&x)[DT],
                                            ScoreAcc& a, bool per_sample) {
    const int last = P.W - 1;
    const int nxt = a.idx + 1 < last ? a.idx + 1 : last;
    float dc = ell_dist<DT>(P, P.ds + (size_t)a.idx * P.d, x);
    const float dn = ell_dist<DT>(P, P.ds + (size_t)nxt * P.d, x);
    const bool mv = ((dc <= 1.0f) || (dn <= dc)) && (a.idx != last);     // theta == 1
    if (mv) { a.idx += 1; dc = dn; }
    const float to_end = P.dl[a.idx] + dc;
    a.score = fmaf(a.prev - to_end, P.gpow[t], a.score);
    a.prev = to_end;
    if (per_sample) {
        const int b0 = a.idx - 1 > 0 ? a.idx - 1 : 0;
        const float* w0 = P.ds + (size_t)b0 * P.d;
        const float* w1 = w0 + P.d;
        float av[DT], bv[DT];
        float ab = 0.f, bb = 0.f;
#pragma unroll
        for (int j = 0; j < DT; ++j)
            if (j < P.d) {
                av[j] = (x[j] - w0[j]) * P.inv_r[j];
                bv[j] = (w1[j] - w0[j]) * P.inv_r[j];
                ab = fmaf(av[j], bv[j], ab);
                bb = fmaf(bv[j], bv[j], bb);
            }
        const float lam = ab / bb;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < DT; ++j)
            if (j < P.d) {
                float df = fmaf(lam, bv[j], -av[j]);
                s = fmaf(df, df, s);
            }
        a.score -= sqrtf(s) * P.pen_scale;
    }
}

// reference mode, second pass: penalty of one point given the global coefficient of its step
template <int DT>
__device__ __forceinline__ float penalty_with_lambda(const PlanView& P, int idx_after_move,
                                                     const float (&x)[DT], float lam) {
    const int b0 = idx_after_move - 1 > 0 ? idx_after_move - 1 : 0;
    const float* w0 = P.ds + (size_t)b0 * P.d;
    const float* w1 = w0 + P.d;
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < DT; ++j)
        if (j < P.d) {
            float av = (x[j] - w0[j]) * P.inv_r[j];
            float bv = (w1[j] - w0[j]) * P.inv_r[j];
            float df = fmaf(lam, bv, -av);
            s = fmaf(df, df, s);
        }
    return sqrtf(s) * P.pen_scale;
}
