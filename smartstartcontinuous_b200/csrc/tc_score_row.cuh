// tc_score_row.cuh -- the scoring step of a row thread, shared by the tcgen05 rollout kernels.
#pragma once

#include "mpc_kernels.cuh"

namespace tc {

// score one trajectory point for this row (ch-1 thread): waypoint move + progress (+ per-sample
// penalty); in reference mode the (state, waypoint index) row is spilled for the penalty passes
// qcol >= 0: this warp's column (tile * 4 + row warp) of the projection-sum table a.qsums
template <int DT>
__device__ __forceinline__ void score_row(const RolloutArgs& a, int t, const float (&x)[DT], ScoreAcc& sc,
                                          bool live, long long k_local, long long qcol, long long n_qcols, int lane) {
    score_point<DT>(a.plan, t, x, sc, a.per_sample != 0);
    if (a.states_out && live) traj_store<DT>(a.states_out, (size_t)t * a.K_local + k_local, a.d, x, sc.idx);
    if (a.qsums && qcol >= 0) {
        // reference penalty: a'.b' and b'.b' of this step (numerical.py:89-93) summed over the warp's
        // 32 sequences -- in the shadow of the layer-2 MMAs, so the separate pass over the spilled rows
        // is not needed.  The 32-term warp sum runs in FP32 (FP64 shuffles + DADDs in these warps cost
        // the kernel 10 %: measured), everything above it in float64; one table column per (tile, row
        // warp) keeps the final sum independent of how tiles are spread over CTAs and launches.
        float ab = 0.f, bb = 0.f;
        if (live) proj_terms<DT>(a.plan, sc.idx, x, ab, bb);
        for (int off = 16; off > 0; off >>= 1) {
            ab += __shfl_down_sync(0xffffffffu, ab, off);
            bb += __shfl_down_sync(0xffffffffu, bb, off);
        }
        if (lane == 0) {
            a.qsums[((size_t)t * 2) * n_qcols + qcol] = (double)ab;
            a.qsums[((size_t)t * 2 + 1) * n_qcols + qcol] = (double)bb;
        }
    }
}

}  // namespace tc
