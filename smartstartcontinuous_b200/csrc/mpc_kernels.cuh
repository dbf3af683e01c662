// mpc_kernels.cuh -- declarations shared between the MPC host code and the kernel files.
#pragma once

#include "common.cuh"
#include "score.cuh"

// everything a rollout kernel needs, passed by value
struct RolloutArgs {
    // model (fp32, zero padded): w[l] is [in_pad_l][out_pad_l] row-major
    const float* w[SS_MAX_LAYERS + 1];
    const float* b[SS_MAX_LAYERS + 1];
    int d, da, L, h;
    int din_pad, h_pad, dout_pad;
    MpcNorm norm;
    ActionSource act;
    PlanView plan;
    float state0[SS_MAX_D];
    int wp_index, H;
    long long K_local, k_offset;
    int per_sample;          // 1: score in the rollout kernel with the per-sample penalty;
                             // 0: reference mode -- the kernel only spills the trajectories, the
                             //    scoring passes (mpc_score.cu) run over them afterwards
    float* states_out;       // [H+1][K_local][d] or null
    double* partial_sums;    // legacy in-kernel projection sums [gridDim.x][H+1][2]; null = off
    float* scores_out;       // [K_local] (per-sample mode: final; reference mode: unused)
};

// fp32 SIMT rollout (mpc_simt.cu)
int mpc_simt_launch(ss_ctx* c, const RolloutArgs& a, int* grid_blocks_out);
int mpc_simt_grid(const RolloutArgs& a);

// tcgen05 rollout (mpc_tc.cu)
bool mpc_tc_shape_supported(const ss_ctx* c);
int mpc_tc_prepare(ss_ctx* c);      // builds the BF16 operand images after ss_mpc_set_model
int mpc_tc_launch(ss_ctx* c, const RolloutArgs& a, int* grid_blocks_out);
int mpc_tc_grid(const ss_ctx* c, const RolloutArgs& a);

// scoring tail (mpc_score.cu)
int mpc_reduce_sums(ss_ctx* c, const double* partial, int blocks, int T, double* sums);
int mpc_sums_reference_blocks(long long K_local);
int mpc_sums_reference(ss_ctx* c, const PlanView& plan, int wp_index, const float* state0, const float* states,
                       long long K_local, int T, double* partial);
int mpc_score_reference(ss_ctx* c, const PlanView& plan, int wp_index, const float* state0,
                        const float* states, long long K_local, int T, const double* sums,
                        float* scores);
int mpc_argmax(ss_ctx* c, const float* scores, long long K_local, long long k_offset,
               double* block_v, long long* block_i, void* result_dev);

struct MpcResult {
    double best_score;
    long long best_k;
    unsigned int blocks_done;
    int pad;
};
