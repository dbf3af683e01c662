// mpc_tc.cu -- tcgen05 rollout kernel (placeholder until the kernel lands in this round).
#include "mpc_kernels.cuh"

bool mpc_tc_shape_supported(const ss_ctx*) { return false; }
int mpc_tc_prepare(ss_ctx*) { return SS_OK; }
int mpc_tc_grid(const ss_ctx*, const RolloutArgs&) { return 1; }
int mpc_tc_launch(ss_ctx* c, const RolloutArgs&, int*) {
    SS_FAIL(c, SS_EUNSUPPORTED, "mpc: tcgen05 kernel not available");
}
