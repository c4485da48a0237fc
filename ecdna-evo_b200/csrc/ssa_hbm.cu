// ssa_hbm.cu -- the SSA kernel with the histogram resident in HBM (one warp per replicate): BASELINE config 5's
// "HBM-resident state path", and the second launch that resumes replicates parked by the shared-memory launch.
#include "engine.cuh"

namespace ecdna {

int launch_hbm(ecdna_b200_ctx* ctx, SsaArgs& a, cudaStream_t st, bool replay, uint32_t* grid_out, uint32_t* bps_out) {
  return replay ? launch_kernel<32, true, true, 0>(ctx, a, st, a.n_runs, grid_out, bps_out, 0)
                : launch_kernel<32, true, false, 0>(ctx, a, st, a.n_runs, grid_out, bps_out, 0);
}

}  // namespace ecdna
