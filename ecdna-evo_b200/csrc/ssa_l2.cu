// ssa_l2.cu -- the shared-memory SSA kernel for tiles of 2 lane(s) per replicate (see ssa_kernel.cuh).
#include "engine.cuh"
ECDNA_DEFINE_LAUNCH_SMEM(2)
