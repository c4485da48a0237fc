// ssa_l16.cu -- the shared-memory SSA kernel for tiles of 16 lane(s) per replicate (see ssa_kernel.cuh).
#include "engine.cuh"
ECDNA_DEFINE_LAUNCH_SMEM(16)
