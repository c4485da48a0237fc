#!/bin/bash
# compile only the L=4 native kernel quickly for SASS inspection: scripts/quick_sass.sh [extra nvcc flags]
cd /root/repo
cat > gpurun_out/quick.cu <<'EOC'
#include "../ecdna-evo_b200/csrc/ssa_kernel.cuh"
void* quick_ptr() { return (void*)ecdna::ssa_kernel<QUICK_L, false, false, QUICK_KG, QUICK_MINB>; }
EOC
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DQUICK_MINB=${MINB:-5} -DQUICK_L=${TILE:-4} -DQUICK_KG=${KG:-2} -cubin -o gpurun_out/quick.cubin gpurun_out/quick.cu "$@" && python scripts/sass_path.py gpurun_out/quick.cubin ssa_kernelILi${TILE:-4}ELb0ELb0ELi${KG:-2} ${TAKEN:-} | sed -n 1,${LINES_OUT:-12}p
