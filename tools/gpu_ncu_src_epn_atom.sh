#!/bin/bash
# Source-level ncu capture of one launch of the EPN bundle kernel and one update launch of the tensor per-atom kernel
# (50 k-molecule inference):   gpurun --timeout 600 -- 'bash tools/gpu_ncu_src_epn_atom.sh'
mkdir -p gpurun_out
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --secondary 0"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"bundle_const_kernel" -s 6 -c 1 -f -o gpurun_out/src_epn $CMD > gpurun_out/ncu_src_epn.log 2>&1; echo "rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"atom_mma_kernel" -s 12 -c 1 -f -o gpurun_out/src_atom $CMD > gpurun_out/ncu_src_atom.log 2>&1; echo "rc=$?"
