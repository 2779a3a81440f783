#!/bin/bash
# Second GPU call of round 2 (after tools/gpu_round2_first.sh is green): launch list + one full ncu capture of the
# pair-per-thread kernels, same recipe as tools/gpu_ncu.sh.
#   gpurun --timeout 500 -- 'bash tools/gpu_round2_ncu.sh'
mkdir -p gpurun_out
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --pair-const 1"
$CMD > gpurun_out/plain_const.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_const.csv $CMD > gpurun_out/ncu1_const.log 2>&1
echo "ncu list rc=$?"
# timed step = launches 10..19 of bundle_const_kernel (5 GNN + 5 EPN); one collapsed GNN step, one live GNN step, one EPN pass, plus the atom kernel
$CMD > gpurun_out/plain_const2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'bundle_const_kernel|atom_const_kernel' -s 22 -c 8 -f -o gpurun_out/prof_const $CMD > gpurun_out/ncu2_const.log 2>&1
echo "ncu full rc=$?"
tail -2 gpurun_out/plain_const.log | cut -c1-400
