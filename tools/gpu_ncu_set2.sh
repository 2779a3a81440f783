#!/bin/bash
# Launch list + one full ncu capture of the default FP32 kernel set (one timed step of 50k QM9-shaped molecules).
#   gpurun --timeout 600 -- 'bash tools/gpu_ncu_set2.sh [extra bench flags]'
mkdir -p gpurun_out
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --secondary 0 $@"
$CMD > gpurun_out/plain_set2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_set2.csv $CMD > gpurun_out/ncu1_set2.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain_set2b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'bundle_run_kernel|bundle_const_kernel|atom_kernel' -s 21 -c 21 -f -o gpurun_out/prof_set2 $CMD > gpurun_out/ncu2_set2.log 2>&1
echo "ncu full rc=$?"
tail -1 gpurun_out/plain_set2.log | cut -c1-300
