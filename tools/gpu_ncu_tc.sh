#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload protein --atoms 20000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --gnn-far-tensor 1"
$CMD > gpurun_out/plain_tc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gnn_far_tc -s 5 -c 1 -f -o gpurun_out/prof_tc $CMD > gpurun_out/ncu_tc.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_tc.log
