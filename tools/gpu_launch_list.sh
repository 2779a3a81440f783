#!/bin/bash
# Default bench line (N = 1) followed by the ncu launch list of one 50 k-molecule step (per-kernel times; cold-cache, serialised):
#   gpurun --timeout 1200 -- 'bash tools/gpu_launch_list.sh'
mkdir -p gpurun_out
bash tools/gpu_bench_default.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv python bench.py --molecules 50000 --steps 1 --warmup 1 --secondary 0 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu rc=$?"
