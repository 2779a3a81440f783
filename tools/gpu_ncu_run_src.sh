#!/bin/bash
# Source-level ncu capture of the row-run GNN kernel (one collapsed + one live launch of a 50 k-molecule inference) and
# the launch list of the same command:  gpurun --timeout 600 -- 'bash tools/gpu_ncu_run_src.sh'
mkdir -p gpurun_out
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --secondary 0"
$CMD > gpurun_out/plain_run_src.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_run_src.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bundle_run -s 7 -c 2 -f -o gpurun_out/run_src $CMD > gpurun_out/ncu_run_src.log 2>&1
echo "ncu full rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launches_final.log 2>&1
echo "ncu list rc=$?"
ls -la gpurun_out | tail -5
