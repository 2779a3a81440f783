#!/bin/bash
# ncu launch list + one full capture of kernels matching $1 (regex), skipping $2 matching launches, capturing $3
mkdir -p gpurun_out
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 -f -o gpurun_out/prof_pair $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"
tail -3 gpurun_out/plain.log | cut -c1-600
