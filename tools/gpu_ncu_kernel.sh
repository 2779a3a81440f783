#!/bin/bash
# One full ncu capture of selected kernels of a 50k-molecule step:  gpurun -- 'bash tools/gpu_ncu_kernel.sh <regex> <skip> <count> <outname> [bench flags]'
mkdir -p gpurun_out
RX=$1; SKIP=$2; CNT=$3; OUT=$4; shift 4
CMD="python bench.py --molecules 50000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --secondary 0 $@"
$CMD > gpurun_out/plain_$OUT.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -f -o gpurun_out/$OUT $CMD > gpurun_out/ncu_$OUT.log 2>&1
echo "ncu full rc=$?"
