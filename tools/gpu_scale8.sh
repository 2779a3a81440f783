#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
N=${1:-8}
python bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_qm9_n$N.json 2>gpurun_out/bq$N.err; tail -1 gpurun_out/bench_qm9_n$N.json | cut -c1-300; tail -2 gpurun_out/bq$N.err
python bench.py --gpus $N --workload protein --atoms 100000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_protein100k_n$N.json 2>gpurun_out/bp$N.err; tail -1 gpurun_out/bench_protein100k_n$N.json | cut -c1-300; tail -2 gpurun_out/bp$N.err
python bench.py --gpus 1 --workload protein --atoms 100000 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_protein100k_n1.json 2>gpurun_out/bp1.err; tail -1 gpurun_out/bench_protein100k_n1.json | cut -c1-300; tail -2 gpurun_out/bp1.err
