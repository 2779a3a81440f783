#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -5
N=${1:-2}
python bench.py --gpus 1 --workload protein --atoms 40000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_protein40k_n1.json 2>gpurun_out/bp1.err; cat gpurun_out/bench_protein40k_n1.json | cut -c1-400; tail -3 gpurun_out/bp1.err
python bench.py --gpus $N --workload protein --atoms 40000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_protein40k_n$N.json 2>gpurun_out/bpN.err; cat gpurun_out/bench_protein40k_n$N.json | cut -c1-400; tail -3 gpurun_out/bpN.err
python bench.py --gpus $N --molecules 300000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_qm9_n$N.json 2>gpurun_out/bqN.err; cat gpurun_out/bench_qm9_n$N.json | cut -c1-400; tail -3 gpurun_out/bqN.err
