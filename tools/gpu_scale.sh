#!/bin/bash
# driver-style scaling check: the default bench at several N on one box
mkdir -p gpurun_out
for N in "$@"; do
  python bench.py --gpus $N --no-cpu-baseline 2>gpurun_out/bs$N.err | grep '^{' > gpurun_out/bench_qm9_final_n$N.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_qm9_final_n$N.json')); print('N', d['n_gpus'], 'atoms/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['clocks']['reasons'])"
done
