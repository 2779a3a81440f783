#!/bin/bash
mkdir -p gpurun_out
for A in 2220 40000; do for TC in 0 1; do
  timeout 200 python bench.py --workload protein --atoms $A --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gnn-far-tensor $TC 2>/dev/null | grep '^{' > gpurun_out/bench_protein${A}_tc$TC.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_protein${A}_tc$TC.json')); print('atoms', $A, 'tc', $TC, 'ms/step', round(d['ms_per_step'],3), 'gnn_pair', round(d['phases_ms_per_step']['ms_gnn_pair'],3), 'atoms/s', round(d['value']), 'sumq', d['checks'])"
done; done
