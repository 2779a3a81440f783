#!/bin/bash
mkdir -p gpurun_out
./build/ubench > gpurun_out/ubench.log 2>&1; cat gpurun_out/ubench.log
for ca in 65536 262144 1048576 4194304 16777216; do
  echo "chunk_atoms=$ca"
  python bench.py --molecules 300000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --chunk-atoms $ca | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['phases_ms_per_step'])"
done 2>&1 | tee gpurun_out/chunk_sweep.log
