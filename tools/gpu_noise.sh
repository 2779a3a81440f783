#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/measure_noise_floor.py 400 2>&1 | tee gpurun_out/noise_floor.log
for P in 32 48 64; do echo "== 300k molecules precision $P model2_weights"; timeout 200 python bench.py --molecules 300000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --secondary 0 --precision $P --checkpoint model2_weights 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})"; done 2>&1 | tee gpurun_out/precision_bench.log
