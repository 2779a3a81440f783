#!/bin/bash
# A/B of the FP32 kernel sets (pair_const 0 / 1 / 2) + their parity tests.   gpurun --timeout 900 -- 'bash tools/gpu_ab_sets.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pair_const.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_sets.log
run() { timeout 150 python bench.py --molecules 300000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --secondary 0 "$@" 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['checks'])"; }
for S in 0 1 2; do echo "== pair_const $S decay"; run --pair-const $S; done
for S in 0 2; do echo "== pair_const $S model_weights"; run --pair-const $S --checkpoint model_weights; done
echo "== protein 40k, set 2"; run --workload protein --atoms 40000 --steps 3
