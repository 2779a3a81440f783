#!/bin/bash
# First GPU call of round 2: validate the experimental pair-per-thread bundle kernels (pair_const) and measure them.
#   gpurun --timeout 600 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
EPNN_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_pair_const.py -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_pair_const.log
run() { timeout 150 python bench.py --molecules 300000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['checks'])"; }
echo "== default (FP32 SIMT, warp tile)"; run
echo "== --pair-const 1 (FP32 SIMT, pair per thread, uniform weights)"; run --pair-const 1
echo "== --pair-tensor 1 (EPN on mma.sync 3xTF32)"; run --pair-tensor 1
echo "== --pair-const 1, model_weights (live hidden state: 4 of 5 steps walk the full far list)"; run --pair-const 1 --checkpoint model_weights
echo "== default, model_weights"; run --checkpoint model_weights
echo "== protein-like 40k atoms, default FP32 (row-group kernels)"; run --workload protein --atoms 40000 --steps 3
echo "== protein-like 40k atoms, --pair-const 1 (row-per-thread far kernel for the live steps)"; run --workload protein --atoms 40000 --steps 3 --pair-const 1
