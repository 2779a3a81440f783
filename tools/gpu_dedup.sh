#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for ck in decay_model_weights model2_weights; do for dd in 1 0; do
  echo -n "$ck dedup=$dd  "
  python bench.py --molecules 300000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --checkpoint $ck --dedup-far $dd 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,1) for k,v in d['phases_ms_per_step'].items()})"
done; done
