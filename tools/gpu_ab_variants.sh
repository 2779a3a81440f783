#!/bin/bash
# A/B of library variants (tools/build_variant.sh):  gpurun -- 'bash tools/gpu_ab_variants.sh "<bench flags>" default nw8x2 ...'
mkdir -p gpurun_out
FLAGS=$1; shift
run() { timeout 150 python bench.py --molecules 300000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --secondary 0 $FLAGS 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()}, d['checks'])"; }
for V in "$@"; do
  if [ "$V" = default ]; then unset EPNN_B200_LIB; else export EPNN_B200_LIB=$PWD/build/variants/libepnn_$V.so; fi
  echo "== $V $FLAGS"; run
done
