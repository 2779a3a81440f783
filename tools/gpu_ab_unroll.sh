#!/bin/bash
# A/B of the k-loop unroll factor of tile_gemm_f32x2 (build/libepnn_U{4,8}.so vs the default 2); arithmetic order unchanged
mkdir -p gpurun_out
run() { timeout 150 python bench.py --molecules 300000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})"; }
echo "== default"; run
for v in U4 U8; do echo "== $v"; EPNN_B200_LIB=$PWD/build/libepnn_$v.so run; done
