#!/bin/bash
# A/B: V1 GNN bundle k-loop unroll 1, V2 per-atom kernel unroll 1, V3 EPN direction loop unrolled, V4 per-atom kernel unroll 4
mkdir -p gpurun_out
run() { timeout 150 python bench.py --molecules 200000 --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})"; }
echo "== default"; run
for v in V1 V2 V3 V4; do echo "== $v"; EPNN_B200_LIB=$PWD/build/libepnn_$v.so run; done
