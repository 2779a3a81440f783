#!/bin/bash
# A/B of the per-atom kernel variants (build/libepnn_{B,C,D}.so: CSR walk 4 entries in flight; + L2 prefetch of the next tile; + 14 warps)
mkdir -p gpurun_out
run() { timeout 150 python bench.py --molecules 300000 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,2) for k,v in d['phases_ms_per_step'].items()})"; }
echo "== default"; run
for v in B C D; do echo "== $v"; EPNN_B200_LIB=$PWD/build/libepnn_$v.so run; done
echo "== default again"; run
EPNN_B200_LIB=$PWD/build/libepnn_C.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_871 or charges_vs_oracle or protein_golden or chunking or determinism" 2>&1 | tail -2
