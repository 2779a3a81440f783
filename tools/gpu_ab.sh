#!/bin/bash
# A/B of two builds of the library on the same box (development aid)
run() { python bench.py --molecules 300000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], {k: round(v,1) for k,v in d['phases_ms_per_step'].items()})"; }
echo "== default"; run
for lib in "$@"; do echo "== $lib"; EPNN_B200_LIB=$lib run; done
