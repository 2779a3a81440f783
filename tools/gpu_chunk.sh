#!/bin/bash
for ca in 524288 1048576 2097152 4194304 8388608; do
  echo -n "chunk_atoms=$ca  "
  python bench.py --molecules 1000000 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --chunk-atoms $ca 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), {k: round(v,1) for k,v in d['phases_ms_per_step'].items()})"
done
