#!/bin/bash
# The driver's multi-GPU invocation:  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_bench_multi.sh N'
N=$1; mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err ) 2>&1 | grep real
echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err | cut -c1-300
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/bench_n{n}.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3))
for k, v in d.get("secondary", {}).items():
    print(k, round(v["value"]), round(v["ms_per_step"], 3), {kk: vv for kk, vv in v.items() if kk in ("bit_identical_to_single_gpu", "exchange", "phases_ms_per_step", "precision_used")})
PY
