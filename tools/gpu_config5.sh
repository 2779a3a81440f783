#!/bin/bash
# BASELINE config 5 on N GPUs: one protein-like system of A atoms sharded by rows.   gpurun --gpus N -- 'bash tools/gpu_config5.sh N A [steps]'
N=$1; A=$2; K=${3:-3}; mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload protein --atoms $A --steps $K --warmup 1 --no-cpu-baseline --no-e2e --secondary 0 > gpurun_out/bench_protein${A}_n$N.json 2> gpurun_out/bench_protein${A}_n$N.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_protein${A}_n$N.err | cut -c1-300
python - $N $A <<'PY'
import json, sys
n, a = sys.argv[1], sys.argv[2]
d = json.loads([l for l in open(f"gpurun_out/bench_protein{a}_n{n}.json") if l.startswith("{")][-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "roofline", d["roofline"]["kernel"][:40], round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3))
print({k: round(v, 2) for k, v in d["phases_ms_per_step"].items()}, d["checks"]["max_abs_sum_q_minus_Q"], d["roofline"].get("far_dedup"))
PY
