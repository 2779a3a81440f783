#!/bin/bash
# Large-system far-column de-duplication: GPU suite, then protein-like systems with dedup_far on / off (1 GPU).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
B="timeout 300 python bench.py --workload protein --no-cpu-baseline"
$B --atoms 2220 --steps 20 --warmup 5 2>gpurun_out/bp2220.err | grep '^{' > gpurun_out/bench_protein2220_dedup1.json
$B --atoms 2220 --steps 20 --warmup 5 --dedup-far 0 2>/dev/null | grep '^{' > gpurun_out/bench_protein2220_dedup0.json
$B --atoms 40000 --steps 5 --warmup 3 2>gpurun_out/bp40k.err | grep '^{' > gpurun_out/bench_protein40000_dedup1.json
$B --atoms 100000 --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/bench_protein100000_dedup1.json
$B --atoms 1000000 --steps 5 --warmup 3 2>gpurun_out/bp1M.err | grep '^{' > gpurun_out/bench_protein1000000_dedup1.json
$B --atoms 40000 --steps 5 --warmup 3 --checkpoint model2_weights --no-e2e 2>/dev/null | grep '^{' > gpurun_out/bench_protein40000_model2.json
python - <<'PY'
import json
for f in ["bench_protein2220_dedup1", "bench_protein2220_dedup0", "bench_protein40000_dedup1", "bench_protein100000_dedup1",
          "bench_protein1000000_dedup1", "bench_protein40000_model2"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as ex:
        print(f, "FAILED", ex); continue
    print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]),
          {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["roofline"].get("far_dedup", {}).get("row_steps_collapsed"),
          d["checks"])
PY
tail -3 gpurun_out/bp1M.err
