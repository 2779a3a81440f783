#!/bin/bash
# final evidence run for the round: full GPU suite, smoke, default bench + reference arm, protein benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
for TC in 0 1; do
python bench.py --workload protein --atoms 2220 --steps 20 --warmup 5 --gnn-far-tensor $TC 2>/dev/null | grep '^{' > gpurun_out/bench_protein2220_tc$TC.json
python bench.py --workload protein --atoms 40000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --gnn-far-tensor $TC 2>/dev/null | grep '^{' > gpurun_out/bench_protein40000_tc$TC.json
done
python - <<'PY'
import json
for f in ["bench", "bench_protein2220_tc0", "bench_protein2220_tc1", "bench_protein40000_tc0", "bench_protein40000_tc1"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as ex:
        print(f, "FAILED", ex); continue
    print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"], 3),
          d["roofline"].get("tensor_far", {}).get("frac"), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]),
          d["roofline"].get("far_dedup", {}).get("row_steps_collapsed"))
PY
