#!/bin/bash
# BASELINE configs[0]/[1] on the reference's own data (QM9 molecules / SSI dimers of data/mixed) + launch list of the Galectin-sized run
mkdir -p gpurun_out
timeout 200 python bench.py --workload qm9_test --steps 50 --warmup 5 2>gpurun_out/bq.err | grep '^{' > gpurun_out/bench_qm9_test.json; cut -c1-260 gpurun_out/bench_qm9_test.json
timeout 200 python bench.py --workload ssi --steps 50 --warmup 5 2>gpurun_out/bs.err | grep '^{' > gpurun_out/bench_ssi.json; cut -c1-260 gpurun_out/bench_ssi.json
CMD="python bench.py --workload protein --atoms 2220 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 120 $CMD > gpurun_out/plain_p2220.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_protein2220.csv $CMD > gpurun_out/ncu_p2220.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import json
for f in ["bench_qm9_test", "bench_ssi"]:
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"]), round(d["ms_per_step"], 3), "e2e", d["e2e"] and round(d["e2e"]["value"]), "cpu", d["cpu_baseline"], {k: round(v, 3) for k, v in d["phases_ms_per_step"].items()}, d["gpu_launches"], d["checks"])
    except Exception as ex:
        print(f, "FAILED", ex)
PY
tail -2 gpurun_out/bq.err
