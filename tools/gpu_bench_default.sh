#!/bin/bash
# The driver's own invocations: default bench line (N = 1) + reference arm.   gpurun --timeout 900 -- 'bash tools/gpu_bench_default.sh'
mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench.json"))
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["value"]), "roofline", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 3),
      "alg", round(d["roofline"]["algorithmic"]["frac"], 3), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]))
print({k: round(v, 2) for k, v in d["phases_ms_per_step"].items()}, d["checks"])
for k, v in d.get("secondary", {}).items():
    print(k, round(v["value"]), round(v["ms_per_step"], 3), {kk: vv for kk, vv in v.items() if kk in ("bit_identical_to_single_gpu", "precision_used", "max_abs_dq_vs_reference_preds_npy", "probe_max_abs_dq_fp32_vs_fp64_kernels", "e2e")})
PY
