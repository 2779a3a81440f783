#!/bin/bash
# tcgen05 far kernel: parity of both implementations + A/B on a 40 k-atom system.   gpurun --timeout 600 -- 'bash tools/gpu_tc2.sh'
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_tensor_far.py -m gpu -x -q 2>&1 | tail -8
cat > /tmp/ab_tc.py <<'PY'
import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth
n = int(sys.argv[1])
w = load_weights("tests/golden/checkpoints/decay_model_weights")
offs, xyz, sp, Q = synth.protein_like(n, 9, seed=1)
res = {}
for tag, opts in (("far_const", {"gnn_far_tensor": 0}), ("tc1", {"gnn_far_tensor": 1, "gnn_far_tensor_impl": 1}), ("tc2", {"gnn_far_tensor": 1, "gnn_far_tensor_impl": 2})):
    eng = Engine(w, 0); eng.set_option("timing", 1)
    for k, v in opts.items(): eng.set_option(k, v)
    for it in range(3):
        q, q64 = eng.infer_batch(offs, xyz, sp, Q, None, want_f64=True)
    res[tag] = q64.copy()
    print(tag, "ms_gnn_pair %.2f  ms_total %.2f  sum-Q %.1e" % (eng.last_stats["ms_gnn_pair"], eng.last_stats["ms_total"], abs(q64.sum() - Q[0])), flush=True)
    eng.close()
print("tc2 vs tc1 max|dq| %.2e   tc1 vs far_const %.2e" % (np.abs(res["tc2"] - res["tc1"]).max(), np.abs(res["tc1"] - res["far_const"]).max()))
PY
timeout 90 python /tmp/ab_tc.py 40000
