#!/bin/bash
# ncu capture of the tcgen05 far kernels on a 20 k-atom system.   gpurun --timeout 600 -- 'bash tools/gpu_ncu_tc2.sh'
mkdir -p gpurun_out
cat > /tmp/run_tc.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth
w = load_weights("tests/golden/checkpoints/model2_weights")
offs, xyz, sp, Q = synth.protein_like(20000, 9, seed=1)
eng = Engine(w, 0); eng.set_option("gnn_far_tensor", 1); eng.set_option("gnn_far_tensor_impl", int(sys.argv[1]))
for it in range(2): eng.infer_batch(offs, xyz, sp, Q, None)
PY
timeout 120 python /tmp/run_tc.py 2 && timeout 400 ncu --set full --clock-control none --import-source on -k regex:gnn_far_tc -s 4 -c 1 -f -o gpurun_out/prof_tc2 python /tmp/run_tc.py 2 > gpurun_out/ncu_tc2.log 2>&1
echo "ncu rc=$?"
