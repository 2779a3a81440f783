#!/bin/bash
# A/B of tcgen05 far-kernel build variants on a 40 k-atom system + their parity tests:  gpurun -- 'bash tools/gpu_tc_variants.sh default coal ...'
mkdir -p gpurun_out
cat > /tmp/ab_tcv.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth
w = load_weights("tests/golden/checkpoints/model2_weights")
offs, xyz, sp, Q = synth.protein_like(40000, 9, seed=1)
eng = Engine(w, 0); eng.set_option("timing", 1); eng.set_option("gnn_far_tensor", 1); eng.set_option("keep_hidden", 1)
for it in range(3):
    q, q64 = eng.infer_batch(offs, xyz, sp, Q, None, want_f64=True)
h = eng.hidden(40000)
print("ms_gnn_pair %.2f  ms_total %.2f  checksum q %.10e  h %.10e" % (eng.last_stats["ms_gnn_pair"], eng.last_stats["ms_total"], float(np.abs(q64).sum()), float(np.abs(h.astype(np.float64)).sum())), flush=True)
PY
for V in "$@"; do
  if [ "$V" = default ]; then unset EPNN_B200_LIB; else export EPNN_B200_LIB=$PWD/build/variants/libepnn_$V.so; fi
  echo "== $V"; timeout 120 python /tmp/ab_tcv.py; timeout 120 python -m pytest tests/test_gpu_tensor_far.py -m gpu -x -q -k "1" 2>&1 | tail -1
done
