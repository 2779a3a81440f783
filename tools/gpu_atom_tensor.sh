#!/bin/bash
# Validation, A/B and ncu capture of the warp-level tensor per-atom kernel (option atom_tensor):
#   gpurun --timeout 1200 -- 'bash tools/gpu_atom_tensor.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for ck in decay_model_weights model2_weights; do CKPT=$ck bash tools/gpu_ab_opt.sh atom_tensor 0 1; done 2>&1 | tee gpurun_out/ab_atom_tensor.log
if [ -f build/variants/libepnn_nw6.so ]; then echo "== 6 warps per CTA"; CKPT=model2_weights EPNN_B200_LIB=build/variants/libepnn_nw6.so bash tools/gpu_ab_opt.sh atom_tensor 1 2>&1 | tee -a gpurun_out/ab_atom_tensor.log; fi
timeout 300 python tools/measure_noise_floor.py 400 atom_tensor=1 2>&1 | grep -v "precision  0" | tee gpurun_out/noise_floor_atom_tensor.log
cat > /tmp/run_am.py <<'PY'
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
from epnn_b200.checkpoint import load_weights
from epnn_b200.engine import Engine
from epnn_b200 import synth
w = load_weights("tests/golden/checkpoints/model2_weights")
offs, xyz, sp, Q = synth.qm9_shaped(50000, w.n_x, seed=0)
npad = np.full(50000, 29, np.int32)
eng = Engine(w, 0)
for it in range(3): eng.infer_batch(offs, xyz, sp, Q, npad)
PY
timeout 120 python /tmp/run_am.py && timeout 500 ncu --set full --clock-control none --import-source on -k regex:atom_mma -s 10 -c 5 -f -o gpurun_out/prof_atom_mma python /tmp/run_am.py > gpurun_out/ncu_atom_mma.log 2>&1
echo "ncu rc=$?"
