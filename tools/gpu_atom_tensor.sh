#!/bin/bash
# Validation + A/B of the warp-level tensor per-atom kernel (option atom_tensor):  gpurun --timeout 900 -- 'bash tools/gpu_atom_tensor.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
bash tools/gpu_ab_opt.sh atom_tensor 0 1 2>&1 | tee gpurun_out/ab_atom_tensor.log
for v in 0 1; do echo "== atom_tensor=$v"; timeout 300 python tools/measure_noise_floor.py 400 atom_tensor=$v 2>&1 | grep -v "precision  0"; done | tee gpurun_out/noise_floor_atom_tensor.log
