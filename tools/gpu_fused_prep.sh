#!/bin/bash
# Validation + A/B of the fused list building (option fused_prep):  gpurun --timeout 1200 -- 'bash tools/gpu_fused_prep.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for ck in decay_model_weights model2_weights; do CKPT=$ck bash tools/gpu_ab_opt.sh fused_prep 0 1; done 2>&1 | tee gpurun_out/ab_fused_prep.log
bash tools/gpu_launch_list.sh
