#!/bin/bash
# Validation + A/B of the two-stream chunk pipeline (option chunk_streams):  gpurun --timeout 1200 -- 'bash tools/gpu_chunk_streams.sh'
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for ck in decay_model_weights model2_weights; do NMOL=1000000 CKPT=$ck bash tools/gpu_ab_opt.sh chunk_streams 1 2; done 2>&1 | tee gpurun_out/ab_chunk_streams.log
bash tools/gpu_bench_default.sh
