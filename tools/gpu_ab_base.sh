#!/bin/bash
# A/B of the working-tree library against build/variants/libepnn_base.so (a copy of the previous build), small-system
# bench at 300 k molecules, after the parity tests:   gpurun --timeout 900 -- 'bash tools/gpu_ab_base.sh'
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pair_const.py -m gpu -q -x > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ab.log
bash tools/gpu_ab_variants.sh "" default base default base 2>&1 | tee gpurun_out/ab_base.log
bash tools/gpu_ab_variants.sh "--checkpoint model2_weights" default base 2>&1 | tee -a gpurun_out/ab_base.log
