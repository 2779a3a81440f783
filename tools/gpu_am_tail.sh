#!/bin/bash
# Error tail (all of data/mixed, model2_weights) + speed of build variants of the tensor per-atom kernel:
#   gpurun --timeout 900 -- 'bash tools/gpu_am_tail.sh am_lolo am_rlo'
mkdir -p gpurun_out
for v in default "$@"; do
  if [ $v = default ]; then unset EPNN_B200_LIB; else export EPNN_B200_LIB=build/variants/libepnn_$v.so; fi
  echo "=== variant $v"
  timeout 300 python tools/measure_fp32_tail.py model2_weights 2>&1 | grep "fp32"
  CKPT=model2_weights bash tools/gpu_ab_opt.sh atom_tensor 1 2>&1 | grep atom_tensor
done 2>&1 | tee gpurun_out/am_tail.log
