#!/bin/bash
# A/B of build variants of the tensor per-atom kernel (tools/build_variant.sh am_X -D...):  gpurun --timeout 900 -- 'bash tools/gpu_am_variants.sh am_b am_c'
mkdir -p gpurun_out
for v in default "$@"; do
  if [ $v = default ]; then unset EPNN_B200_LIB; else export EPNN_B200_LIB=build/variants/libepnn_$v.so; fi
  echo "=== variant $v"
  for ck in decay_model_weights model2_weights; do CKPT=$ck bash tools/gpu_ab_opt.sh atom_tensor 1; done 2>&1 | grep atom_tensor
  timeout 300 python tools/measure_noise_floor.py 400 atom_tensor=1 2>&1 | grep "precision 32" | grep -v decay
done 2>&1 | tee gpurun_out/am_variants.log
