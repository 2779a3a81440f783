#!/bin/bash
# Build a variant of the library with extra -D flags for A/B measurements:  tools/build_variant.sh NAME [-DRUN_NW=8 ...]
# -> build/variants/libepnn_NAME.so, used through EPNN_B200_LIB=build/variants/libepnn_NAME.so (development only).
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
OUT=build/variants/$NAME; mkdir -p $OUT
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off --cudart static"
for s in epnn_b200/csrc/*.cu epnn_b200/csrc/*.cpp; do
  o=$OUT/$(basename ${s%.*}).o
  nvcc $FLAGS "$@" -c $s -o $o &
done
wait
nvcc -shared --cudart static -gencode arch=compute_100a,code=sm_100a -o build/variants/libepnn_$NAME.so $OUT/*.o
echo build/variants/libepnn_$NAME.so
