#!/bin/bash
# Final validation of a build: GPU test suite, smoke(), the driver's default bench line and the reference arm.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh'
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
bash tools/gpu_bench_default.sh
( time python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_ref.json
