#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_profile.py -x -q > gpurun_out/r2_q_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_q_tests.log
for v in rev0 ""; do
  lib=para-suite_b200/lib/libparasuite_b200${v:+_$v}.so
  PARASUITE_B200_LIB=$PWD/$lib timeout 600 python tools/bench_kernels.py --iters 16 >> gpurun_out/r2_rev_variants.json 2>> gpurun_out/r2_rev_variants.err
  PARASUITE_B200_LIB=$PWD/$lib timeout 600 python tools/bench_kernels.py --iters 16 --len 50 >> gpurun_out/r2_rev_variants.json 2>> gpurun_out/r2_rev_variants.err
done
timeout 600 python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 > gpurun_out/r2_ragged_bench36.json 2>> gpurun_out/r2_rev_variants.err
