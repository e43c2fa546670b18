#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_swz_variants.json
for v in swz0 swz1 swz0 swz1; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200_$v.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 >> gpurun_out/r2_swz_variants.json 2>> gpurun_out/r2_swz_variants.err
done
for v in swz0 swz1; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200_$v.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 --len 50 >> gpurun_out/r2_swz_variants.json 2>> gpurun_out/r2_swz_variants.err
done
timeout 1200 python -m pytest tests/test_gpu_pileup.py tests/test_gpu_fused.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2_swz_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_swz_tests.log
