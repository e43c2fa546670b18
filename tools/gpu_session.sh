#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_coop_variants.json
for v in coop0 coop1; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200_$v.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 --check >> gpurun_out/r2_coop_variants.json 2>> gpurun_out/r2_coop_variants.err
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 --len 50 --check >> gpurun_out/r2_coop_variants.json 2>> gpurun_out/r2_coop_variants.err
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 --check >> gpurun_out/r2_coop_variants.json 2>> gpurun_out/r2_coop_variants.err
done
timeout 1200 python -m pytest tests/test_gpu_profile.py tests/test_gpu_ragged.py tests/test_gpu_fused.py tests/test_gpu_golden.py -x -q > gpurun_out/r2_coop_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_coop_tests.log
