#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_b5_variants.json
for v in b5 "" b5 ""; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200${v:+_$v}.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 --check >> gpurun_out/r2_b5_variants.json 2>> gpurun_out/r2_b5_variants.err
done
for v in b5 ""; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200${v:+_$v}.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --iters 16 --len 50 --check >> gpurun_out/r2_b5_variants.json 2>> gpurun_out/r2_b5_variants.err
done
