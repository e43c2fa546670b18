#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_rg_variants.json
for v in rg3 rg4 rg3 rg4; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200_$v.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 --check >> gpurun_out/r2_rg_variants.json 2>> gpurun_out/r2_rg_variants.err
done
for v in rg3 rg4; do
  lib=$PWD/para-suite_b200/lib/libparasuite_b200_$v.so
  PARASUITE_B200_LIB=$lib timeout 600 python tools/bench_kernels.py --reads 4000000 --len 60 --max-len 64 --trim 30 --check >> gpurun_out/r2_rg_variants.json 2>> gpurun_out/r2_rg_variants.err
done
