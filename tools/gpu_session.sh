#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_all_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_all_tests.log
rm -f gpurun_out/r2_len_variants.json
for L in 40 51 32; do
  timeout 600 python tools/bench_kernels.py --iters 12 --len $L --max-len 64 --check >> gpurun_out/r2_len_variants.json 2>> gpurun_out/r2_len_variants.err
done
