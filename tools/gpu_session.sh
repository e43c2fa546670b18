#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/r2_gen_full.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_generic_kernel' -s 2 -c 1 -o gpurun_out/r2_gen_full python tools/bench_kernels.py --reads 4000000 --len 150 --mode 1 --max-len 176 --iters 4 > gpurun_out/r2_gen_ncu.log 2>&1
