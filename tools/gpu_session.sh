#!/bin/bash
# scratch: one gpurun call -- single-GPU bench record of the final code
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_final_ncu_list.log 2>&1
