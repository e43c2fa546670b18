#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_all_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_all_tests.log
timeout 600 python tools/bench_kernels.py --reads 4000000 --len 44 --trim 18 --check > gpurun_out/r2_ragged_bench.json 2> gpurun_out/r2_ragged_bench.err
timeout 600 python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 --check > gpurun_out/r2_ragged_bench36.json 2>> gpurun_out/r2_ragged_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_ragged_launches.csv python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 --iters 4 > gpurun_out/r2_ragged_ncu.log 2>&1
