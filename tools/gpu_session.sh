#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ragged.py tests/test_gpu_bam.py tests/test_gpu_profile.py tests/test_config1.py -m gpu -x -q > gpurun_out/r2_ragged_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_ragged_tests.log
timeout 600 python tools/bench_kernels.py --reads 4000000 --len 44 --trim 18 --check > gpurun_out/r2_ragged_bench.json 2> gpurun_out/r2_ragged_bench.err
timeout 600 python tools/bench_kernels.py --reads 10000000 --len 36 --trim 20 --check > gpurun_out/r2_ragged_bench36.json 2>> gpurun_out/r2_ragged_bench.err
