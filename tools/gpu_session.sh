#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_rb_bench.json 2> gpurun_out/r2_rb_bench.err
timeout 900 python -m pytest tests/test_gpu_profile.py tests/test_gpu_api.py tests/test_gpu_fused.py tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/r2_rb_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_rb_tests.log
