#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_bam.py tests/test_gpu_stream.py -m gpu -x -q > gpurun_out/r2_compact_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_compact_tests.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
