#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_all_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_all_tests.log
timeout 1500 python tools/bench_extra.py --bam-reads 1000000 --bam-repeat 1 > gpurun_out/r2_bench_extra_tools.json 2> gpurun_out/r2_bench_extra_tools.err
PARASUITE_B200_WINDOW_READS=250000 PARASUITE_B200_BATCH_READS=250000 timeout 1500 python tools/bench_extra.py --bam-reads 1000000 --bam-repeat 1 > gpurun_out/r2_bench_extra_tools_w.json 2> gpurun_out/r2_bench_extra_tools_w.err
