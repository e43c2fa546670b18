#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
PARASUITE_B200_BATCHER_TIMING=1 timeout 1500 python tools/bench_extra.py --bam-repeat 40 > gpurun_out/r2_bench_extra40.json 2> gpurun_out/r2_bench_extra40.err
