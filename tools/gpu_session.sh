#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_pipe2_bench.json 2> gpurun_out/r2_pipe2_bench.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_pipe2_bench_n2.json 2> gpurun_out/r2_pipe2_bench_n2.err
