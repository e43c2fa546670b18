#!/bin/bash
# scratch: one gpurun call
set -x
cd /root/repo
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 1800 $TR --master-port 29502 tools/run_config.py --configs 4 --reads4 250000000 > gpurun_out/r2_n${N}_config4.json 2> gpurun_out/r2_n${N}_config4.err
timeout 900 $TR --master-port 29501 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_n${N}_bench.json 2> gpurun_out/r2_n${N}_bench.err
