#!/bin/bash
# scratch: one gpurun call -- bench record of the final code at N = 8
set -x
cd /root/repo
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 2950$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_final_n${N}_bench.json 2> gpurun_out/r2_final_n${N}_bench.err
