set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29501 bench.py --gpus 2 --small --steps 3 --warmup 1 > gpurun_out/r2_n2_bench_small.json 2> gpurun_out/r2_n2_bench_small.err
timeout 600 $TR --master-port 29502 tools/run_config.py --configs 3,4,5 --small-ref --reads3 600000 --reads4 700000 --reads5 200000 --prefix 300000 --steps 2 > gpurun_out/r2_n2_cfg_small.json 2> gpurun_out/r2_n2_cfg_small.err
timeout 900 $TR --master-port 29503 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_n2_bench.json 2> gpurun_out/r2_n2_bench.err
