set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt 2>&1
nproc > gpurun_out/r2_n8_nproc.txt; free -g >> gpurun_out/r2_n8_nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29501 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_n8_bench.json 2> gpurun_out/r2_n8_bench.err
timeout 1500 $TR --master-port 29502 tools/run_config.py --configs 3,4,5 > gpurun_out/r2_n8_configs.json 2> gpurun_out/r2_n8_configs.err
