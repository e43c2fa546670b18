#!/bin/bash
# scratch: one gpurun call -- final single-GPU records of the round
set -x
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/r2_final_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_final_tests.log
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_final_ncu_list.log 2>&1
