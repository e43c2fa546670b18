#!/bin/bash
# scratch: one gpurun call -- last check of the final library
set -x
cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1
