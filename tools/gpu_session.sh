set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/para-suite_b200/lib
timeout 1200 python -m pytest tests/test_gpu_fused.py tests/test_gpu_profile.py tests/test_gpu_pileup.py tests/test_gpu_api.py -x -q > gpurun_out/r2_gpu3_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_gpu3_tests.log
timeout 300 python tools/bench_kernels.py --check > gpurun_out/r2_gpu3_k_default.json 2> gpurun_out/r2_gpu3_k_default.err
PARASUITE_B200_LIB=$L/libparasuite_b200_b5.so timeout 300 python tools/bench_kernels.py --check > gpurun_out/r2_gpu3_k_b5.json 2> gpurun_out/r2_gpu3_k_b5.err
timeout 300 python tools/bench_kernels.py --len 50 --check > gpurun_out/r2_gpu3_k_L50.json 2> gpurun_out/r2_gpu3_k_L50.err
timeout 300 python bench.py --small --steps 3 --warmup 1 > gpurun_out/r2_gpu3_bench_small.json 2> gpurun_out/r2_gpu3_bench_small.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_gpu3_bench.json 2> gpurun_out/r2_gpu3_bench.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_gpu3_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_fast_kernel' -s 3 -c 1 -o gpurun_out/r2_gpu3_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2_gpu3_ncu.log 2>&1
ls -la gpurun_out | tail -12
