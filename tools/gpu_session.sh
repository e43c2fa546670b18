set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
L=$PWD/para-suite_b200/lib
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_profile.py -x -q > gpurun_out/r2_gpu2_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_gpu2_tests.log
timeout 300 python tools/bench_kernels.py --check > gpurun_out/r2_gpu2_k_default.json 2> gpurun_out/r2_gpu2_k_default.err
for v in b4 c256 c64; do
PARASUITE_B200_LIB=$L/libparasuite_b200_$v.so timeout 300 python tools/bench_kernels.py --check > gpurun_out/r2_gpu2_k_$v.json 2> gpurun_out/r2_gpu2_k_$v.err
done
PARASUITE_B200_LIB=$L/libparasuite_b200_b4.so timeout 300 python tools/bench_kernels.py --len 50 --check > gpurun_out/r2_gpu2_k_b4_L50.json 2> gpurun_out/r2_gpu2_k_b4_L50.err
export PARASUITE_B200_LIB=$L/libparasuite_b200_b4.so
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_gpu2_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'profile_fast_kernel|pl_cluster_kernel' -s 6 -c 2 -o gpurun_out/r2_gpu2_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_gpu2_ncu.log 2>&1
ls -la gpurun_out | tail -12
