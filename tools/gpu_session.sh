set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2_gpu5_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2_gpu5_tests.log
