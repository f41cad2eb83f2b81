cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_pools.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -25 gpurun_out/pytest_filter.log
