set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -30 gpurun_out/pytest_filter.log
