cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_emit.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_emit.log
tail -15 gpurun_out/pytest_emit.log
python tools/rows_sweep.py
python tools/rows_sweep.py --no-shadow
