cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fullsize.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_shadow.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_shadow.log
tail -12 gpurun_out/pytest_shadow.log
timeout 600 python tools/soak.py 600 2000000 2>&1 | tail -7
python tools/rows_sweep.py
