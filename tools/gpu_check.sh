cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_tc.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_check.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_check.log
tail -6 gpurun_out/pytest_check.log
python tools/rows_sweep.py
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-regimes --no-configs --no-gpu-composition | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('bench Q64 x30', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), d['e2e']['matches_device_path'])
"
