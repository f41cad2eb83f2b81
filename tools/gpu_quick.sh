cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_filter.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -6 gpurun_out/pytest_filter.log
for Q in 300 1100; do timeout 200 python bench.py --rows 10000000 --queries $Q --steps 5 --warmup 2 --no-cpu-baseline --no-regimes 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q$Q', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'launches/step', d['gpu_launches']/d['steps'], 'fallbacks', d['exact_fallback_queries'], d['e2e']['matches_device_path'])
    else: print(ln.rstrip()[:300])
"; done
