# all GPU tests + smoke on one B200 (results under gpurun_out/)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/filter_error_bound.json
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=900 > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -30 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-regimes | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('bench Q64 x30', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'launches/step', d['gpu_launches']/d['steps'], d['e2e']['matches_device_path'])
"
