cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/c1_probe.py 6345; DEVQA_FUSED_MERGE=1 python tools/c1_probe.py 6345
DEVQA_FUSED_MERGE=1 python tools/c1_probe.py 19035
DEVQA_FUSED_MERGE=1 timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -5 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
DEVQA_FUSED_MERGE=1 timeout 200 python bench.py --rows 10000000 --queries 1 --steps 100 --warmup 5 --no-cpu-baseline --no-regimes 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('10M Q1 x100', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'launches/step', d['gpu_launches']/d['steps'], d['e2e']['matches_device_path'])
"
