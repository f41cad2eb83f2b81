cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -5 gpurun_out/pytest_filter.log
for Q in 64 128; do
timeout 300 python bench.py --rows 10000000 --queries $Q --steps 20 --warmup 3 --no-cpu-baseline --no-regimes 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q $Q', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'scan ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    else: print(ln.rstrip()[:200])
"; done 2>&1 | tee gpurun_out/filt4_sweep.txt
DEVQA_TC_TRACE=1 python de-vqa_b200/build.py --force > /dev/null 2>&1
echo "== Q=128"; python tools/tc_trace.py 10000000 128 0 | tail -5
echo "== Q=64"; python tools/tc_trace.py 10000000 64 0 | tail -5
