cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do for f in 268435456 0; do timeout 200 python bench.py --rows 10000000 --queries 4096 --steps 6 --warmup 2 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q4096 flags $f', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'scan ms/launch', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'fallbacks', d['exact_fallback_queries'])
    else: print(ln.rstrip()[:300])
"; done; done
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_filter.py -x -q -m gpu 2>&1 | tail -3
