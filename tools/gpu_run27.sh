cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for rep in 1 2; do for Q in 64 128; do for f in 0 536870912; do timeout 200 python bench.py --rows 10000000 --queries $Q --steps 150 --warmup 5 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q$Q flags $f', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'fallbacks', d['exact_fallback_queries'], d['e2e']['matches_device_path'])
    else: print(ln.rstrip()[:300])
"; done; done; done
