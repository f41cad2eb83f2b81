cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for f in 268435456 0; do timeout 120 python bench.py --rows 10000000 --queries 1024 --steps 5 --warmup 2 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q1024 flags $f', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'scan ms/launch', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'fallbacks', d['exact_fallback_queries'], 'e2e==dev', d['e2e']['matches_device_path'])
    else: print(ln.rstrip()[:300])
"; echo "exit $?"; done
DEVQA_TC_TRACE=1 python de-vqa_b200/build.py --force > /dev/null 2>&1
echo "== Q=256 pair"; timeout 120 python tools/tc_trace.py 10000000 256 0 | sed -n '1,6p' | cut -c1-250;  timeout 120 python tools/tc_trace.py 10000000 256 0 | tail -5
