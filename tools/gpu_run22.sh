cd $GRAFT_REPO_ROOT
for f in 0 524288 65536; do timeout 200 python bench.py --rows 10000000 --queries 1024 --steps 5 --warmup 2 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q1024 flags $f', 'ms', round(d['ms_per_step'],3), 'scan ms/launch', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    else: print(ln.rstrip()[:200])
"; done
