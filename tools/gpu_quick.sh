cd $GRAFT_REPO_ROOT
for f in 0 65536 1073807360 1074331648; do timeout 200 python bench.py --rows 10000000 --queries 64 --steps 300 --warmup 5 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
names={0:'baseline',65536:'no_insert',1073807360:'no_insert+no_mma',1074331648:'no_insert+no_mma+no_query_tma (pure bank TMA stream)'}
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q64 x300', names[$f].ljust(52), 'ms', round(d['ms_per_step'],3), 'scan ms', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    else: print(ln.rstrip()[:300])
"; done
