cd $GRAFT_REPO_ROOT
for Q in 1 2 4; do timeout 200 python bench.py --rows 10000000 --queries $Q --steps 400 --warmup 5 --no-cpu-baseline --no-regimes 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('stream Q$Q x400', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks'])
"; done
