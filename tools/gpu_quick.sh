cd $GRAFT_REPO_ROOT
for rep in 1 2; do for f in 2147483648 0; do timeout 200 python bench.py --rows 10000000 --queries 64 --steps 300 --warmup 5 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
names={0:'wide prefetch',2147483648:'box prefetch'}
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q64 x300', names[$f].ljust(14), 'ms', round(d['ms_per_step'],3), 'scan ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['e2e']['matches_device_path'])
    else: print(ln.rstrip()[:300])
"; done; done
for f in 2147483648 0; do timeout 200 python bench.py --rows 10000000 --queries 256 --steps 40 --warmup 5 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
names={0:'wide prefetch',2147483648:'box prefetch'}
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q256 x40', names[$f].ljust(14), 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']))
"; done
