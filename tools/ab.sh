# A/B of a debug flag on the bench workload: usage tools/ab.sh <flagbits> <rows> <queries...>
F=$1; ROWS=$2; shift 2
for q in "$@"; do for rep in 1 2; do for f in 0 $F; do timeout 200 python bench.py --rows $ROWS --queries $q --steps 20 --warmup 3 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('rows $ROWS Q $q flags $f ms', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'])
"; done; done; done
