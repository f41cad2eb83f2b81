# performance bisection of the tcgen05 scan with the undocumented debug bits (results are wrong when set)
Q=${1:-64}
for f in 0 65536 131072 262144 524288 1048576 983040; do timeout 200 python bench.py --rows 4000000 --queries $Q --steps 5 --warmup 3 --no-cpu-baseline --flags $f 2>&1 | python -c "
import sys,json
names={0:'baseline',65536:'no_insert',131072:'no_split',262144:'no_mma2',524288:'no_query_tma',1048576:'no_l2_prefetch',983040:'none_of_them'}
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q $Q', names[$f].ljust(16), 'ms', round(d['ms_per_step'],3), 'cycles/kblock', round(d['ms_per_step']*1e-3*1.965e9/(4000000/128/148*24)))
    else: print(ln.rstrip()[:200])
"; done
