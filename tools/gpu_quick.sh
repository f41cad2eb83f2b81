cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --rows $3 --queries $4 --steps $5 --warmup 5 --no-cpu-baseline --no-regimes --flags $2 $6 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$1'.ljust(34), 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
    else: print(ln.rstrip()[:300])
"; }
for Q in 16 128 256; do
run "10M Q$Q x150 prefetch" 0 10000000 $Q 150
run "10M Q$Q x150 no prefetch" 1048576 10000000 $Q 150
done
run "10M Q1024 x12 prefetch" 0 10000000 1024 12
run "10M Q1024 x12 no prefetch" 1048576 10000000 1024 12
run "1M Q64 x300 prefetch" 0 1000000 64 300
run "1M Q64 x300 no prefetch" 1048576 1000000 64 300
run "10M bf16 Q64 x150 prefetch" 0 10000000 64 150 "--bank-dtype bf16"
run "10M bf16 Q64 x150 no prefetch" 1048576 10000000 64 150 "--bank-dtype bf16"
run "10M Q64 x20 prefetch" 0 10000000 64 20
run "10M Q64 x20 no prefetch" 1048576 10000000 64 20
