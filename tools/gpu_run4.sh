cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for Q in 128 64; do
for st in 0 2 4 8 15; do f=$((st << 24)); timeout 200 python bench.py --rows 10000000 --queries $Q --steps 20 --warmup 3 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q $Q seed_tiles $st', 'ms', round(d['ms_per_step'],3), 'scan ms', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    else: print(ln.rstrip()[:200])
"; done; done 2>&1 | tee gpurun_out/seed_sweep.txt
