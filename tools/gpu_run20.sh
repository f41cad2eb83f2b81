cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -12 gpurun_out/pytest_gpu_all.log
for Q in 16 64; do
timeout 300 python bench.py --rows 10000000 --queries $Q --steps 20 --warmup 3 --no-cpu-baseline --no-regimes --bank-dtype bf16 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('bf16 Q $Q', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'scan ms/launch', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'])
    else: print(ln.rstrip()[:200])
"; done
timeout 300 python bench.py --rows 10000000 --queries 64 --steps 20 --warmup 3 --no-cpu-baseline --no-regimes --flags 1024 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('fp32 x3 Q 64', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']))
    else: print(ln.rstrip()[:200])
"
