cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -5 gpurun_out/pytest_filter.log
for Q in 64 128 256 1024; do
steps=20; if [ $Q -ge 1024 ]; then steps=5; fi
timeout 300 python bench.py --rows 10000000 --queries $Q --steps $steps --warmup 3 --no-cpu-baseline --no-regimes 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q $Q', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'scan ms/launch', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'], d['clocks']['reasons'], 'fallbacks', d['exact_fallback_queries'])
    else: print(ln.rstrip()[:200])
"; done 2>&1 | tee gpurun_out/filt6_sweep.txt
python tools/bench_configs.py --only C1,C2,C3 > gpurun_out/configs_r01b.jsonl 2> gpurun_out/configs_r01b.err
python - <<'PY'
import json
for ln in open("gpurun_out/configs_r01b.jsonl"):
    d=json.loads(ln)
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items()})
PY
