cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py -q -m gpu -x --timeout=600 -k "shadow" > gpurun_out/pytest_shadow.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_shadow.log
tail -25 gpurun_out/pytest_shadow.log
for sh in 1 0; do
python bench.py --steps 100 --warmup 5 --no-configs --no-gpu-composition --no-cpu-baseline $( [ $sh = 1 ] && echo --shadow ) | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('shadow $sh', 'ms', round(d['ms_per_step'],4), 'q/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'scan ms', round(d['roofline']['launch_ms'],4), d['clocks']['sm_mhz'], d['e2e']['matches_device_path'], 'fb', d['exact_fallback_queries'], 'viol', d['bound_violations'], d['result_digest'][:16])
        for r in d.get('regimes', []): print('   ', r['queries_per_call'], round(r['ms_per_call'],3), round(r['queries_per_s']), round(r['frac_of_measured_hbm'],3), round(r['useful_tflops'],1), r['exact_fallback_queries'], r['bound_violations'])
"
done
