cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tc.py tests/test_gpu_fullsize.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -4 gpurun_out/pytest_filter.log
python tools/bench_configs.py --only C2 2> gpurun_out/configs_r01c.err | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items()})
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python tools/c2_probe.py > gpurun_out/c2_probe.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_c2.csv
