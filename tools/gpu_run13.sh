cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -4 gpurun_out/pytest_gpu_all.log
python tools/bench_configs.py --only C2,C3 > gpurun_out/configs_r01c.jsonl 2> gpurun_out/configs_r01c.err
python - <<'PY'
import json
for ln in open("gpurun_out/configs_r01c.jsonl"):
    d=json.loads(ln)
    print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items()})
PY
