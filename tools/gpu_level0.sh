cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filter.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_level0.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_level0.log
tail -8 gpurun_out/pytest_level0.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-regimes --no-configs > gpurun_out/bench_level0.json 2> gpurun_out/bench_level0.err
tail -3 gpurun_out/bench_level0.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_level0.json") if l.startswith("{")][-1])
print("headline", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]))
print(json.dumps(d.get("level0")))
print(json.dumps(d.get("reference_gpu_composition"))[:400])
PY
