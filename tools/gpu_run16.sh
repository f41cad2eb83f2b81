cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -x -q -m gpu > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log
tail -4 gpurun_out/pytest_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r01b_g2.json 2> gpurun_out/bench_r01b_g2.err
tail -2 gpurun_out/bench_r01b_g2.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r01b_g2.json") if l.startswith("{")][-1])
print("2gpu headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"], "fallbacks", d["exact_fallback_queries"])
for r in d["regimes"]: print(r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), r["exact_fallback_queries"])
PY
