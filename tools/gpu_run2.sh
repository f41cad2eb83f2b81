set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -5 gpurun_out/pytest_gpu_all.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_q64_filt.json 2> gpurun_out/bench_q64_filt.err
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-regimes --flags 1024 > gpurun_out/bench_q64_x3.json 2>> gpurun_out/bench_q64_filt.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-regimes --queries 128 > gpurun_out/bench_q128_filt.json 2>> gpurun_out/bench_q64_filt.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-regimes --queries 1024 > gpurun_out/bench_q1024_filt.json 2>> gpurun_out/bench_q64_filt.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-regimes --queries 16 > gpurun_out/bench_q16_filt.json 2>> gpurun_out/bench_q64_filt.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-regimes --queries 32 > gpurun_out/bench_q32_filt.json 2>> gpurun_out/bench_q64_filt.err
for f in gpurun_out/bench_q*_filt.json gpurun_out/bench_q64_x3.json; do python - "$f" <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], d["config"]["queries_per_step"], "ms", round(d["ms_per_step"],3), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"], [ (r["queries_per_call"], round(r["ms_per_call"],3)) for r in d.get("regimes",[])])
PY
done
