# new tests + the full N=1 bench line + the reference arm, as the driver runs them
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pools.py -q -m gpu -x --timeout=600 -k "level0 or select or scope or hice or liveedit" > gpurun_out/pytest_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_g1.json 2> gpurun_out/bench_r02_g1.err ) 2>&1 | tail -3
tail -5 gpurun_out/bench_r02_g1.err
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_ref.json 2>> gpurun_out/bench_r02_g1.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_g1.json") if l.startswith("{")][-1])
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["clocks"], "fallbacks", d["exact_fallback_queries"], "viol", d["bound_violations"], "digest", d["result_digest"][:16], "launches", d["gpu_launches"])
for r in d.get("regimes", []): print(r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), round(r["useful_tflops"],1), r["exact_fallback_queries"], r["bound_violations"])
print(json.dumps(d.get("reference_gpu_composition"), indent=0))
print(json.dumps(d.get("configs"), indent=0))
print(d.get("cpu_baseline"))
print(open("gpurun_out/bench_r02_ref.json").read()[:1500])
PY
free -g | head -2; nproc
