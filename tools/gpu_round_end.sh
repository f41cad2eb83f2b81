cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -4 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_r01c_g1.json 2> gpurun_out/bench_r01c_g1.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r01c_ref.json 2>> gpurun_out/bench_r01c_g1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r01c_g1.json"))
print("headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], d["clocks"], "fallbacks", d["exact_fallback_queries"], "launches", d["gpu_launches"])
for r in d["regimes"]: print(r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), round(r["useful_tflops"],1), round(r["frac_of_tf32_tensor_peak"],3), r["exact_fallback_queries"])
print(d.get("cpu_baseline"))
print(open("gpurun_out/bench_r01c_ref.json").read()[:600])
PY
python tools/bench_configs.py --cpu > gpurun_out/configs_r01c.jsonl 2> gpurun_out/configs_r01c.err
tail -3 gpurun_out/configs_r01c.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_filt_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_launch_filt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 2 -c 1 -o gpurun_out/filt_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_full_filt.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_tc_kernel<.int.256' -s 2 -c 1 -o gpurun_out/filt_pair_r01 -f python bench.py --queries 256 --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_full_pair.log 2>&1
tail -1 gpurun_out/ncu_full_pair.log
