# round-2 evidence run on one B200: all GPU tests, smoke, the bench line as the driver runs it, the reference arm,
# ncu launch list + full capture of the headline kernel (each ncu pass only after the plain command exited 0)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/filter_error_bound.json
timeout 2400 python -m pytest tests -q -m gpu -x --timeout=900 > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -4 gpurun_out/pytest_gpu_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_g1.json 2> gpurun_out/bench_r02_g1.err ) 2>&1 | grep real
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02_ref.json 2>> gpurun_out/bench_r02_g1.err ) 2>&1 | grep real
python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-configs --no-gpu-composition > gpurun_out/bench_r02_g1_200steps.json 2>> gpurun_out/bench_r02_g1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_r02_g1.json", "gpurun_out/bench_r02_g1_200steps.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, "ms", round(d["ms_per_step"],4), "q/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["clocks"], "fb", d["exact_fallback_queries"], "viol", d["bound_violations"], d["result_digest"][:16], "no-shadow", d["without_shadow"])
    for r in d.get("regimes", []): print("   ", r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), round(r["useful_tflops"],1), round(r["frac_of_tf32_tensor_peak"],3), r["exact_fallback_queries"], r["bound_violations"])
    if "reference_gpu_composition" in d: print(json.dumps(d["reference_gpu_composition"])[:600])
    if "configs" in d: print(json.dumps(d["configs"])[:1500])
    print(d.get("cpu_baseline"))
print(open("gpurun_out/bench_r02_ref.json").read()[:700])
PY
tail -3 gpurun_out/bench_r02_g1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02_headline.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-regimes --no-configs --no-gpu-composition > gpurun_out/ncu_launch_r02.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 4 -c 1 -o gpurun_out/r02_shadow_scan -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes --no-configs --no-gpu-composition > gpurun_out/ncu_full_r02.log 2>&1
tail -2 gpurun_out/ncu_full_r02.log
ncu -i gpurun_out/r02_shadow_scan.ncu-rep --page raw --csv > gpurun_out/r02_shadow_scan_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_shadow_scan.ncu-rep --page details > gpurun_out/r02_shadow_scan_details.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r02_c2.csv python tools/c2_probe.py > gpurun_out/ncu_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r02_shard.csv python bench.py --rows 1250000 --steps 4 --warmup 3 --no-cpu-baseline --no-regimes --no-configs --no-gpu-composition > gpurun_out/ncu_launch_shard.log 2>&1
ls -la gpurun_out | tail -12
