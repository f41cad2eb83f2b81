cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DEVQA_P2P=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02_g8_driver.json 2> gpurun_out/bench_r02_g8_driver.err
tail -2 gpurun_out/bench_r02_g8_driver.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_g8_driver.json") if l.startswith("{")][-1])
print("8gpu", round(d["ms_per_step"],4), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "scan_share", round(d["roofline"]["scan_share_of_step"],3), "scan ms", round(d["roofline"]["launch_ms"],4), "agg frac", round(d["aggregate_hbm_frac_of_measured"],3), d["exchange"], d["result_digest"][:16], d["ranks_agree"], d["clocks"]["sm_mhz"])
for r in d.get("regimes", []): print("   ", r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), round(r["useful_tflops"],1), r["exact_fallback_queries"], r["bound_violations"])
PY
