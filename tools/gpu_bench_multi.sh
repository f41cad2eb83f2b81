cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r01b_g8.json 2> gpurun_out/bench_r01b_g8.err
tail -2 gpurun_out/bench_r01b_g8.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r01b_g8.json") if l.startswith("{")][-1])
print("8gpu headline", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "scan_share", d["roofline"]["scan_share_of_step"], d["clocks"], "fallbacks", d["exact_fallback_queries"])
for r in d["regimes"]: print(r["queries_per_call"], round(r["ms_per_call"],3), round(r["queries_per_s"]), round(r["frac_of_measured_hbm"],3), r["exact_fallback_queries"])
PY
