cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 900 python -m pytest tests/test_gpu_dist.py -q -m gpu -x --timeout=600 > gpurun_out/pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_dist.log
tail -30 gpurun_out/pytest_dist.log
cat gpurun_out/dist_exchange_2gpu.json
for p2p in 1 0; do
DEVQA_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 5 --no-regimes > gpurun_out/bench_r02_g2_p2p$p2p.json 2> gpurun_out/bench_r02_g2.err
tail -3 gpurun_out/bench_r02_g2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r02_g2_p2p$p2p.json") if l.startswith("{")][-1])
print("2gpu p2p=$p2p", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "scan_share", d["roofline"]["scan_share_of_step"], "digest", d["result_digest"][:16], d["ranks_agree"], "launches", d["gpu_launches"])
PY
done
