cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_ab.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ab.log
tail -5 gpurun_out/pytest_ab.log
timeout 600 python tools/ab_epi.py 1250000 1,16,32,64,128 2>&1 | tee gpurun_out/ab_resident_1250k.txt
timeout 900 python tools/ab_epi.py 10000000 16,64 1 2>&1 | tee gpurun_out/ab_resident_10m.txt
timeout 900 python tools/ab_epi.py 10000000 16,32 0 2>&1 | tee -a gpurun_out/ab_resident_10m.txt
for f in 0 64; do
  python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-regimes --no-gpu-composition --no-configs --flags $f > gpurun_out/resident_s200_$f.json 2> gpurun_out/resident_s200_$f.err
  python - $f <<'PY'
import json, sys
d=json.loads([l for l in open(f"gpurun_out/resident_s200_{sys.argv[1]}.json") if l.startswith("{")][-1])
print("200 steps, flags", sys.argv[1], "ms", round(d["ms_per_step"],4), "q/s", round(d["value"]), "clocks", d["clocks"], d["result_digest"][:16])
PY
done 2>&1 | tee gpurun_out/resident_200steps.txt
