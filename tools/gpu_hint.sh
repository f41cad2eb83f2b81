# sustained (200 steps) headline step with different suspend-time hints on the scan kernel's mbarrier waits
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for h in 0 2000 100000 0; do
  DEVQA_WAIT_HINT_NS=$h python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline --no-regimes --no-gpu-composition --no-configs > gpurun_out/hint_$h.json 2> gpurun_out/hint_$h.err
  python - $h <<'PY'
import json, sys
d=json.loads([l for l in open(f"gpurun_out/hint_{sys.argv[1]}.json") if l.startswith("{")][-1])
print("hint", sys.argv[1], "ms", round(d["ms_per_step"],4), "q/s", round(d["value"]), "clocks", d["clocks"])
PY
done 2>&1 | tee gpurun_out/wait_hint_200steps.txt
