cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu -x --timeout=900 > gpurun_out/pytest_gpu_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_all.log
tail -8 gpurun_out/pytest_gpu_all.log
for pdl in 0 1; do
  echo "== DEVQA_PDL=$pdl"
  DEVQA_PDL=$pdl timeout 300 python tools/c2_time.py 2>&1 | grep -v "3xTF32\|list epi" | tee gpurun_out/c2_time_pdl$pdl.txt
  for rows in 1250000 10000000; do
    DEVQA_PDL=$pdl python bench.py --rows $rows --steps 200 --warmup 5 --no-regimes --no-configs --no-gpu-composition --no-cpu-baseline | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('rows', d['config']['bank_rows'], 'ms', round(d['ms_per_step'],4), 'q/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'scan share', round(d['roofline']['scan_share_of_step'],4), 'scan ms', round(d['roofline']['launch_ms'],4), d['clocks']['sm_mhz'], d['e2e']['matches_device_path'])
"
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python tools/c2_probe.py > gpurun_out/ncu_c2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches_c2.csv") if not l.startswith("=="))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki][:70], float(r[vi].replace(",",""))/1e3) for r in rows[1:]]
for k,v in seq[-7:]: print(f"{v:9.1f} us  {k}")
PY
