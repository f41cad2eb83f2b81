cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_tc.py -q -m gpu -x --timeout=600 > gpurun_out/pytest_filter.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_filter.log
tail -15 gpurun_out/pytest_filter.log
timeout 300 python tools/c2_time.py 2>&1 | tee gpurun_out/c2_time.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python tools/c2_probe.py $C2FLAGS > gpurun_out/ncu_c2.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/launches_c2.csv") if not l.startswith("=="))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki][:70], float(r[vi].replace(",",""))/1e3) for r in rows[1:]]
for k,v in seq[-14:]: print(f"{v:9.1f} us  {k}")
PY
