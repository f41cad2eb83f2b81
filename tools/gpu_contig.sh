# timing experiment: scan kernel duration with strided (real) vs contiguous (garbage results) tile boxes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for fl in 0 128; do
  for rows in 10000000 1250000; do
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k scan_tc_kernel -c 12 --csv --log-file gpurun_out/contig_${fl}_${rows}.csv python tools/one_call.py $rows 64 1 $fl 3 > /dev/null 2>&1
    python - $fl $rows <<'PY'
import csv, sys
fl, rows = sys.argv[1], sys.argv[2]
rs=[r for r in csv.reader(l for l in open(f"gpurun_out/contig_{fl}_{rows}.csv") if not l.startswith("=="))]
h=rs[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value")
d=[(r[ki][:60], r[mi], float(r[vi].replace(",",""))) for r in rs[1:]]
main=[v for k,m,v in d if m=="gpu__time_duration.sum" and v>50000]
print(f"flags={fl} rows={rows}: main scan durations (us):", [round(v/1e3,1) for v in main])
PY
  done
done 2>&1 | tee gpurun_out/contig_experiment.txt
