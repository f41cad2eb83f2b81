# ncu launch list (cold-cache, serialised per-kernel durations) of ONE search: ROWS Q SHADOW TAG
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  timeout 300 python tools/one_call.py $1 $2 $3 || exit 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$4.csv python tools/one_call.py $1 $2 $3 > gpurun_out/ncu_$4.log 2>&1
  python - $4 <<'PY'
import csv, sys
tag = sys.argv[1]
rows=[r for r in csv.reader(l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("=="))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki][:90], float(r[vi].replace(",",""))/1e3) for r in rows[1:]]
idx=[i for i,(k,_) in enumerate(seq) if "prep_queries_filt" in k]
last=seq[idx[-1]:] if idx else seq[-12:]
with open(f"gpurun_out/launches_{tag}.txt","w") as f:
    for k,v in last:
        line=f"{v:9.1f} us  {k}"; print(line); f.write(line+"\n")
    tot=sum(v for _,v in last); line=f"{tot:9.1f} us  total"; print(line); f.write(line+"\n")
PY
}
for spec in $SPECS; do
  IFS=: read rows q sh tag <<< "$spec"
  run $rows $q $sh $tag
done
