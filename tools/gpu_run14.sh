cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c2.csv python tools/c2_probe.py > gpurun_out/c2_probe.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_c2.csv
