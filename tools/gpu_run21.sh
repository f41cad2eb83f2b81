cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/c1_probe.py
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 40 --csv --log-file gpurun_out/launches_c1.csv python tools/c1_probe.py > gpurun_out/c1_probe.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_c1.csv
