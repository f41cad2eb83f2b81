# ncu --set full of ONE 256-query CTA-pair scan launch (10M x 768 fp16 shadow) -- after the plain command exited 0
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/one_call.py 10000000 1024 1 0 2 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k scan_tc_kernel -s 21 -c 1 -o gpurun_out/r02_pair_scan -f python tools/one_call.py 10000000 1024 1 0 2 > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
ncu -i gpurun_out/r02_pair_scan.ncu-rep --page raw --csv > gpurun_out/r02_pair_scan_raw.csv 2>/dev/null
python profiles/summarize.py kernel gpurun_out/r02_pair_scan.ncu-rep | tee gpurun_out/r02_pair_scan_summary.txt | head -40
