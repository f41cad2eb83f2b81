cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:scan_tc_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/filt_q128 -f python bench.py --rows 4000000 --queries 128 --steps 1 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_filt_q128.log 2>&1
tail -2 gpurun_out/ncu_filt_q128.log
DEVQA_TC_TRACE=1 python de-vqa_b200/build.py --force > /dev/null 2>&1
python tools/tc_trace.py 10000000 128 > gpurun_out/trace_filt_q128.txt 2>&1
python tools/tc_trace.py 10000000 64 > gpurun_out/trace_filt_q64.txt 2>&1
cat gpurun_out/trace_filt_q128.txt | cut -c1-400
