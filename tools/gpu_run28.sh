cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:scan_tc_kernel<.int.256' -s 2 -c 1 -o gpurun_out/filt_pair_r01 -f python bench.py --queries 256 --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_full_pair.log 2>&1
tail -2 gpurun_out/ncu_full_pair.log
