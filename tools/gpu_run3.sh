cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for Q in 128 64; do
for f in 0 65536 524288 1048576 2097152; do timeout 200 python bench.py --rows 4000000 --queries $Q --steps 10 --warmup 3 --no-cpu-baseline --no-regimes --flags $f 2>&1 | python -c "
import sys,json
names={0:'baseline',65536:'no_insert',524288:'no_query_tma',1048576:'no_l2_prefetch',2097152:'no_seed'}
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('Q $Q', names[$f].ljust(16), 'ms', round(d['ms_per_step'],3), 'scan ms', round(d['roofline']['launch_ms'],3), 'clk', d['clocks']['sm_mhz'], 'cycles/kblock@1.965', round(d['roofline']['launch_ms']*1e-3*1.965e9/(4000000/128/148*24)))
    else: print(ln.rstrip()[:200])
"; done; done 2>&1 | tee gpurun_out/bisect_filt.txt
ncu --set full --clock-control none --import-source on --kernel-name regex:scan_tc_kernel --launch-skip 3 --launch-count 1 -o gpurun_out/filt_q128 -f python bench.py --rows 4000000 --queries 128 --steps 1 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_filt_q128.log 2>&1
tail -3 gpurun_out/ncu_filt_q128.log
