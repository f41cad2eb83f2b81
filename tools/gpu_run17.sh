cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-regimes | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('sustained Q64 x200', 'ms', round(d['ms_per_step'],3), 'q/s', round(d['value']), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks'])
"
ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 2 -c 1 -o gpurun_out/filt_r01 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_full_filt.log 2>&1
tail -2 gpurun_out/ncu_full_filt.log
