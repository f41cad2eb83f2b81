cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pools.py -x -q -m gpu > gpurun_out/pytest_pools.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_pools.log
tail -12 gpurun_out/pytest_pools.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches_shard.csv python bench.py --rows 1250000 --steps 4 --warmup 3 --no-cpu-baseline --no-regimes > gpurun_out/ncu_launch_shard.log 2>&1
python profiles/summarize.py launches gpurun_out/launches_shard.csv
