cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py -q -m gpu -x --timeout=900 > gpurun_out/pytest_ab.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ab.log
tail -3 gpurun_out/pytest_ab.log
timeout 600 python tools/ab_epi.py 1250000 64,128,256,1024 1 2>&1 | tee gpurun_out/ab_respair_1250k.txt
AB_DIM=384 timeout 600 python tools/ab_epi.py 2500000 128,256,1024 2>&1 | tee gpurun_out/ab_respair_d384.txt
timeout 900 python tools/ab_epi.py 10000000 128 1 2>&1 | tee gpurun_out/ab_respair_10m.txt
