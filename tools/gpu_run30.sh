cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_filter.py tests/test_gpu_pools.py -x -q -m gpu -k "bit_for_bit or midsize or duplicated or nan_rows or hice or liveedit or closest" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|error" gpurun_out/memcheck.log | head -20
