cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DEVQA_TC_TRACE=1 python de-vqa_b200/build.py --force > /dev/null 2>&1
echo "== Q=256 pair"; timeout 120 python tools/tc_trace.py 10000000 256 0 | sed -n '1,6p' | cut -c1-330
echo "== Q=256 no pair"; timeout 120 python tools/tc_trace.py 10000000 256 268435456 | sed -n '1,6p' | cut -c1-330
