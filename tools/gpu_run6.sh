cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
DEVQA_TC_TRACE=1 python de-vqa_b200/build.py --force > /dev/null 2>&1
for st in 0 4 15; do
echo "== Q=128 seed_tiles $st"; python tools/tc_trace.py 10000000 128 $((st << 24)) | tail -4
done
echo "== Q=128 no_insert"; python tools/tc_trace.py 10000000 128 65536 | tail -4
echo "== Q=64 seed 0"; python tools/tc_trace.py 10000000 64 0 | tail -4
echo "== Q=64 no_insert"; python tools/tc_trace.py 10000000 64 65536 | tail -4
echo "== Q=32 seed 0"; python tools/tc_trace.py 10000000 32 0 | tail -4
