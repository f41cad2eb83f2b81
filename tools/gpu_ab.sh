cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/ab_epi.py 1250000 16,64,128,256 2>&1 | tee gpurun_out/ab_defer_1250k.txt
timeout 900 python tools/ab_epi.py 10000000 64,256 1 2>&1 | tee gpurun_out/ab_defer_10m.txt
