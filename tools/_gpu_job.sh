cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tools/tc_timeline.py > gpurun_out/r2w_timeline.txt 2>&1; echo "timeline rc=$?"
