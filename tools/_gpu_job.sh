cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 90 python tools/tc_timeline.py > gpurun_out/r2l_timeline.txt 2>&1; rc=$?; echo "timeline rc=$rc"
cat gpurun_out/r2l_timeline.txt | grep -v "acc_wait done\|waiting for h2" | tail -42
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q > gpurun_out/r2l_parity.log 2>&1; echo "parity rc=$?"
tail -8 gpurun_out/r2l_parity.log
timeout 120 python tools/_hang_probe.py dev 2>&1 | tail -2
timeout 120 python tools/_hang_probe.py cg 2>&1 | tail -2
