cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2p_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
grep -E "bench path|worst over|^seed [0-9]|n=25|passed|failed|^FAILED|^E  +assert|worst per-step" gpurun_out/r2p_gpu_tests.log | tail -40
timeout 120 python tools/_hang_probe.py dev 2>&1 | tail -2
