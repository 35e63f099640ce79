cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2c_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/r2c_parity.log
timeout 300 python tools/tc_timeline.py > gpurun_out/r2c_timeline.txt 2>&1; echo "timeline rc=$?"
cat gpurun_out/r2c_timeline.txt
PYTHONFAULTHANDLER=1 timeout -s ABRT 400 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
cat gpurun_out/r2c_bench.json; tail -30 gpurun_out/r2c_bench.err
