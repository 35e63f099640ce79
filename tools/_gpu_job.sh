cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
PYTHONFAULTHANDLER=1 timeout -s ABRT 300 python bench.py --no-cpu-baseline > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"
cat gpurun_out/r2m_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','updates_per_s','gpu_launches')}, d['e2e'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['clocks'])"
tail -5 gpurun_out/r2m_bench.err
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
tail -8 gpurun_out/r2m_gpu_tests.log
