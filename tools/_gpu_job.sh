cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2z_tests.log
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1fM ms %.3f e2e %.1fM (%.3f ms) kernel_ms %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['ms_per_update'], d['roofline']['kernel_ms']))"
timeout 200 python tools/tc_timeline.py > gpurun_out/r2z_timeline.txt 2>&1; head -12 gpurun_out/r2z_timeline.txt
