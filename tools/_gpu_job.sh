cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 60 python tools/_wd_probe.py 256 2>&1 | tail -3
timeout 90 python tools/tc_timeline.py > gpurun_out/r2s_timeline.txt 2>&1; rc=$?; echo "timeline rc=$rc"
grep "forward step\|step end\|Ed1 done\|g_h1 ready" gpurun_out/r2s_timeline.txt
if [ $rc -ne 0 ]; then tail -5 gpurun_out/r2s_timeline.txt; exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2s_tests.log
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1fM ms %.3f e2e %.1fM kernel_ms %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['kernel_ms']))"
