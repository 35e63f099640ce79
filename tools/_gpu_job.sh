cd $GRAFT_REPO_ROOT
timeout 90 python tools/tc_timeline.py 2>&1 | grep -v "acc_wait\|waiting for\|load issued"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "tc or golden" 2>&1 | tail -2
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value %.1fM ms %.3f e2e %.1fM kernel_ms %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['kernel_ms']))"
