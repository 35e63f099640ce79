timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['roofline']['kernel_ms'])"
timeout 300 python tools/tc_timeline.py 2>&1 | head -14
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q 2>&1 | tail -3
