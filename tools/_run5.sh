timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b5.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['roofline']['kernel_ms'])"
MPG_TAIL_OVERLAP=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('no overlap', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['roofline']['kernel_ms'])"
