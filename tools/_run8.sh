timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "nadp or wave_tail" 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b8.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['e2e']['value'], d['roofline']['kernel_ms'])"
timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -7
