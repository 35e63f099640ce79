timeout 200 python tools/gemm_probe.py 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/b6.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('overlap', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['roofline']['kernel_ms'])"
