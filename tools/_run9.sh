for lib in libmpg_b200.so libmpg_b200_late.so; do
MPG_B200_LIB=$PWD/mpg_b200/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', d['value'], d['ms_per_step'], d['e2e']['ms_per_update'], d['roofline']['kernel_ms'])"
done
MPG_B200_LIB=$PWD/mpg_b200/libmpg_b200_late.so timeout 300 python tools/tc_timeline.py 2>&1 | head -14
MPG_B200_LIB=$PWD/mpg_b200/libmpg_b200_late.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "nadp or wave_tail or full_size" 2>&1 | tail -3
