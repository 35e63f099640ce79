set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; tail -c 600 gpurun_out/bench_r1_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_rollout_kernel -s 4 -c 2 -o gpurun_out/prof_rollout_r1b python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_dw_kernel -s 4 -c 2 -o gpurun_out/prof_dw_r1b python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
ls -la gpurun_out/
