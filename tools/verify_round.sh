#!/bin/bash
# End-of-round check on a GPU box: GPU test suite, smoke, both bench arms at N=1, and the N=2 torchrun launch.
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/verify_bench_1gpu.json 2> gpurun_out/verify_bench_1gpu.err; wc -l < gpurun_out/verify_bench_1gpu.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/verify_bench_ref.json 2>/dev/null; wc -l < gpurun_out/verify_bench_ref.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 280 bash tools/run_scaling.sh 2 > gpurun_out/verify_bench_2gpu.json 2> gpurun_out/verify_bench_2gpu.err; wc -l < gpurun_out/verify_bench_2gpu.json
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/verify_bench_*.json')):
    d = json.load(open(f))
    print(f, d.get('impl', 'ours'), d['n_gpus'], round(d['value'] / 1e6, 2), round(d['ms_per_step'], 3), d.get('e2e', {}).get('value'), d.get('clocks'))
PY
