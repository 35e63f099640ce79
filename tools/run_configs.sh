#!/bin/bash
# One bench.py line per BASELINE.json configuration on ONE GPU -> gpurun_out/r2_configs.jsonl
# (config 4 on 1 GPU is the strong-scaling base; tools/run_scaling.sh adds the 2/4/8-GPU lines)
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
OUT=gpurun_out/r2_configs.jsonl
: > $OUT
run() { timeout 600 python bench.py --no-cpu-baseline "$@" >> $OUT 2>> gpurun_out/r2_configs.err; echo "bench $* rc=$?"; }
run --config 1
run --config 2
for env in ip idp; do
  for r in 1024 16384 131072 1048576; do run --config 3 --env $env --rows $r --steps 5; done
done
run --rows 262144
run --config 4 --steps 3
run --config 5 --steps 5
