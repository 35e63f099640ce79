#!/bin/bash
# bench.py at N GPUs of one box the way the driver launches it (one rank per GPU over NCCL):
#   weak scaling of the headline workload (config 2), strong scaling of config 4 (1,048,576-row MPG) and of one 65,536-row
#   NADP batch.  Lines are appended to gpurun_out/r2_scaling_<N>gpu.jsonl.
N=${1:-2}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
OUT=gpurun_out/r2_scaling_${N}gpu.jsonl
run() {
  if [ "$N" = 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu-baseline "$@" >> $OUT 2>> gpurun_out/r2_scaling.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
         bench.py --gpus $N "$@" >> $OUT 2>> gpurun_out/r2_scaling.err; fi
  echo "N=$N bench $* rc=$?"
}
WHICH=${2:-all}      # all | weak | weak+c4
run --steps 10 --warmup 3
if [ "$WHICH" != weak ]; then run --config 4 --steps 5 --warmup 3; fi
if [ "$WHICH" = all ]; then run --global-rows 65536 --steps 10 --warmup 3; fi
tail -3 $OUT | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['n_gpus'], d['scaling'], d['config']['global_batch'], 'value %.1fM ms %.3f e2e %.1fM' % (d['value'] / 1e6, d['ms_per_step'], d['e2e']['value'] / 1e6), d.get('mgpu_check'), d.get('mgpu_check_detail', {}).get('rel_l2_vs_single_rank_global_batch'))
"
