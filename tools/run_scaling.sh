#!/bin/bash
# bench.py at N GPUs of one box the way the driver launches it (one rank per GPU over NCCL); stdout = the JSON line
N=${1:-2}
if [ "$N" = 1 ]; then exec python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline; fi
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3
