#!/bin/bash
# usage: tools/run_scale.sh N [extra bench args]   -- bench.py on N GPUs of this box under torchrun
N=$1; shift
if [ "$N" = "1" ]; then python bench.py --gpus 1 "$@"; else
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; fi
