#!/bin/bash
# usage: bash tools/gpu_multi.sh <n_gpus> <tag>
set -u
N=$1; T=$2; O=gpurun_out; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 2000 --warmup 100 --no-cpu-baseline > $O/${T}_bench_g$N.json 2> $O/${T}_bench_g$N.err
tail -2 $O/${T}_bench_g$N.err; cat $O/${T}_bench_g$N.json | cut -c1-600
