#!/bin/bash
set -u
N=8; T=$1; O=gpurun_out; mkdir -p $O
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --envs 8388608 --steps 400 --warmup 20 --no-cpu-baseline --e2e-steps 0 --rollout-steps 16 > $O/${T}_bench_g8_64m.json 2> $O/${T}_bench_g8_64m.err
tail -2 $O/${T}_bench_g8_64m.err; cat $O/${T}_bench_g8_64m.json | cut -c1-400
