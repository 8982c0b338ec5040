#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
export SKYJO_RANGES=1
for S in 1100 1101; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $S -c 1 -f -o $O/c6_step_n8_$S \
    python bench.py --players 8 --envs 4194304 --preroll 1024 --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
done
for S in 700 701; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s $S -c 1 -f -o $O/c6_step_n4_$S \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 44 -c 1 -f -o $O/c6_rollout_n4 \
    python bench.py --steps 16 --warmup 3 --preroll 320 --e2e-steps 0 --no-cpu-baseline --rollout-steps 16 > $O/c6_ro.log 2>&1
ls -la $O | grep c6_
