#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 600 python bench.py --steps 2000 --warmup 100 > $O/c3_bench_n4.json 2> $O/c3_bench_n4.err; tail -3 $O/c3_bench_n4.err; cat $O/c3_bench_n4.json
export SKYJO_RANGES=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 2 -f -o $O/c3_step_n4 \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -f -o $O/c3_rollout_n4 \
    python bench.py --steps 16 --warmup 3 --preroll 320 --e2e-steps 0 --no-cpu-baseline --rollout-steps 16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1100 -c 2 -f -o $O/c3_step_n8 \
    python bench.py --players 8 --envs 4194304 --preroll 1024 --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
ls -la $O | grep c3_
