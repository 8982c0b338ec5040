#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
{
python tools/variants.py run --reset next_step base
python tools/variants.py run --reset next_step --players 8 --envs 4194304 --steps 256 --preroll 1024 base
} > $O/c7_variants.log 2>&1
cat $O/c7_variants.log
export SKYJO_RANGES=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deal_kernel -s 20 -c 1 -f -o $O/c7_deal_n4 \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
