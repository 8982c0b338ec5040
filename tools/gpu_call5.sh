#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
{
python tools/variants.py run --reset next_step --rollout 64 base pf0 a4
python tools/variants.py run --reset same_step --rollout 64 base pf0
python tools/variants.py run --reset next_step --rollout 64 --players 8 --envs 4194304 --steps 256 --preroll 1024 base n8_a9
} > $O/c5_variants.log 2>&1
cat $O/c5_variants.log
