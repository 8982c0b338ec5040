#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
{
python tools/variants.py run --reset next_step --rollout 64 base n4_r28 n4_r24
python tools/variants.py run --reset next_step --rollout 64 --players 2 base n2_r28
python tools/variants.py run --reset next_step --players 8 --envs 4194304 --steps 256 --preroll 1024 --rollout 64 base
} > $O/c12_variants.log 2>&1
grep rollout $O/c12_variants.log; grep '"N": 8' $O/c12_variants.log
