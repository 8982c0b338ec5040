#!/bin/bash
# ncu evidence of the final tree (single-stream path, as in collect_profiles.sh): launch list of the bench loop and a
# full capture of two consecutive step-kernel launches (a draw slot and a place slot)
T=r1_v7; O=gpurun_out; mkdir -p $O
export SKYJO_RANGES=1
F="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -s 730 -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 400 --warmup 10 $F > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 2 -f -o $O/${T}_step_full \
    python bench.py --steps 100 --warmup 10 $F > /dev/null 2>&1
ls -la $O | grep $T | grep -v json
