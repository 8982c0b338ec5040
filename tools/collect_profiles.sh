#!/bin/bash
# Collects the round's measurement evidence on a B200 (run through gpurun); outputs land in gpurun_out/.
# usage: bash tools/collect_profiles.sh <tag>      e.g. r1_v4
set -u
T=${1:-rX}
O=gpurun_out
mkdir -p $O
python bench.py > $O/${T}_bench_n4.json 2> $O/${T}_bench_n4.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_reference_arm.json 2>/dev/null
python bench.py --players 2 --steps 1000 --warmup 20 --no-cpu-baseline > $O/${T}_bench_n2.json 2>/dev/null
python bench.py --indirect --steps 1000 --warmup 20 --no-cpu-baseline > $O/${T}_bench_n4ind.json 2>/dev/null
python bench.py --envs 4194304 --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 0 > $O/${T}_bench_n4_4m.json 2>/dev/null
python bench.py --players 8 --envs 4194304 --steps 500 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 10 > $O/${T}_bench_n8.json 2>/dev/null
python bench.py --players 8 --envs 16777216 --steps 300 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 > $O/${T}_bench_n8_16m.json 2>/dev/null
# the same-step reset mode (episodes desynchronised in phase) for comparison
python bench.py --reset same_step --steps 2000 --warmup 100 --no-cpu-baseline > $O/${T}_bench_n4_same_step.json 2>/dev/null
python bench.py --reset same_step --players 8 --envs 4194304 --steps 500 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 0 > $O/${T}_bench_n8_same_step.json 2>/dev/null
# ncu serialises kernels, so the profiled runs use the single-stream path (SKYJO_RANGES=1: full-batch launches)
export SKYJO_RANGES=1
# launch list of the timed loop (cold-cache, serialised per-launch times: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 730 -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 400 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > /dev/null 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 2 -f -o $O/${T}_step_full \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -f -o $O/${T}_rollout_full \
    python bench.py --steps 16 --warmup 3 --preroll 320 --e2e-steps 0 --no-cpu-baseline --rollout-steps 16 --other-reset-steps 0 --policy-steps 0 > $O/${T}_rollout_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1100 -c 2 -f -o $O/${T}_step_n8_full \
    python bench.py --players 8 --envs 4194304 --preroll 1024 --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:deal_kernel -s 20 -c 1 -f -o $O/${T}_deal_full \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:pack_host_kernel -s 2 -c 1 -f -o $O/${T}_pack_full \
    python bench.py --steps 8 --warmup 3 --preroll 64 --e2e-steps 3 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > /dev/null 2>&1
unset SKYJO_RANGES
ls -la $O
