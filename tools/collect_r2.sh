#!/bin/bash
# Round-2 measurement evidence on one B200 (run through gpurun); outputs land in gpurun_out/.
# usage: bash tools/collect_r2.sh <tag>      e.g. r2_v6
set -u
T=${1:-r2_vX}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/${T}_pytest_gpu.txt
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_reference_arm.json 2> $O/${T}_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/${T}_bench_n4.json 2> $O/${T}_bench_n4.err
python bench.py --players 2 --steps 20 --warmup 5 --no-cpu-baseline --no-configs --policy-steps 0 > $O/${T}_bench_n2.json 2>/dev/null
python bench.py --indirect --steps 20 --warmup 5 --no-cpu-baseline --no-configs --policy-steps 0 > $O/${T}_bench_n4ind.json 2>/dev/null
python tools/policy_kernel_bench.py > $O/${T}_policy_kernel.txt 2>&1
python tools/policy_trace.py >> $O/${T}_policy_kernel.txt 2>&1
# ncu serialises kernels, so the profiled runs use the single-stream path (SKYJO_RANGES=1: full-batch launches)
export SKYJO_RANGES=1
F="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 --no-configs"
# launch list of the bench command's timed loop (cold-cache, serialised per-launch times: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 740 -c 480 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 7 --warmup 1 $F > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 720 -c 2 -f -o $O/${T}_step_full \
    python bench.py --steps 2 --warmup 1 $F > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:deal_kernel -s 12 -c 1 -f -o $O/${T}_deal_full \
    python bench.py --steps 2 --warmup 1 $F > /dev/null 2>&1
unset SKYJO_RANGES
ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 3 -c 1 -f -o $O/${T}_policy_full \
    python tools/policy_kernel_bench.py --launches 2 > /dev/null 2>&1
ls -la $O | grep $T
