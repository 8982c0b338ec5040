T=r1_v5; O=gpurun_out; mkdir -p $O
export SKYJO_RANGES=1
ncu --metrics gpu__time_duration.sum --clock-control none -s 730 -c 400 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 400 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 2 -f -o $O/${T}_step_full \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
ls -la $O | tail -4
