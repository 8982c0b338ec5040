#!/bin/bash
# round-1 session-4 call 1: GPU suite on a fresh box, quick timings, source-level ncu captures
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
tail -5 $O/c1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/c1_smoke.log 2>&1; tail -2 $O/c1_smoke.log
{
timeout 300 python tools/quick_bench.py --tag base --rollout 64
timeout 300 python tools/quick_bench.py --tag base --players 8 --envs 4194304 --steps 256 --preroll 1024 --rollout 64
timeout 300 python tools/quick_bench.py --tag base --players 2 --rollout 64
} > $O/c1_quick.log 2>&1
cat $O/c1_quick.log
export SKYJO_RANGES=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 700 -c 1 -f -o $O/c1_step_n4 \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 4 -c 1 -f -o $O/c1_rollout_n4 \
    python bench.py --steps 16 --warmup 3 --preroll 320 --e2e-steps 0 --no-cpu-baseline --rollout-steps 16 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 1100 -c 1 -f -o $O/c1_step_n8 \
    python bench.py --players 8 --envs 4194304 --preroll 1024 --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deal_kernel -s 20 -c 1 -f -o $O/c1_deal_n4 \
    python bench.py --steps 100 --warmup 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 > /dev/null 2>&1
ls -la $O | tail -12
