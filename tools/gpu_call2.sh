#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c4_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c4_pytest.log
tail -15 $O/c4_pytest.log
{
for R in same_step next_step; do
timeout 300 python tools/quick_bench.py --tag c4 --reset $R --rollout 64
timeout 300 python tools/quick_bench.py --tag c4 --reset $R --players 8 --envs 4194304 --steps 256 --preroll 1024 --rollout 64
timeout 300 python tools/quick_bench.py --tag c4 --reset $R --players 2 --rollout 64
done
} > $O/c4_quick.log 2>&1
cat $O/c4_quick.log
