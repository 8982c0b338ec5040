#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/c9_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c9_pytest.log
tail -4 $O/c9_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c9_smoke.log 2>&1; tail -2 $O/c9_smoke.log
{
for R in same_step next_step; do
python tools/quick_bench.py --tag c9 --reset $R --players 8 --envs 4194304 --steps 256 --preroll 1024 --rollout 64
done
python tools/quick_bench.py --tag c9 --reset next_step --rollout 64
python tools/quick_bench.py --tag c9 --reset next_step --players 12 --envs 1048576 --steps 256 --preroll 1400
python tools/quick_bench.py --tag c9 --reset next_step --players 6 --envs 1048576 --steps 256 --preroll 1000
python tools/quick_bench.py --tag c9 --reset next_step --players 5 --envs 1048576 --steps 256 --preroll 1000
} > $O/c9_quick.log 2>&1
cat $O/c9_quick.log
