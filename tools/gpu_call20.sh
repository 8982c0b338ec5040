#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
( time python bench.py > $O/c20_bench.json 2> $O/c20.err ) 2> $O/c20_time.txt
tail -3 $O/c20.err; cat $O/c20_time.txt | tail -4
python - <<'P'
import json
d=json.load(open("gpurun_out/c20_bench.json"))
print(d["policy_rollout"]); print("value %.4g"%d["value"], d["roofline"]["frac"], d["roofline"]["traffic_frac"], "e2e %.3g"%d["e2e"]["value"], d["cpu_baseline"]["value"])
P
( time python bench.py --impl reference --steps 3 --warmup 1 > $O/c20_ref.json ) 2>&1 | tail -3
