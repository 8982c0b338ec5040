#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --e2e-steps 0 > $O/c14_n4.json 2>$O/c14.err
python bench.py --players 8 --envs 4194304 --steps 500 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 0 > $O/c14_n8.json 2>>$O/c14.err
python bench.py --players 2 --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 0 > $O/c14_n2.json 2>>$O/c14.err
python - <<'P'
import json
for f in ("c14_n4","c14_n8","c14_n2"):
    d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
    print(f, "value %.4g"%d["value"], "ms/step %.4f"%d["ms_per_step"], "kernel_us %.2f"%r["kernel_us"], "frac %.3f"%r["frac"], "loop_frac %.3f"%r["loop_frac"], "deal_share %.3f"%r["deal_kernel_share"])
P
tail -3 $O/c14.err
