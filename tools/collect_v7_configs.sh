T=r1_v7; O=gpurun_out; mkdir -p $O
python bench.py --players 2 --steps 1000 --warmup 20 --no-cpu-baseline --policy-steps 0 > $O/${T}_bench_n2.json 2>/dev/null
python bench.py --indirect --steps 1000 --warmup 20 --no-cpu-baseline --policy-steps 0 > $O/${T}_bench_n4ind.json 2>/dev/null
python bench.py --envs 4194304 --steps 500 --warmup 20 --no-cpu-baseline --e2e-steps 0 --policy-steps 0 > $O/${T}_bench_n4_4m.json 2>/dev/null
python bench.py --players 8 --envs 4194304 --steps 500 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 10 --policy-steps 0 > $O/${T}_bench_n8.json 2>/dev/null
python bench.py --players 8 --envs 16777216 --steps 300 --warmup 20 --preroll 1024 --no-cpu-baseline --e2e-steps 0 --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 > $O/${T}_bench_n8_16m.json 2>/dev/null
for f in n2 n4ind n4_4m n8 n8_16m; do python - <<PY
import json
d=json.load(open("$O/${T}_bench_$f.json"))
print("$f", "%.4g"%d["value"], "kernel_us %.1f"%d["roofline"]["kernel_us"], "frac %.3f"%d["roofline"]["frac"], "e2e", ("%.3g"%d["e2e"]["value"]) if d.get("e2e") else None, "rollout", ("%.4g"%d["rollout"]["value"]) if d.get("rollout") else None, d["episode_stats"]["reshuffles"])
PY
done
