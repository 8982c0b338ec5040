#!/bin/bash
# e2e (skyjo_step_host) sweep over wire mode / host threads / env ranges on a B200 box; outputs in gpurun_out/.
# usage: bash tools/e2e_sweep.sh <tag> [players]
T=${1:-e2e}
N=${2:-4}
O=gpurun_out
mkdir -p $O
nproc > $O/${T}_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> $O/${T}_nproc.txt
COMMON="--players $N --steps 8 --warmup 3 --preroll 64 --e2e-steps 40 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0"
run() {  # name, env assignments...
  name=$1; shift
  env SKYJO_HOSTIO_TRACE=1 "$@" python bench.py $COMMON > $O/${T}_$name.json 2> $O/${T}_$name.err
  python - <<PY
import json
try:
    d = json.load(open("$O/${T}_$name.json"))["e2e"]
    print("$name", "%.3e" % d["value"], "env-steps/s", d.get("wire_bytes_per_env"), "B/env")
except Exception as ex:
    print("$name failed", ex)
PY
  grep skyjo_step_host $O/${T}_$name.err | tail -1
}
run raw_t4_c4 SKYJO_HOST_WIRE=raw
run cmp_t4_c8 SKYJO_HOST_THREADS=4
run cmp_t8_c8 SKYJO_HOST_THREADS=8
run cmp_t12_c8 SKYJO_HOST_THREADS=12
run cmp_t16_c8 SKYJO_HOST_THREADS=16
run cmp_t8_c4 SKYJO_HOST_THREADS=8 SKYJO_HOST_CHUNKS=4
run cmp_t16_c4 SKYJO_HOST_THREADS=16 SKYJO_HOST_CHUNKS=4
