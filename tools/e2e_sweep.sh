#!/bin/bash
# e2e (skyjo_step_host) sweep over wire mode / host threads / env ranges on a B200 box; outputs in gpurun_out/.
# usage: bash tools/e2e_sweep.sh <tag> [players] [ngpus]
T=${1:-e2e}
N=${2:-4}
G=${3:-1}
O=gpurun_out
mkdir -p $O
nproc > $O/${T}_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> $O/${T}_nproc.txt
RUN="python"
[ "$G" != "1" ] && RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29547"
COMMON="--gpus $G --players $N --steps 2 --warmup 1 --preroll 64 --e2e-steps 40 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 --no-configs"
run() {  # name, env assignments...
  name=$1; shift
  env SKYJO_HOSTIO_TRACE=1 "$@" $RUN bench.py $COMMON > $O/${T}_$name.json 2> $O/${T}_$name.err
  python - <<PY
import json
try:
    d = json.load(open("$O/${T}_$name.json"))["e2e"]
    print("$name", "%.3e" % d["value"], "env-steps/s", d.get("wire_bytes_per_env"), "B/env", "%.3f ms/call" % d["ms_per_call"])
except Exception as ex:
    print("$name failed", ex)
PY
  grep skyjo_step_host $O/${T}_$name.err | tail -$G
}
run raw_default SKYJO_HOST_WIRE=raw
run raw_t4 SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=4
run raw_nopin SKYJO_HOST_WIRE=raw SKYJO_HOST_PIN=0
run cmp_default SKYJO_HOST_WIRE=compact
run cmp_t4 SKYJO_HOST_WIRE=compact SKYJO_HOST_THREADS=4
run cmp_t8 SKYJO_HOST_WIRE=compact SKYJO_HOST_THREADS=8
run cmp_c4 SKYJO_HOST_WIRE=compact SKYJO_HOST_CHUNKS=4
