#!/bin/bash
# e2e (skyjo_step_host) sweep over wire mode / host threads / env ranges on a B200 box; outputs in gpurun_out/.
# usage: bash tools/e2e_sweep.sh <tag> [players] [ngpus] [final|multi|nt|threads]
#   (default) one run per wire mode / fixed share;  final: measured share vs raw, twice;  multi: the same set under
#   torchrun at <ngpus>;  nt: streaming vs plain stores A/B (SKYJO_HOST_NT_MASK / SKYJO_HOST_NT are experiment knobs of
#   csrc/skyjo_hostsimd.cpp);  threads: worker count, spin and pinning knobs.
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
if [ "$4" = "final" ]; then
COMMON="${COMMON/--e2e-steps 40/--e2e-steps 100}"
for rep in 1 2; do
run auto_$rep
run raw_$rep SKYJO_HOST_WIRE=raw
run raw_t4_$rep SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=4
done
python - <<PY
import json
for n in ("auto_1", "auto_2"):
    print(n, "compact ranges of 8 at the end:", json.load(open("$O/${T}_%s.json" % n))["e2e"].get("compact_ranges_of_8"))
PY
exit 0
fi
if [ "$4" = "multi" ]; then
COMMON="${COMMON/--e2e-steps 40/--e2e-steps 100}"
run auto
run raw SKYJO_HOST_WIRE=raw
run mix4 SKYJO_HOST_MIX=4
run mix6 SKYJO_HOST_MIX=6
run cmp SKYJO_HOST_WIRE=compact
python - <<PY
import json
print("auto: compact ranges of 8 at the end (rank 0):", json.load(open("$O/${T}_auto.json"))["e2e"].get("compact_ranges_of_8"))
PY
exit 0
fi
if [ "$4" = "nt" ]; then
for rep in 1 2; do
run raw_t4_nt_$rep SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=4 SKYJO_HOST_NT_MASK=1
run raw_t4_plain_$rep SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=4
run raw_t2_plain_$rep SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=2
run mix5_t8_nt_$rep SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=8
run mix5_t8_plain_$rep SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=8 SKYJO_HOST_NT=0
run mix6_t8_nt_$rep SKYJO_HOST_MIX=6 SKYJO_HOST_THREADS=8
run mix4_t8_nt_$rep SKYJO_HOST_MIX=4 SKYJO_HOST_THREADS=8
run mix5_t6_nt_$rep SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=6
run mix5_t10_nt_$rep SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=10
done
exit 0
fi
if [ "$4" = "threads" ]; then
for t in 2 4 6 8 12; do run raw_t$t SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=$t; done
run raw_t12_nospin SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=12 SKYJO_HOST_SPIN=0
run raw_t12_nopin SKYJO_HOST_WIRE=raw SKYJO_HOST_THREADS=12 SKYJO_HOST_PIN=0
for t in 4 8 12 15; do run mix5_t$t SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=$t; done
run mix5_t12_nospin SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=12 SKYJO_HOST_SPIN=0
run mix5_t12_nopin SKYJO_HOST_MIX=5 SKYJO_HOST_THREADS=12 SKYJO_HOST_PIN=0
exit 0
fi
run mixed_default
run raw_default SKYJO_HOST_WIRE=raw
run mix4 SKYJO_HOST_MIX=4
run mix5 SKYJO_HOST_MIX=5
run mix6 SKYJO_HOST_MIX=6
run mix7 SKYJO_HOST_MIX=7
run cmp_default SKYJO_HOST_WIRE=compact

