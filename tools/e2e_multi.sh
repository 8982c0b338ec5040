#!/bin/bash
# e2e (skyjo_step_host) at N GPUs of one box with the host-side lap trace of every rank, plus the concurrent copy
# bandwidth of the same GPUs.   usage: bash tools/e2e_multi.sh <tag> <ngpus> [extra env assignments...]
T=$1; G=$2; shift 2
O=gpurun_out
mkdir -p $O
nproc > $O/${T}_host.txt; grep -m1 "model name" /proc/cpuinfo >> $O/${T}_host.txt; lscpu | grep -i -E "numa|socket|thread|core" >> $O/${T}_host.txt
nvidia-smi topo -m >> $O/${T}_host.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541"
[ "$G" = "1" ] && RUN="python"
$RUN tools/d2h_bw.py > $O/${T}_bw_g$G.json 2> $O/${T}_bw_g$G.err
cat $O/${T}_bw_g$G.json
env SKYJO_HOSTIO_TRACE=1 "$@" $RUN bench.py --gpus $G --steps 4 --warmup 2 --preroll 64 --e2e-steps 40 --no-cpu-baseline --rollout-steps 0 \
    --other-reset-steps 0 --policy-steps 0 --no-configs > $O/${T}_e2e_g$G.json 2> $O/${T}_e2e_g$G.err
python - <<PY
import json
try:
    d = json.load(open("$O/${T}_e2e_g$G.json"))["e2e"]
    print("e2e g$G", "%.3e" % d["value"], "env-steps/s", d.get("ms_per_call"), "ms/call")
except Exception as ex:
    print("failed", ex)
PY
grep skyjo_step_host $O/${T}_e2e_g$G.err | tail -$G
