#!/bin/bash
# A/B of the mask expansion's store type in skyjo_step_host at N GPUs: plain 64-byte stores (default) against
# streaming stores (SKYJO_HOST_NT_MASK=1).   usage: bash tools/e2e_nt_ab.sh <ngpus>
G=$1
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29571"
[ "$G" = "1" ] && RUN="python"
F="--gpus $G --steps 2 --warmup 1 --preroll 64 --e2e-steps 40 --no-cpu-baseline --rollout-steps 0 --other-reset-steps 0 --policy-steps 0 --no-configs"
for v in 0 1 0 1; do
  SKYJO_HOSTIO_TRACE=1 SKYJO_HOST_NT_MASK=$v $RUN bench.py $F > $O/nt_ab_$v.json 2> $O/nt_ab_$v.err
  python - <<PY
import json
d = json.loads([l for l in open("$O/nt_ab_$v.json") if l.startswith("{")][-1])["e2e"]
print("NT_MASK=$v g$G", "%.3e" % d["value"], "env-steps/s", round(d.get("ms_per_call"), 3), "ms/call, compact ranges", d.get("compact_ranges_of_8"))
PY
  grep skyjo_step_host $O/nt_ab_$v.err | tail -1
done
