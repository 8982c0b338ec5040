"""Concurrent device->host / host->device copy bandwidth of the GPUs of one box (pinned host memory), per rank and in
aggregate -- the ceiling of the host-buffer entry skyjo_step_host at N GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/d2h_bw.py
"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"world": world, "host_cores": os.cpu_count()}
    for mb in (8, 74, 256):
        n = mb << 20
        d = torch.empty(n, dtype=torch.uint8, device=dev)
        h = torch.empty(n, dtype=torch.uint8).pin_memory()
        for direction in ("d2h", "h2d"):
            def go():
                if direction == "d2h":
                    h.copy_(d, non_blocking=True)
                else:
                    d.copy_(h, non_blocking=True)
            for _ in range(3):
                go()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            reps = max(4, 2048 // mb)
            t0 = time.perf_counter()
            for _ in range(reps):
                go()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            mine = n * reps / float(dt.item()) / 1e9
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            agg = world * n * reps / float(dt.item()) / 1e9
            per = torch.tensor([mine], device=dev, dtype=torch.float64)
            if world > 1:
                g = [torch.empty_like(per) for _ in range(world)]
                dist.all_gather(g, per)
                per_all = [round(float(x.item()), 1) for x in g]
            else:
                per_all = [round(mine, 1)]
            out[f"{direction}_{mb}MB"] = {"aggregate_GBs": round(agg, 1), "per_rank_GBs": per_all}
        del d, h
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
