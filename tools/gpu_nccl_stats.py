"""Multi-rank check of skyjo_stats_allreduce (run under torchrun on N GPUs of one box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/gpu_nccl_stats.py
Every rank steps its own env shard; the in-library ncclAllReduce of the statistics must equal torch.distributed's
all_reduce of the same vectors, and the sum of the per-rank vectors gathered on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200 import BatchedSkyjoEnv  # noqa: E402
from skyjo_rl_b200.nccl import StatsComm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = 1 << 16
    env = BatchedSkyjoEnv(num_envs=B, num_players=4, seed=1, device=f"cuda:{local}", first_global_env_id=rank * B,
                          auto_reset="next_step")
    env.reset()
    comm = StatsComm(env.device)
    for it in range(8):
        env.step_random(64 + 8 * rank + it)             # ranks differ in their counters
        own = env.stats_tensor()
        lib = env.stats(comm=comm)
        ref = env.stats(all_reduce=True)
        gathered = [torch.empty_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)
        total = torch.stack(gathered).sum(0).tolist()
        diff = {k: (v, t) for (k, v), t in zip(lib.items(), total) if v != t}
        assert lib == ref and not diff, (rank, it, diff, {k: (lib[k], ref[k]) for k in lib if lib[k] != ref[k]})
        side = env.stats_allreduce_async(comm)            # the same collective on the library's side stream
        env.step_random(5)                               # queued behind the snapshot, must not leak into it
        env.stats_allreduce_wait()
        assert side.tolist() == total, (rank, it, "side-stream all-reduce")
        env.clear_stats()
    env.check()
    comm.close()
    if rank == 0:
        print(f"skyjo_stats_allreduce over {world} ranks: steps {lib['steps']}, episodes {lib['episodes']} == torch.distributed: OK")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
