"""World-size-2 test of the multi-GPU host logic on CPU (gloo): env sharding by global env id and
the statistics all-reduce (the only collective of the env, never on the step path).  Each rank
drives its shard with the host-compiled device functions (tests/hostsim); rank 0 checks that the
all-reduced statistics and the gathered observations equal those of the unsharded batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

B, N, T, SEED = 96, 4, 150, 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    from hostsim.sim import HostSimEnv
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = B // world
    env = HostSimEnv(num_envs=per, num_players=N, seed=SEED, first_global_env_id=rank * per)
    env.reset()
    env.step_random(T)
    vec = torch.tensor([env.stats()[k] for k in sorted(env.stats())], dtype=torch.int64)
    dist.all_reduce(vec)                                   # BatchedSkyjoEnv.stats(all_reduce=True)
    obs = torch.from_numpy(env.observations.copy())
    gathered = [torch.empty_like(obs) for _ in range(world)]
    dist.all_gather(gathered, obs)
    if rank == 0:
        np.save(out + ".stats.npy", vec.numpy())
        np.save(out + ".obs.npy", torch.cat(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_envs_and_stats_allreduce(tmp_path):
    from hostsim.sim import HostSimEnv, build
    build()                                                # compile once before the ranks race for it
    out = str(tmp_path / "r0")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    whole = HostSimEnv(num_envs=B, num_players=N, seed=SEED)
    whole.reset()
    whole.step_random(T)
    st = whole.stats()
    np.testing.assert_array_equal(np.load(out + ".stats.npy"), np.array([st[k] for k in sorted(st)]))
    np.testing.assert_array_equal(np.load(out + ".obs.npy"), whole.observations)
    assert st["steps"] == B * T and st["episodes"] > 0
