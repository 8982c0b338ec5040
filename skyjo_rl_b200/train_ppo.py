"""Self-play PPO on the batched env -- the B200 counterpart of the reference's training script
(rlskyjo/models/train_model_simple_rllib.py:34-137: one shared action-mask policy for all seats,
PPO, periodic checkpoints).

    python -m skyjo_rl_b200.train_ppo --iters 50 [--envs 16384] [--players 3] [--checkpoint out.pt]
    python -m torch.distributed.run --nproc-per-node N -m skyjo_rl_b200.train_ppo ...   # one rank per GPU

Every rank owns a disjoint range of global env ids; gradients are averaged with an NCCL
all-reduce, the episode statistics with the env's own int64[32] all-reduce.
"""
import argparse
import json
import os
import time

import torch

from . import BatchedSkyjoEnv, DEFAULT_CONFIG
from .ppo import PPOTrainer


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--envs", type=int, default=16384, help="envs per GPU")
    ap.add_argument("--players", type=int, default=DEFAULT_CONFIG["num_players"])
    ap.add_argument("--direct", action="store_true", help="observe_other_player_indirect=False")
    ap.add_argument("--rollout-len", type=int, default=64)
    ap.add_argument("--lr", type=float, default=3e-4)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--minibatches", type=int, default=8)
    ap.add_argument("--ent-coef", type=float, default=0.01)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--fused", action="store_true",
                    help="rollouts through the library's fused tensor-core policy kernel (csrc/skyjo_policy.cu)")
    ap.add_argument("--tf32", action="store_true",
                    help="let the learner's fp32 GEMMs (ATen) use TF32 tensor cores; the env path is integer and unaffected")
    a = ap.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    if a.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(a.seed)                      # identical initial weights on every rank
    cfg = dict(DEFAULT_CONFIG, num_players=a.players, observe_other_player_indirect=not a.direct)
    env = BatchedSkyjoEnv(num_envs=a.envs, device=dev, seed=a.seed, first_global_env_id=rank * a.envs, **cfg)
    env.reset()
    tr = PPOTrainer(env, rollout_len=a.rollout_len, lr=a.lr, epochs=a.epochs, minibatches=a.minibatches,
                    ent_coef=a.ent_coef, seed=a.seed + rank, fused=a.fused)
    def rank_path(path):      # env state and generator are per rank (disjoint global env ids): one file per rank
        return path if world == 1 else f"{path}.rank{rank}"

    if a.resume:
        tr.load_state_dict(torch.load(rank_path(a.resume), map_location=dev, weights_only=False))
    for _ in range(a.iters):
        t0 = time.perf_counter()
        m = tr.train_iteration()
        torch.cuda.synchronize(dev)
        m["seconds"] = time.perf_counter() - t0
        m["env_steps_per_s"] = m["env_steps"] * world / m["seconds"]
        if rank == 0:
            print(json.dumps({k: (round(v, 5) if isinstance(v, float) else v) for k, v in m.items()}), flush=True)
    if a.checkpoint:
        torch.save(tr.state_dict(), rank_path(a.checkpoint))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
