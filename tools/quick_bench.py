#!/usr/bin/env python
"""Quick device-time probe of the fused step kernel (development aid, not the bench contract).

    python tools/quick_bench.py [--players 4] [--envs 1048576] [--steps 512] [--indirect] [--tag x]
Prints one line: mean step-kernel / deal-kernel device time from per-launch CUDA events."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from skyjo_rl_b200 import BatchedSkyjoEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--players", type=int, default=4)
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--steps", type=int, default=512)
ap.add_argument("--preroll", type=int, default=640)
ap.add_argument("--indirect", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--rollout", type=int, default=0, help="time rollout_random with this many steps per call")
ap.add_argument("--reset", default="same_step", help="same_step | next_step (phase-locked)")
a = ap.parse_args()
env = BatchedSkyjoEnv(num_envs=a.envs, num_players=a.players, observe_other_player_indirect=a.indirect, seed=0,
                      auto_reset=a.reset)
env.reset()
env.step_random(a.preroll)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = env.stats()["steps"]
ev0.record()
env.step_random(a.steps)
ev1.record()
torch.cuda.synchronize()
wall = ev0.elapsed_time(ev1) * 1e3 / a.steps
counted = env.stats()["steps"] - n0      # env-steps actually played (reset slots of next_step mode excluded)
prof = env.step_random_profile(a.steps)
env.check()
if a.rollout:
    T = a.rollout
    ro = env.rollout_random(T)
    torch.cuda.synchronize()
    r0 = env.stats()["steps"]
    ev0.record()
    reps = max(1, a.steps // T)
    for _ in range(reps):
        env.rollout_random(T, ro)
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / (reps * T)
    rcounted = env.stats()["steps"] - r0
    env.profile_begin()
    for _ in range(reps):
        env.rollout_random(T, ro)
    pr = env.profile_end()
    env.check()
    print(json.dumps({"tag": a.tag + " rollout", "N": a.players, "B": a.envs, "T": T, "us_per_step_all": round(us, 2),
                      "rollout_kernel_us_per_step": round(1e3 * pr["step_ms"] / (reps * T), 2),
                      "deal_kernel_us": round(1e3 * pr["deal_ms"] / max(pr["deal_launches"], 1), 2),
                      "steps_per_s": round(rcounted / (us * 1e-6 * reps * T), 0)}))
print(json.dumps({"tag": a.tag, "reset": a.reset, "N": a.players, "B": a.envs, "indirect": a.indirect,
                  "us_per_step_all": round(wall, 2),
                  "step_kernel_us": round(1e3 * prof["step_ms"] / prof["step_launches"], 2),
                  "deal_kernel_us": round(1e3 * prof["deal_ms"] / max(prof["deal_launches"], 1), 2),
                  "counted_frac": round(counted / (a.envs * a.steps), 5),
                  "steps_per_s": round(counted / (wall * 1e-6 * a.steps), 0)}))
