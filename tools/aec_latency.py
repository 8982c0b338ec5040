"""Single-env latency of the PettingZoo-style AEC view (skyjo_rl_b200/aec.py) on the GPU: iterations of the reference's
consumer loop (vanilla_env_example.py:14-35: last() + step()) per second, one env, uniform legal policy on the host.
A GPU is the wrong tool for ONE game -- every iteration is a kernel launch plus a few small device-to-host copies --
this number documents the cost of the compatibility view, next to the reference's own per-step cost."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200.aec import env as make_env  # noqa: E402
from skyjo_rl_b200.policy import policy_ra  # noqa: E402

cfg = {"num_players": 3, "score_penalty": 2.0, "observe_other_player_indirect": True, "mean_reward": 1.0,
       "reward_refunded": 0.001}                     # skyjo_env.DEFAULT_CONFIG
e = make_env(**cfg)
rng = np.random.default_rng(0)
its = games = 0
t0 = None
while True:
    e.reset()
    for agent in e.agent_iter(max_iter=2000):
        obs, reward, done, info = e.last()
        e.step(None if done else policy_ra(obs["observations"], obs["action_mask"], rng))
        its += 1
    games += 1
    if games == 3:                                   # warm-up
        t0, its0 = time.perf_counter(), its
    if games > 3 and time.perf_counter() - t0 > 5.0:
        break
dt = time.perf_counter() - t0
print(f"AEC view, one env on the GPU: {(its - its0) / dt:.0f} loop iterations/s ({1e6 * dt / (its - its0):.0f} us each), "
      f"{games - 3} games in {dt:.1f} s")
