#!/usr/bin/env python
"""Development aid: device time per step of the EXTERNAL-action path (sample_actions kernel + skyjo_step with its
per-step refill deal) next to step_random, on one GPU.   python tools/ext_bench.py [--players 4] [--envs 1048576]"""
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
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--reset", default="same_step")
a = ap.parse_args()
env = BatchedSkyjoEnv(num_envs=a.envs, num_players=a.players, seed=0, auto_reset=a.reset)
env.reset()
env.step_random(640)
logits = torch.zeros((a.envs, 26), device="cuda")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = {}
for name in ("sample+step", "step_only", "step_random"):
    act, _ = env.sample_actions(logits)
    torch.cuda.synchronize()
    if name == "step_only":
        # legal for every env in a draw slot only when phases are locked: use the take-discard/draw pair via mask
        pass
    ev0.record()
    if name == "sample+step":
        for _ in range(a.steps):
            act, _ = env.sample_actions(logits)
            env.step(act)
    elif name == "step_only":
        for _ in range(a.steps):
            env.sample_actions(logits, actions=act)
        ev1.record()
        torch.cuda.synchronize()
        out["sample_only_us"] = round(ev0.elapsed_time(ev1) * 1e3 / a.steps, 2)
        continue
    else:
        env.step_random(a.steps)
    ev1.record()
    torch.cuda.synchronize()
    out[name + "_us"] = round(ev0.elapsed_time(ev1) * 1e3 / a.steps, 2)
env.check()
print(json.dumps({"N": a.players, "B": a.envs, "reset": a.reset, **out, "illegal": env.stats()["illegal"]}))
