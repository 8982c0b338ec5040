"""The fused policy kernel alone (csrc/skyjo_policy.cu) on the observations of a stepped env: time per launch by CUDA
events, or a short run for ncu (`--launches 3`).   python tools/policy_kernel_bench.py [--envs 262144] [--launches 50]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200 import BatchedSkyjoEnv  # noqa: E402
from skyjo_rl_b200.policy import ActionMaskPolicy, FusedPolicy  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 18)
    ap.add_argument("--players", type=int, default=4)
    ap.add_argument("--launches", type=int, default=50)
    a = ap.parse_args()
    env = BatchedSkyjoEnv(num_envs=a.envs, num_players=a.players, seed=0)
    env.reset()
    env.step_random(40)
    torch.manual_seed(0)
    fused = FusedPolicy(ActionMaskPolicy(env.obs_len).to(env.device), env, with_value=False)
    acts = torch.empty(a.envs, dtype=torch.uint8, device=env.device)
    logp = torch.empty(a.envs, dtype=torch.float32, device=env.device)
    for t in range(3):
        fused.sample(t, acts, logp)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for t in range(a.launches):
        fused.sample(t, acts, logp)
    ev1.record()
    torch.cuda.synchronize()
    us = 1e3 * ev0.elapsed_time(ev1) / a.launches
    flop = 2.0 * a.envs * ((96 + 256) * 256 + 256 * 32)
    print(f"policy_kernel: {us:.1f} us per launch of {a.envs} envs, {flop / us / 1e6:.0f} TFLOP/s issued, "
          f"{a.envs / us:.0f} envs/us")


if __name__ == "__main__":
    main()
