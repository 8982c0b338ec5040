"""Host time to ENQUEUE one bench iteration (skyjo_step_random(64) + the side-stream statistics all-reduce) against
the device time it takes: how far the loop is from being launch-bound.   python tools/enqueue_time.py [--envs N]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200 import BatchedSkyjoEnv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--players", type=int, default=4)
ap.add_argument("--iters", type=int, default=20)
a = ap.parse_args()
env = BatchedSkyjoEnv(num_envs=a.envs, num_players=a.players, seed=0, auto_reset="next_step")
env.reset()
env.step_random(640)
for _ in range(3):            # the call shape of the loop below is captured into a CUDA graph the second time it is seen
    env.step_random(64)
    env.stats_allreduce_async(None)
torch.cuda.synchronize()
enq = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(a.iters):
    t0 = time.perf_counter()
    env.step_random(64)
    env.stats_allreduce_async(None)
    enq.append(time.perf_counter() - t0)
ev1.record()
torch.cuda.synchronize()
dev_ms = ev0.elapsed_time(ev1) / a.iters
enq.sort()
print(f"enqueue per iteration: median {1e3 * enq[len(enq) // 2]:.3f} ms, min {1e3 * enq[0]:.3f}, max {1e3 * enq[-1]:.3f}; "
      f"device {dev_ms:.3f} ms per iteration; launches per iteration {env.launch_count // (a.iters + 13)}")
