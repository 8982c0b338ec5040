"""Small invocations of every device path added in round 2, for compute-sanitizer (tools/gpu_sanitize.sh runs
smoke() the same way).  No oracle here: the point is memcheck / racecheck / synccheck coverage, the results are
compared with the device buffers where that is free.

    compute-sanitizer --tool memcheck python tools/sanitize_paths.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyjo_rl_b200 import BatchedSkyjoEnv  # noqa: E402
from skyjo_rl_b200.policy import ActionMaskPolicy, FusedPolicy  # noqa: E402


def host_buffers(env):
    B, D, N = env.num_envs, env.obs_len, env.num_players
    pin = lambda *s, dt: torch.empty(s, dtype=dt).pin_memory()  # noqa: E731
    return (pin(B, D, dt=torch.int8), pin(B, 26, dt=torch.int8), pin(B, dt=torch.int8), pin(B, dt=torch.uint8),
            pin(B, N, dt=torch.float64))


def legal_actions(env):
    return torch.multinomial(env.action_mask.float(), 1).squeeze(1).to(torch.uint8).cpu().pin_memory()


def step_host_modes():
    # ragged batch (not a multiple of the 128-env tile), three wire modes, forced chunking
    os.environ["SKYJO_HOST_CHUNKS"] = "4"
    for N, ind in ((4, False), (2, True), (8, False)):
        env = BatchedSkyjoEnv(num_envs=1000, num_players=N, seed=3, device="cuda:0", observe_other_player_indirect=ind)
        env.reset()
        env.step_random(40)
        obs, mask, agent, done, rew = host_buffers(env)
        for mode in ("raw", "compact", "mixed"):
            env.set_host_wire(mode)
            for _ in range(3):
                env.step_host(legal_actions(env), obs, mask, agent, done, rew)
                assert torch.equal(obs, env.observations.cpu()) and torch.equal(mask, env.action_mask.cpu()), (N, mode)
        env.check()
    del os.environ["SKYJO_HOST_CHUNKS"]
    print("step_host wire modes ok")


def step_host_small():
    for B in (1, 7, 256):
        env = BatchedSkyjoEnv(num_envs=B, num_players=3, seed=5, device="cuda:0")
        env.reset()
        obs, mask, agent, done, rew = host_buffers(env)
        for _ in range(40):
            env.step_host(legal_actions(env), obs, mask, agent, done, rew)
            assert torch.equal(obs, env.observations.cpu()) and torch.equal(agent, env.agent_selection.cpu()), B
        env.check()
    print("step_host small-batch path ok")


def ranges_and_stats():
    env = BatchedSkyjoEnv(num_envs=3000, num_players=4, seed=9, device="cuda:0", auto_reset="next_step")
    env.reset()
    env.set_env_ranges(4)
    for _ in range(3):
        env.step_random(70)          # two refill windows and a remainder per call, on four streams (direct launches)
        out = env.stats_allreduce_async(None)
    for _ in range(4):
        env.step_random(64)          # an even number of windows: captured into a CUDA graph on the second call, replayed after
        out = env.stats_allreduce_async(None)
    env.step_random(5)               # direct launches in between: the device-side lockstep counter is brought up to date
    env.step_random(64)
    env.stats_allreduce_wait()
    torch.cuda.synchronize()
    assert int(out[0]) >= 0
    env.check()
    ro = env.rollout_random(40)
    assert ro["observations"].shape[0] == 40
    env.check()
    print("env ranges, side-stream statistics, rollout kernel ok")


def policy_kernel():
    for N, ind, B in ((4, False, 1000), (6, False, 300), (8, True, 129)):
        env = BatchedSkyjoEnv(num_envs=B, num_players=N, seed=11, device="cuda:0", observe_other_player_indirect=ind)
        env.reset()
        env.step_random(30)
        torch.manual_seed(0)
        fused = FusedPolicy(ActionMaskPolicy(env.obs_len).to(env.device), env)
        for t in range(4):
            acts, logp = fused.sample(t)
            assert bool((env.action_mask.gather(1, acts.long().unsqueeze(1)) == 1).all())
            v = fused.value()
            assert bool(torch.isfinite(logp).all()) and bool(torch.isfinite(v).all())
            env.step(acts)
        p1, p2, lg = fused.debug()
        assert bool(torch.isfinite(lg).all())
        env.check()
    print("policy kernel ok")


if __name__ == "__main__":
    which = sys.argv[1:] or ["host", "small", "ranges", "policy"]
    if "host" in which:
        step_host_modes()
    if "small" in which:
        step_host_small()
    if "ranges" in which:
        ranges_and_stats()
    if "policy" in which:
        policy_kernel()
    torch.cuda.synchronize()
    print("done")
