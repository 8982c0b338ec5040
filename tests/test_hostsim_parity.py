"""CPU parity of the CUDA library's per-env device functions, compiled for the host by
tests/hostsim (g++ on the same .cuh sources the kernels inline), against the reference-generated
golden fixtures and the C oracle.  This is the no-GPU mirror of tests/test_gpu_parity.py: it pins
the transition / scoring / auto-reset / observation-stream / tile-staging logic; the GPU tests pin
the kernels that run it."""
import numpy as np
import pytest

from golden_util import golden_names
from hostsim.sim import HostSimEnv
from parity_util import golden_replay, rng_rollout


@pytest.mark.parametrize("name", golden_names())
def test_hostsim_golden_replay(name):
    golden_replay(HostSimEnv, name)


@pytest.mark.parametrize("N,indirect,penalty,mr,rr,B,T", [
    (1, False, 2.0, 1.0, 0.0, 40, 120),
    (2, False, 2.0, 1.0, 0.0, 130, 200),
    (3, True, 2.0, 1.0, 0.001, 70, 260),
    (4, False, 2.0, 1.0, 0.0, 140, 330),
    (4, True, 1.5, 0.0, 0.01, 64, 330),
    (8, False, 2.0, 1.0, 0.0, 48, 560),
    (12, False, 1.1, 1.0, 0.01, 33, 760),
    (12, True, 2.0, 1.0, 0.0, 33, 760),
])
def test_hostsim_rng_rollout_matches_oracle(N, indirect, penalty, mr, rr, B, T):
    st = rng_rollout(HostSimEnv, N, indirect, penalty, mr, rr, B, T)
    if N >= 8:
        assert st["reshuffles"] > 0      # the lazy draw-pile path is exercised


def test_hostsim_illegal_and_truncation():
    B, N = 96, 3
    env = HostSimEnv(num_envs=B, num_players=N, seed=5, auto_reset=True)
    env.reset()
    agent0 = env.agent_selection.copy()
    actions = np.full(B, 24, dtype=np.int32)
    actions[::2] = 3
    actions[1::4] = 26
    env.step(actions)
    bad = np.zeros(B, bool)
    bad[::2] = True
    bad[1::4] = True
    assert np.all(env.done_code[bad] == 2) and np.all(env.done_code[~bad] == 0)
    for i in np.flatnonzero(bad):
        exp = np.zeros(N)
        exp[agent0[i]] = -1.0
        np.testing.assert_array_equal(env.rewards[i], exp)
    env.step_random(1)
    assert np.all(env.rewards == 0)
    env2 = HostSimEnv(num_envs=B, num_players=2, seed=9, auto_reset=True, max_episode_steps=20)
    env2.reset()
    trunc = 0
    for _ in range(60):
        env2.step_random(1)
        trunc += int((env2.done_code == 3).sum())
    assert trunc == 3 * B and env2.stats()["truncated"] == 3 * B
