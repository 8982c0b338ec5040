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


@pytest.mark.parametrize("N,indirect,penalty,mr,rr,B,T", [
    (2, False, 2.0, 1.0, 0.0, 130, 260),
    (4, False, 2.0, 1.0, 0.01, 140, 420),
    (3, True, 1.5, 0.0, 0.0, 70, 330),
    (8, False, 2.0, 1.0, 0.0, 40, 620),
])
def test_hostsim_next_step_reset_is_phase_locked_and_matches_oracle(N, indirect, penalty, mr, rr, B, T):
    # SKYJO_RESET_NEXT_STEP: terminal observation in the ending step, reset in the env's next slot
    rng_rollout(HostSimEnv, N, indirect, penalty, mr, rr, B, T, reset_mode=2)


def test_hostsim_next_step_reset_keeps_the_lock_after_illegal_and_truncated_ends():
    B, N = 64, 3
    env = HostSimEnv(num_envs=B, num_players=N, seed=21, auto_reset=2, max_episode_steps=31)
    env.reset()
    rng = np.random.default_rng(0)
    ends = {1: 0, 2: 0, 3: 0}
    waiting = np.zeros(B, bool)         # envs showing a terminal observation, reset slot ahead
    for t in range(400):
        m = env.action_mask
        live = m[~waiting, 24]
        assert np.all(live == live[0]), f"phase lock broken at step {t}"
        draw_slot = live[0] == 1
        a = np.array([rng.choice(np.flatnonzero(m[i])) for i in range(B)], dtype=np.int32)
        bad = rng.random(B) < 0.03
        a[bad] = np.where(m[bad, 24] == 1, 5, 24)   # an action of the other phase: illegal
        env.step(a)
        for c in ends:
            ends[c] += int((env.done_code == c).sum())
        # an end in a draw slot waits for the next slot; an end in a place slot is replaced at once
        waiting = (env.done_code != 0) & draw_slot
    assert ends[2] > 0 and ends[3] > 0
    assert env.stats()["illegal"] == ends[2] and env.stats()["truncated"] == ends[3]


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


@pytest.mark.parametrize("N,indirect,mode,max_steps,B,T", [
    (2, False, 1, 0, 48, 260), (3, False, 2, 0, 48, 330), (3, True, 0, 0, 40, 200),
    (4, False, 2, 45, 40, 300), (5, False, 1, 60, 32, 300), (2, True, 2, 33, 48, 260),
])
def test_hostsim_external_actions_match_the_oracle_model(N, indirect, mode, max_steps, B, T):
    # random external actions with 2 % illegal ones, all three reset modes, truncation on and off
    from parity_util import external_actions_rollout
    seen, st = external_actions_rollout(HostSimEnv, N, indirect, mode, max_steps, B, T)
    assert seen[2] > 0 and st["illegal"] > 0
    if max_steps:
        assert seen[3] > 0


@pytest.mark.parametrize("N,B,T,mode", [(1, 1, 120, 2), (2, 31, 200, 2), (12, 3, 700, 2), (7, 20, 450, 2)])
def test_hostsim_small_batches_in_next_step_mode(N, B, T, mode):
    rng_rollout(HostSimEnv, N, False, 2.0, 1.0, 0.0, B, T, reset_mode=mode)


@pytest.mark.parametrize("N,indirect,B,chunks,mode", [
    (4, False, 96, [64, 7, 13, 64, 2, 64, 64, 33], 2), (2, True, 70, [5, 64, 64, 9, 64], 1),
    (8, False, 40, [64] * 6 + [3, 64, 64], 2),
])
def test_hostsim_chunked_rollout_driver(N, indirect, B, chunks, mode):
    # the chunk-boundary / sampled-env checker the GPU tests run at 2^20 and 2^24 envs (tests/test_gpu_scale_parity.py),
    # here on the host-compiled kernels with a strided sample of the batch replayed on the oracle
    from parity_util import chunked_rollout, sample_blocks
    ids = sample_blocks(B, n_blocks=3, width=8, ranges=4)
    assert 0 in ids and B - 1 in ids and len(ids) < B
    steps, ended, _ = chunked_rollout(HostSimEnv, N, indirect, 2.0, 1.0, 0.01, B, chunks, reset_mode=mode, ids=ids,
                                      first_env=5000)
    assert ended > 0 and steps > 0
    steps, ended, st = chunked_rollout(HostSimEnv, N, indirect, 2.0, 1.0, 0.01, 33, chunks[:4], reset_mode=mode)
    assert st["steps"] == steps
