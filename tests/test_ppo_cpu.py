"""CPU checks of the learner-side host logic: turn-based GAE against a naive
per-env, per-seat restatement on synthetic trajectories."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from skyjo_rl_b200.ppo import gae_turn_based  # noqa: E402


def naive_gae(value, agent, done, reward, gamma, lam):
    T, B = done.shape
    N = reward.shape[2]
    adv = np.zeros((T, B), np.float64)
    valid = np.zeros((T, B), bool)
    for b in range(B):
        # split into episodes; within an episode build each seat's list of decision times
        start = 0
        bounds = [t for t in range(T) if done[t, b] != 0]
        segs = []
        for tb in bounds:
            segs.append((start, tb, True))
            start = tb + 1
        if start < T:
            segs.append((start, T - 1, False))
        for (s, e, ended) in segs:
            for q in range(N):
                ts = [t for t in range(s, e + 1) if agent[t, b] == q]
                if not ts:
                    continue
                # successor values / rewards
                nxt_v, nxt_a, known = 0.0, 0.0, ended
                if not ended and agent[T, b] == q:
                    nxt_v, known = float(value[T, b]), True
                pending = float(reward[e, b, q]) if ended else 0.0
                for t in reversed(ts):
                    if known:
                        delta = pending + gamma * nxt_v - float(value[t, b])
                        a = delta + gamma * lam * nxt_a
                        adv[t, b], valid[t, b] = a, True
                    else:
                        a = 0.0
                    pending = 0.0
                    nxt_v, nxt_a, known = float(value[t, b]), a, True
    return adv, valid


@pytest.mark.parametrize("N,lam", [(2, 1.0), (4, 0.9), (3, 0.5)])
def test_gae_turn_based_matches_naive(N, lam):
    rng = np.random.default_rng(N)
    T, B, gamma = 60, 17, 0.97
    agent = np.zeros((T + 1, B), np.int64)
    done = np.zeros((T, B), np.uint8)
    reward = np.zeros((T, B, N), np.float64)
    for b in range(B):
        cur, phase = int(rng.integers(N)), 0
        for t in range(T):
            agent[t, b] = cur
            if rng.random() < 0.06:                       # episode ends on this step
                done[t, b] = 1 + int(rng.integers(3))
                reward[t, b] = rng.normal(size=N)
                cur, phase = int(rng.integers(N)), 0      # new episode, new starter
            elif phase == 1:
                cur, phase = (cur + 1) % N, 0             # place -> next seat draws
            else:
                phase = 1                                 # draw -> same seat places
        agent[T, b] = cur
    value = rng.normal(size=(T + 1, B)).astype(np.float32)
    adv, ret, valid = gae_turn_based(torch.from_numpy(value), torch.from_numpy(agent), torch.from_numpy(done),
                                     torch.from_numpy(reward), gamma=gamma, lam=lam)
    exp_adv, exp_valid = naive_gae(value, agent, done, reward, gamma, lam)
    assert np.array_equal(valid.numpy(), exp_valid)
    assert exp_valid.mean() > 0.8
    np.testing.assert_allclose(adv.numpy()[exp_valid], exp_adv[exp_valid], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(ret.numpy(), adv.numpy() + value[:T], rtol=1e-6, atol=1e-6)
    assert np.all(adv.numpy()[~exp_valid] == 0)
