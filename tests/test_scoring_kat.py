"""End-of-game arithmetic against known answers computed by the UNMODIFIED reference (tests/golden/scoring_kat.npz,
made by tests/golden/make_scoring_kat.py with `SkyjoGame._evaluate_game`, skyjo.py:477-498, and
`SimpleSkyjoEnv._calc_final_rewards`, skyjo_env.py:293-312): 160 whole games on low-card boards built to hit ties for
the best score (penalty forgiven), negative finisher scores that still get multiplied, equal columns, penalties < 1
and fractional.  Replayed through the oracle and the host-compiled kernels here, through the GPU in
tests/test_gpu_parity.py::test_scoring_known_answers_of_the_reference."""
import os

import numpy as np

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load_kat():
    z = np.load(os.path.join(HERE, "golden", "scoring_kat.npz"))
    return {k: z[k] for k in z.files}


def flip_action(mask):
    if mask[24]:
        return 24
    hidden = np.flatnonzero(mask[12:24])
    return 12 + int(hidden[0]) if len(hidden) else int(np.flatnonzero(mask)[0])


def test_oracle_reproduces_the_reference_scores_and_rewards():
    z = load_kat()
    seen = {"tie": 0, "negative_penalised": 0}
    for i in range(len(z["N"])):
        N, pen, mr, rr = int(z["N"][i]), float(z["penalty"][i]), float(z["mr"][i]), float(z["rr"][i])
        g = O.OracleGame(N, pen, False)
        g.reset_injected(z["deck"][i], z["flips"][i][:N])
        steps = 0
        while True:
            pid = g.expected_action[0]
            _, mask = g.collect_observation(pid)
            steps += 1
            if g.act(pid, flip_action(mask)):
                break
        assert pid == z["finisher"][i] and steps == z["steps"][i]
        np.testing.assert_array_equal(g.players_cards, z["cards"][i][:N])
        score = np.array(g.game_metrics["final_score"], np.float64)
        assert score.tobytes() == z["score"][i][:N].tobytes(), i
        assert g.final_rewards(mr, rr).tobytes() == z["reward"][i][:N].tobytes(), i
        assert O.evaluate_game(z["cards"][i][:N], pid, pen).tobytes() == z["score"][i][:N].tobytes()
        raw = O.evaluate_game(z["cards"][i][:N], pid, 1.0)
        if raw[pid] != raw.min():
            seen["negative_penalised"] += raw[pid] < 0
        elif (raw == raw[pid]).sum() > 1:
            seen["tie"] += 1
    assert seen["tie"] >= 3 and seen["negative_penalised"] >= 3, seen


def replay_kat_batched(make_env):
    """All cases of one player count as ONE batch (one env per case) through env.step()."""
    z = load_kat()
    total = 0
    for N in (2, 3, 4):
        for pen in sorted(set(z["penalty"].tolist())):
            for mr in sorted(set(z["mr"].tolist())):
                for rr in sorted(set(z["rr"].tolist())):
                    idx = np.flatnonzero((z["N"] == N) & (z["penalty"] == pen) & (z["mr"] == mr) & (z["rr"] == rr))
                    if len(idx) == 0:
                        continue
                    B = len(idx)
                    env = make_env(num_envs=B, num_players=N, score_penalty=pen, mean_reward=mr, reward_refunded=rr,
                                   auto_reset=False)
                    env.reset_injected(z["deck"][idx], z["flips"][idx][:, :N])
                    to_np = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)  # noqa: E731
                    for t in range(int(z["steps"][idx].max())):
                        mask = to_np(env.action_mask)
                        env.step(np.array([flip_action(m) for m in mask], dtype=np.uint8))
                    assert (to_np(env.done_code) == 1).all()
                    assert to_np(env.final_scores).tobytes() == np.ascontiguousarray(z["score"][idx][:, :N]).tobytes()
                    assert to_np(env.rewards).tobytes() == np.ascontiguousarray(z["reward"][idx][:, :N]).tobytes()
                    env.check()
                    total += B
    assert total == len(z["N"])
    return total


def test_hostsim_reproduces_the_reference_scores_and_rewards():
    from hostsim.sim import HostSimEnv
    assert replay_kat_batched(HostSimEnv) == 160
