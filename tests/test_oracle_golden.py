"""The C oracle (oracle/skyjo_oracle.c) against fixtures produced by the live reference.

CPU-only.  Every step of every recorded game is replayed; obs, mask, acting agent, hand,
discard top, final cards/masks, scores, float64 rewards and metrics must be bit-identical.
"""
import numpy as np
import pytest

from golden_util import Golden, golden_names
from oracle import oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_games(name):
    G = Golden(name)
    steps = 0
    for gi in range(G.games):
        g = G.game(gi)
        og = O.OracleGame(G.N, G.score_penalty, G.indirect)
        og.reset_injected(g["decks"], g["flips"])
        og.set_rng_reshuffle(G.seed, gi, 0)
        for t in range(len(g["action"])):
            pid, phase = og.expected_action
            assert pid == g["agent"][t]
            assert (0 if phase == "draw" else 1) == g["phase"][t]
            obs, mask = og.collect_observation(pid)
            np.testing.assert_array_equal(obs, g["obs"][t])
            np.testing.assert_array_equal(mask, g["mask"][t])
            oo, mo = og.collect_observation((pid + 1) % G.N)
            np.testing.assert_array_equal(oo, g["obs_other"][t])
            np.testing.assert_array_equal(mo, g["mask_other"][t])
            assert og.hand_card == g["hand"][t]
            over = og.act(pid, int(g["action"][t]))
            assert over == (t == len(g["action"]) - 1)
            steps += 1
        assert og.is_terminated
        m = og.game_metrics
        np.testing.assert_array_equal(np.array(m["final_score"]), g["final_score"])
        np.testing.assert_array_equal(np.array(m["num_refunded"]), g["num_refunded"])
        np.testing.assert_array_equal(np.array(m["num_placed"]), g["num_placed"])
        np.testing.assert_array_equal(og.players_cards, g["final_cards"])
        np.testing.assert_array_equal(og.players_masked, g["final_masked"])
        assert og.n_reshuffles == g["n_reshuffles"]
        r = og.final_rewards(G.mean_reward, G.reward_refunded)
        assert r.tobytes() == g["reward"].tobytes(), (r, g["reward"])
        fo, fm = og.collect_observation(og.expected_action[0])
        np.testing.assert_array_equal(fo, g["final_obs"])
        np.testing.assert_array_equal(fm, g["final_mask"])
        assert og.expected_action[0] == g["final_agent"]
    assert steps == int(G.lengths.sum())


def test_evaluate_game_known_answer():
    # /root/reference/notebooks/trainpettingzoo.ipynb:52745-52758 (cell output):
    # table._evaluate_game(table.players_cards, 0, score_penalty=2) -> [76.0, 41.0, 21.0]
    cards = np.array([[-1, 9, 7, -2, 4, 2, 0, 7, 4, 0, 3, 5],
                      [0, 7, 1, 10, 7, 2, 0, 6, 1, -1, -1, 9],
                      [-1, 6, 5, -2, 4, 2, 1, 4, -2, 3, -2, 3]], dtype=np.int8)
    np.testing.assert_array_equal(O.evaluate_game(cards, 0, 2.0), [76.0, 41.0, 21.0])


def test_notebook_episode_result():
    # notebooks/trainpettingzoo.ipynb:2732,2746-2747: Results {0: 46, 1: 118, 2: 87} and rewards
    # 38.66.., -33.33.., -2.33.. with mean_reward 1.0 (finisher 1 had 59, doubled).
    r = O.calc_final_rewards([46.0, 118.0, 87.0], [0, 0, 0], 1.0, 0.0)
    ref = -np.array([46.0, 118.0, 87.0]) + np.mean([46.0, 118.0, 87.0]) + 1.0
    assert r.tobytes() == ref.tobytes()
    assert abs(r[0] - 38.666666666666664) < 1e-12 and abs(r[1] + 33.33333333333333) < 1e-12


def test_numpy_sum_order():
    rng = np.random.default_rng(5)
    for n in range(1, 13):
        for _ in range(300):
            score = rng.random(n) * rng.choice([1.0, 1e3, 1e-3], n)
            refunded = rng.integers(0, 5, n)
            mr, rr = float(rng.normal()), float(rng.choice([0.0, 0.01, 0.37]))
            ref = -score + np.mean(score) + mr
            if rr:
                ref += np.array(refunded) * rr
            got = O.calc_final_rewards(score, refunded, mr, rr)
            assert got.tobytes() == ref.tobytes()


def test_scoring_quirks():
    # equal columns score 0 even when never refunded; ties forgiven;
    # a negative finisher score is multiplied too.
    cards = np.array([[5, 5, 5, 1, 2, 3, -2, -2, -2, 0, 0, 1],
                      [1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3]], dtype=np.int8)
    np.testing.assert_array_equal(O.evaluate_game(cards, 0, 2.0), [7.0, 7.0])   # tie forgiven
    np.testing.assert_array_equal(O.evaluate_game(cards, 1, 2.0), [7.0, 7.0])
    cards2 = np.array([[-2, -2, -1, 0, 0, -1, 1, 1, 1, 2, 2, 2],
                       [-2, -2, -2, -2, -2, -1, -2, -1, -2, 0, 0, 0]], dtype=np.int8)
    # p0 = -5 - 1 = -6, p1 = -5 - 5 = -10; finisher 0 is not the minimum -> -12
    np.testing.assert_array_equal(O.evaluate_game(cards2, 0, 2.0), [-12.0, -10.0])
