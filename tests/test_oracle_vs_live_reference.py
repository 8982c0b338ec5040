"""Differential test of the C oracle against the LIVE reference (rlskyjo imported from /root/reference, numba JIT):
fresh seeded random games on every run, beyond the committed fixtures.  CPU-only, and only where the reference is
mounted (the build container); it is skipped on the GPU box, where /root/reference does not exist.  The games are
generated with the recipe of tests/golden/make_golden.py (injected decks / flips / keyed reshuffle)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(
    not os.path.isdir(os.path.join(REF, "rlskyjo")) or importlib.util.find_spec("numba") is None,
    reason="the reference is only mounted in the build container")


@pytest.fixture(scope="module")
def mg():
    spec = importlib.util.spec_from_file_location(
        "make_golden_live", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)          # imports rlskyjo, patches the reshuffle hook; does not write fixtures
    return m


@pytest.mark.parametrize("N,indirect,kind,games", [(2, False, "standard", 6), (3, True, "dense", 5), (4, False, "dense", 5),
                                                   (6, False, "standard", 4), (9, True, "standard", 3),
                                                   (11, False, "dense", 3)])
def test_oracle_matches_live_reference_on_fresh_games(mg, N, indirect, kind, games):
    rng = np.random.default_rng(20260000 + 31 * N + (7 if indirect else 0))
    penalty = float(rng.choice([1.0, 1.5, 2.0, 3.0]))
    mr, rr = float(rng.choice([0.0, 1.0])), float(rng.choice([0.0, 0.01]))
    seed = int(rng.integers(1, 2**40))
    steps = 0
    for gi in range(games):
        deck = mg.make_deck(rng, kind)
        flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
        g = mg.play(N, indirect, penalty, mr, rr, deck, flips, rng, seed, gi, kind == "dense")
        og = O.OracleGame(N, penalty, indirect)
        og.reset_injected(deck, flips)
        og.set_rng_reshuffle(seed, gi, 0)
        for t in range(len(g["action"])):
            pid = og.expected_action[0]
            assert pid == g["agent"][t]
            obs, mask = og.collect_observation(pid)
            np.testing.assert_array_equal(obs, g["obs"][t])
            np.testing.assert_array_equal(mask, g["mask"][t])
            oo, mo = og.collect_observation((pid + 1) % N)
            np.testing.assert_array_equal(oo, g["obs_other"][t])
            np.testing.assert_array_equal(mo, g["mask_other"][t])
            assert og.act(pid, int(g["action"][t])) == (t == len(g["action"]) - 1)
            steps += 1
        m = og.game_metrics
        assert np.array(m["final_score"]).tobytes() == g["final_score"].tobytes()
        np.testing.assert_array_equal(np.array(m["num_refunded"]), g["num_refunded"])
        np.testing.assert_array_equal(og.players_cards, g["final_cards"])
        np.testing.assert_array_equal(og.players_masked, g["final_masked"])
        assert og.n_reshuffles == g["n_reshuffles"]
        assert og.final_rewards(mr, rr).tobytes() == g["reward"].tobytes()
    assert steps > 0
