"""Games played by STRATEGIES instead of the uniform legal policy, to reach the states random play almost never
visits at small player counts (bench stats: 0 reshuffles in 3e7 four-player episodes):

  * "hoarder": always draws from the draw pile and swaps the card into a slot that is already open, so no hidden
    card is ever revealed, the game cannot end and the draw pile runs dry again and again -- many in-game reshuffles
    per episode (skyjo.py:361-365) at N = 2..4, in both observation modes, with the phantom [0,0,0] entries that
    column removals left in the discard pile (skyjo.py:451-460) coming back as drawable cards;
  * "hunter": prefers the swap that completes a column of three equal open cards (removals), takes the discard
    when it matches an open card;
  * "closer": reveals hidden cards as fast as possible to end the game.
A game runs hoarder + hunter for a while, then the closer finishes it.

Three comparisons of the same games, every observation, mask and the final float64 rewards bit for bit:
  live reference (imported from /root/reference, build container only) vs the C oracle;
  the host-compiled kernels vs the C oracle (here); the GPU vs the C oracle (tests/test_gpu_parity.py twin).
"""
import importlib.util
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def strategy_action(obs, mask, rng, mode, own_cards):
    """obs / mask as published for the agent on turn; own_cards = offset of the agent's own 12 slots in obs."""
    legal = np.flatnonzero(mask)
    if mask[24]:                                       # draw phase
        if mode == "hoard":
            return 24
        top = int(obs[17])
        own = obs[own_cards:own_cards + 12]
        if mode == "hunt" and top in own.tolist():
            return 25
        return int(rng.choice(legal))
    hand = int(obs[18])
    own = obs[own_cards:own_cards + 12]
    open_slots = [s for s in range(12) if mask[s] and not mask[12 + s]]      # swappable and already revealed
    if mode in ("hoard", "hunt"):
        for s in open_slots:                           # complete a column of three equal open cards
            col = [int(own[3 * (s // 3) + j]) for j in range(3)]
            col[s % 3] = hand
            if col[0] == col[1] == col[2] and 15 not in col:
                return s
        if mode == "hunt":                             # build columns: put the hand card next to an equal open card
            for s in range(12):
                col = [int(own[3 * (s // 3) + j]) for j in range(3)]
                if mask[s] and col[s % 3] != hand and hand in col:
                    return s
        if open_slots:
            return int(rng.choice(open_slots))
    if mode == "close":
        hidden = [a for a in legal if a >= 12]
        if hidden:
            return int(hidden[0])
    return int(rng.choice(legal))


def game_plan(rng, N):
    """(steps of hoarding / hunting before the closer takes over)"""
    return int(rng.integers(300, 900)) * N // 2


def play_against_oracle(make_env, N, indirect, penalty, mr, rr, seed, env_id, rng):
    env = make_env(num_envs=1, num_players=N, score_penalty=penalty, observe_other_player_indirect=indirect,
                   mean_reward=mr, reward_refunded=rr, seed=seed, auto_reset=False, first_global_env_id=env_id)
    env.reset()
    g = O.OracleGame(N, penalty, indirect)
    g.reset_rng(seed, env_id, 0)
    hoard_for = game_plan(rng, N)
    to_np = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)  # noqa: E731
    t = 0
    while True:
        pid = g.expected_action[0]
        o, m = g.collect_observation(pid)
        np.testing.assert_array_equal(to_np(env.observations)[0], o, err_msg=f"obs at {t}")
        np.testing.assert_array_equal(to_np(env.action_mask)[0], m, err_msg=f"mask at {t}")
        assert int(to_np(env.agent_selection)[0]) == pid
        mode = ("hoard" if t % 5 else "hunt") if t < hoard_for else "close"
        a = strategy_action(o, m, rng, mode, 19 if indirect else 19 + 12 * pid)
        over = g.act(pid, a)
        env.step(np.array([a], dtype=np.uint8))
        t += 1
        assert int(to_np(env.done_code)[0]) == (1 if over else 0)
        if over:
            break
        assert t < 5000 * N
    assert to_np(env.rewards)[0].tobytes() == g.final_rewards(mr, rr).tobytes()
    assert to_np(env.final_scores)[0].tobytes() == np.array(g.game_metrics["final_score"], np.float64).tobytes()
    env.check()
    return t, g.n_reshuffles, int(np.sum(g.game_metrics["num_refunded"]))


STRATEGY_CONFIGS = [(2, False, 2.0, 1.0, 0.0), (3, True, 2.0, 1.0, 0.001), (4, False, 1.5, 0.0, 0.01)]


def run_strategy_games(make_env, games=2):
    steps = resh = refunds = 0
    for ci, (N, indirect, penalty, mr, rr) in enumerate(STRATEGY_CONFIGS):
        rng = np.random.default_rng(500 + ci)
        for gi in range(games):
            t, r, f = play_against_oracle(make_env, N, indirect, penalty, mr, rr, 4242 + ci, gi, rng)
            steps, resh, refunds = steps + t, resh + r, refunds + f
    return steps, resh, refunds


def test_hostsim_matches_oracle_on_strategy_games():
    from hostsim.sim import HostSimEnv
    steps, resh, refunds = run_strategy_games(HostSimEnv)
    assert resh >= 10 and refunds >= 3 and steps > 3000, (steps, resh, refunds)


def play_reference_vs_oracle(mg, N, indirect, penalty, mr, rr, deck, flips, seed, env_id, rng):
    """One strategy-played game through the LIVE reference (dealt the injected deck with the recipe of make_golden.play,
    its in-game reshuffles keyed like ours) and the oracle side by side; returns (steps, reshuffles, removals)."""
    import types
    from rlskyjo.environment.skyjo_env import SimpleSkyjoEnv
    from rlskyjo.game.skyjo import SkyjoGame
    mg._RESHUFFLE["ctx"] = None
    g = SkyjoGame(num_players=N, score_penalty=penalty, observe_other_player_indirect=indirect)
    g.players_cards = deck[: 12 * N].reshape(N, 12).astype(np.int8).copy()
    masks = np.full((N, 12), 2, dtype=np.int8)
    for p in range(N):
        masks[p, flips[p, 0]] = masks[p, flips[p, 1]] = 1
    g.players_masked = masks
    rest = [int(x) for x in deck[12 * N:]]
    g.discard_pile, g.drawpile = [rest[-1]], rest[:-1]
    g._reset_start_player()
    mg._RESHUFFLE["ctx"] = {"seed": seed, "env": env_id, "episode": 0, "q": 0, "n": 0}
    og = O.OracleGame(N, penalty, indirect)
    og.reset_injected(deck, flips)
    og.set_rng_reshuffle(seed, env_id, 0)
    hoard_for = game_plan(rng, N)
    t = 0
    while not g.is_terminated:
        pid = g.expected_action[0]
        assert og.expected_action[0] == pid
        obs, mask = g.collect_observation(pid)
        oo, om = og.collect_observation(pid)
        np.testing.assert_array_equal(oo, obs, err_msg=f"obs at {t}")
        np.testing.assert_array_equal(om, mask, err_msg=f"mask at {t}")
        mode = ("hoard" if t % 5 else "hunt") if t < hoard_for else "close"
        a = strategy_action(obs, mask, rng, mode, 19 if indirect else 19 + 12 * pid)
        assert bool(g.act(pid, a)) == bool(og.act(pid, a))
        t += 1
    metrics = g.get_game_metrics()
    reward = SimpleSkyjoEnv._calc_final_rewards(types.SimpleNamespace(mean_reward=mr, reward_refunded=rr), **metrics)
    assert np.asarray(reward, np.float64).tobytes() == og.final_rewards(mr, rr).tobytes()
    assert np.asarray(metrics["final_score"], np.float64).tobytes() == \
        np.asarray(og.game_metrics["final_score"], np.float64).tobytes()
    np.testing.assert_array_equal(og.players_cards, g.players_cards)
    np.testing.assert_array_equal(og.players_masked, g.players_masked)
    assert og.n_reshuffles == mg._RESHUFFLE["ctx"]["n"]
    return t, og.n_reshuffles, int(np.sum(metrics["num_refunded"]))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rlskyjo")) or importlib.util.find_spec("numba") is None,
                    reason="the reference is only mounted in the build container")
def test_oracle_matches_live_reference_on_strategy_games():
    spec = importlib.util.spec_from_file_location("make_golden_strategy", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    total_resh = 0
    for ci, (N, indirect, penalty, mr, rr) in enumerate(STRATEGY_CONFIGS):
        rng = np.random.default_rng(900 + ci)
        deck = mg.make_deck(rng, "dense" if ci == 2 else "standard")
        flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
        _, resh, _ = play_reference_vs_oracle(mg, N, indirect, penalty, mr, rr, deck, flips, 31337 + ci, ci, rng)
        total_resh += resh
    assert total_resh >= 5, total_resh
