"""Whole-episode traces of the reference's PettingZoo env (`skyjo_env.env(**cfg)`, skyjo_env.py:19-26, driven by
the consumer loop of vanilla_env_example.py:14-35) replayed through `skyjo_rl_b200.aec.SkyjoAECView`.

tests/golden/env_trace.npz was recorded by tests/golden/make_env_trace.py from the UNMODIFIED
`SimpleSkyjoEnv` (its own reset / step / observe / _calc_final_rewards) on the pettingzoo / gym stand-ins of
tests/shims: 14 games of 2-5 players, half of them ended by one illegal action
(`TerminateIllegalWrapper(illegal_reward=-1)`).  Compared per iteration of the loop: the agent on turn, the
observation dict (also for done agents: what observe() returns after the game), last()'s cumulative reward
(float64, bit for bit), done, and after the step the next agent, the number of live agents and every remaining
cumulative reward.

  * here: the view on the host-compiled kernels (CPU suite), plus -- in the build container only -- a live
    re-recording on fresh decks compared the same way;
  * GPU twin: tests/test_gpu_policy.py::test_gpu_aec_replays_reference_env_traces.
"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_env_traces():
    z = np.load(os.path.join(HERE, "golden", "env_trace.npz"))
    out = []
    for name in z["names"]:
        pre = str(name) + "/"
        out.append((str(name), {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}))
    return out


def env_kwargs(c):
    return dict(num_players=int(c["num_players"]), score_penalty=float(c["score_penalty"]),
                observe_other_player_indirect=bool(c["observe_other_player_indirect"]),
                mean_reward=float(c["mean_reward"]), reward_refunded=float(c["reward_refunded"]))


def replay_game(aec, deck, flips, rec):
    """rec: dict of per-iteration arrays of ONE game (see make_env_trace.play_env)."""
    N = len(aec.possible_agents)
    aec.reset_injected(deck, flips)
    n = len(rec["agent"])
    t = 0
    for agent in aec.agent_iter(max_iter=n + 5):
        assert t < n, "the view plays longer than the reference"
        obs, reward, done, info = aec.last()
        assert agent == f"player_{rec['agent'][t]}", f"agent at {t}"
        assert done == bool(rec["done"][t]) and info == {}
        assert np.float64(reward).tobytes() == np.float64(rec["reward"][t]).tobytes(), f"reward at {t}"
        np.testing.assert_array_equal(obs["observations"], rec["obs"][t], err_msg=f"obs at {t}")
        np.testing.assert_array_equal(obs["action_mask"], rec["mask"][t], err_msg=f"mask at {t}")
        a = int(rec["action"][t])
        assert (a < 0) == done
        aec.step(None if a < 0 else a)
        assert len(aec.agents) == rec["n_agents"][t]
        if aec.agents:
            assert aec.agent_selection == f"player_{rec['next_agent'][t]}", f"next agent at {t}"
        cum = np.full(N, np.nan)
        for name, r in aec._cumulative_rewards.items():
            cum[int(name.split("_")[-1])] = r
        assert cum.tobytes() == np.asarray(rec["cumulative"][t], np.float64).tobytes(), f"cumulative at {t}"
        t += 1
    assert t == n and not aec.agents


def replay_config(make_aec, c):
    off = 0
    for gi, L in enumerate(c["lengths"]):
        # the recording keyed in-game reshuffles by (seed, env = game index, episode 0)
        aec = make_aec(dict(env_kwargs(c), seed=int(c["seed"]), first_global_env_id=gi))
        rec = {k: c[k][off:off + L] for k in ("agent", "reward", "done", "obs", "mask", "action", "next_agent",
                                              "n_agents", "cumulative")}
        replay_game(aec, c["decks"][gi], c["flips"][gi], rec)
        off += L
    return off


def hostsim_aec(kw):
    from hostsim.sim import HostSimEnv
    from skyjo_rl_b200.aec import SkyjoAECView

    class Backend(HostSimEnv):
        def observation_space(self, agent):
            return None

        def action_space(self, agent):
            return None
    return SkyjoAECView(Backend(num_envs=1, auto_reset=False, **kw), 0)


@pytest.mark.parametrize("name,c", load_env_traces(), ids=[n for n, _ in load_env_traces()])
def test_hostsim_aec_view_replays_reference_env_traces(name, c):
    assert replay_config(hostsim_aec, c) > 50


def test_fixture_holds_both_kinds_of_game_end():
    ends = {"illegal": 0, "played": 0}
    for _, c in load_env_traces():
        off = 0
        for L in c["lengths"]:
            r = c["reward"][off:off + L][c["done"][off:off + L] == 1]
            ends["illegal" if sorted(r.tolist()) == [-1.0] + [0.0] * (len(r) - 1) else "played"] += 1
            off += L
    assert ends["illegal"] >= 6 and ends["played"] >= 6


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rlskyjo")) or importlib.util.find_spec("numba") is None,
                    reason="the reference is only mounted in the build container")
def test_hostsim_aec_view_matches_live_reference_env_on_fresh_games():
    spec = importlib.util.spec_from_file_location("make_env_trace_live", os.path.join(HERE, "golden", "make_env_trace.py"))
    met = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(met)
    mg = met._load_make_golden()
    from rlskyjo.environment import skyjo_env
    rng = np.random.default_rng(20261017)
    for gi in range(6):
        N = int(rng.integers(2, 7))
        cfg = dict(num_players=N, score_penalty=float(rng.choice([1.0, 2.0, 3.0])),
                   observe_other_player_indirect=bool(rng.integers(2)), mean_reward=float(rng.choice([0.0, 1.0])),
                   reward_refunded=float(rng.choice([0.0, 0.01])))
        deck = mg.make_deck(rng, "dense" if gi % 3 == 0 else "standard")
        flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
        illegal_at = int(rng.integers(0, 80)) if gi % 2 else -1
        rec, live = met.play_env(mg, skyjo_env, cfg, deck, flips, rng, 555, gi, illegal_at)
        rec = {k: np.asarray(v) for k, v in rec.items()}
        replay_game(hostsim_aec(dict(cfg, seed=555, first_global_env_id=gi)), deck, flips, rec)
