"""Generate tests/golden/env_trace.npz: whole-episode traces of the reference's PettingZoo env,
`rlskyjo.environment.skyjo_env.env(**cfg)` (skyjo_env.py:19-26), driven by the reference's consumer
loop (`vanilla_env_example.py:14-35`: agent_iter / last / step(action | None)).

Run here (the reference does not exist on the GPU box):   python tests/golden/make_env_trace.py

What is the reference's and what is ours:
  * `SimpleSkyjoEnv.reset / step / observe / _calc_final_rewards` and the whole `SkyjoGame` under it are
    the UNMODIFIED reference, imported from /root/reference;
  * gym / pettingzoo are not installed and cannot be (no network): `AECEnv` and the four wrappers come
    from tests/shims, a restatement of pettingzoo 1.14.0 -- so the reward visibility / dead-step order
    recorded here is pinned by the reference's own `step()` code running on restated AEC helpers (the
    notebook trace, tests/golden/notebook_trace.npz, is the recording made with the real package);
  * inputs are injected as in make_golden.py: deck order, flipped slots, seeded legal actions, and in
    about half of the games ONE illegal action (mask == 0) at a random point, which
    `TerminateIllegalWrapper(illegal_reward=-1)` turns into the end of the game.
Each record is one iteration of the consumer loop: what last() returned, the action stepped
(-1 = None for a done agent), and the AEC attributes after the step.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_make_golden():
    spec = importlib.util.spec_from_file_location("make_golden_env", os.path.join(HERE, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)  # imports rlskyjo over tests/shims, patches the keyed reshuffle
    return m


CONFIGS = [
    # name, env kwargs, deck kind, games
    ("n2_direct", dict(num_players=2, score_penalty=2.0, observe_other_player_indirect=False, mean_reward=1.0,
                       reward_refunded=0.0), "standard", 4),
    ("n3_default", dict(num_players=3, score_penalty=2.0, observe_other_player_indirect=True, mean_reward=1.0,
                        reward_refunded=0.001), "standard", 4),      # skyjo_env.DEFAULT_CONFIG
    ("n4_direct_dense", dict(num_players=4, score_penalty=1.5, observe_other_player_indirect=False, mean_reward=0.0,
                             reward_refunded=0.01), "dense", 4),
    ("n5_indirect", dict(num_players=5, score_penalty=2.0, observe_other_player_indirect=True, mean_reward=-1.0,
                         reward_refunded=0.0), "standard", 2),
]


def play_env(mg, skyjo_env, cfg, deck, flips, rng, seed, env_id, illegal_at):
    N = cfg["num_players"]
    e = skyjo_env.env(**cfg)
    table = e.unwrapped.table
    ref_reset = table.reset

    def injected_reset():
        mg._RESHUFFLE["ctx"] = None
        ref_reset()
        table.players_cards = deck[: 12 * N].reshape(N, 12).astype(np.int8).copy()
        masks = np.full((N, 12), 2, dtype=np.int8)
        for p in range(N):
            masks[p, flips[p, 0]] = 1
            masks[p, flips[p, 1]] = 1
        table.players_masked = masks
        rest = [int(x) for x in deck[12 * N:]]
        table.discard_pile, table.drawpile = [rest[-1]], rest[:-1]
        table._reset_start_player()
        mg._RESHUFFLE["ctx"] = {"seed": seed, "env": env_id, "episode": 0, "q": 0, "n": 0}

    table.reset = injected_reset
    e.reset()
    pid_of = lambda a: int(a.split("_")[-1])  # noqa: E731
    rec = {k: [] for k in ("agent", "reward", "done", "obs", "mask", "action", "next_agent", "n_agents", "cumulative")}
    live = 0
    for agent in e.agent_iter(max_iter=400 * N):
        obs, reward, done, info = e.last()
        assert info == {}
        rec["agent"].append(pid_of(agent))
        rec["reward"].append(float(reward))
        rec["done"].append(bool(done))
        rec["obs"].append(obs["observations"].copy())
        rec["mask"].append(obs["action_mask"].copy())
        if not done:
            mask = obs["action_mask"]
            pool = np.flatnonzero(mask == 0) if live == illegal_at else np.flatnonzero(mask)
            action = int(rng.choice(pool))
            live += 1
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                e.step(action)
        else:
            action = -1
            e.step(None)
        rec["action"].append(action)
        rec["next_agent"].append(pid_of(e.agent_selection) if e.agents else -1)
        rec["n_agents"].append(len(e.agents))
        cum = np.full(N, np.nan)
        for a, r in e._cumulative_rewards.items():
            cum[pid_of(a)] = r
        rec["cumulative"].append(cum)
    assert not e.agents
    return rec, live


def main():
    mg = _load_make_golden()
    from rlskyjo.environment import skyjo_env
    out = {}
    for ci, (name, cfg, kind, games) in enumerate(CONFIGS):
        rng = np.random.default_rng(4200 + ci)
        seed = 8800 + ci
        N = cfg["num_players"]
        per, decks, flipss = [], [], []
        for gi in range(games):
            deck = mg.make_deck(rng, kind)
            flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
            illegal_at = int(rng.integers(0, 60)) if gi % 2 else -1
            rec, live = play_env(mg, skyjo_env, cfg, deck, flips, rng, seed, gi, illegal_at)
            per.append(rec)
            decks.append(deck)
            flipss.append(flips)
            print(f"{name} game {gi}: {len(rec['agent'])} iterations, {live} live steps, illegal_at={illegal_at}, "
                  f"final rewards {[r for r, d in zip(rec['reward'], rec['done']) if d]}")
        pre = name + "/"
        for k, v in cfg.items():
            out[pre + k] = v
        out[pre + "seed"] = seed
        out[pre + "decks"] = np.stack(decks).astype(np.int8)
        out[pre + "flips"] = np.stack(flipss).astype(np.uint8)
        out[pre + "lengths"] = np.array([len(p["agent"]) for p in per], dtype=np.int32)
        cat = lambda k, dt: np.concatenate([np.asarray(p[k]) for p in per]).astype(dt)  # noqa: E731
        for k, dt in (("agent", np.int8), ("reward", np.float64), ("done", np.uint8), ("obs", np.int8), ("mask", np.int8),
                      ("action", np.int8), ("next_agent", np.int8), ("n_agents", np.int8), ("cumulative", np.float64)):
            out[pre + k] = cat(k, dt)
    out["names"] = np.array([c[0] for c in CONFIGS])
    np.savez_compressed(os.path.join(HERE, "env_trace.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
