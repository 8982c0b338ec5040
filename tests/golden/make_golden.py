"""Generate tests/golden/*.npz by running the UNMODIFIED reference (rlskyjo, imported from
/root/reference, numba JIT) on injected decks and seeded action sequences.

Run here (the reference does not exist on the GPU box):   python tests/golden/make_golden.py

What is the reference's and what is ours:
  * every obs / mask / score / reward / metric in the fixtures is computed by the reference's
    own code: SkyjoGame.collect_observation / act (rlskyjo/game/skyjo.py:148-199, 308-335) and
    SimpleSkyjoEnv._calc_final_rewards (rlskyjo/environment/skyjo_env.py:293-312);
  * the deck order, the two flipped slots, the action sequence and the permutation used by
    in-game reshuffles are INPUTS, injected as SURVEY.md 9.8 describes
    (SkyjoGame._reshuffle_discard_pile is monkey-patched with tests/rng_twin.reshuffle so the
    GPU can reproduce np.random.shuffle's role deterministically).
gym / pettingzoo are not installed; the stand-ins of tests/shims are put on the path so that skyjo_env.py
imports -- none of their code is exercised by _calc_final_rewards.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")

import rng_twin  # noqa: E402


def _stub_modules():
    """gym / pettingzoo are not installed: put the stand-ins of tests/shims on the path (the same ones
    tests/golden/make_env_trace.py runs the whole env on), unless the real packages are importable."""
    import importlib.util
    if importlib.util.find_spec("gym") is None or importlib.util.find_spec("pettingzoo") is None:
        sys.path.insert(0, os.path.join(os.path.dirname(HERE), "shims"))


_stub_modules()
from rlskyjo.environment.skyjo_env import SimpleSkyjoEnv  # noqa: E402
from rlskyjo.game.skyjo import SkyjoGame  # noqa: E402

# one context shared by every copy of this module a test session loads (the hook below is installed on the class, so
# the copy that patched last must see the context whichever copy sets it)
_RESHUFFLE = getattr(SkyjoGame, "_b200_reshuffle_ctx", None) or {"ctx": None}
SkyjoGame._b200_reshuffle_ctx = _RESHUFFLE
_orig_reshuffle = SkyjoGame._reshuffle_discard_pile


def _patched_reshuffle(old_pile):
    ctx = _RESHUFFLE["ctx"]
    if ctx is None:  # during the constructor's own reset(): result is overwritten anyway
        lst = [int(x) for x in old_pile]
    else:
        lst = rng_twin.reshuffle(ctx["seed"], ctx["env"], ctx["episode"], ctx["q"], list(old_pile))
        ctx["q"] += 1
        ctx["n"] += 1
    top = lst.pop()
    return lst, [top]


SkyjoGame._reshuffle_discard_pile = staticmethod(_patched_reshuffle)


def make_deck(rng, kind):
    if kind == "standard":
        deck = np.repeat(np.arange(-2, 13, dtype=np.int8), 10)
    elif kind == "dense":  # 10 values x 15 copies: many column removals, bins stay <= 15
        vals = rng.choice(np.arange(-2, 13), size=10, replace=False)
        deck = np.repeat(vals.astype(np.int8), 15)
    else:
        raise ValueError(kind)
    rng.shuffle(deck)
    return deck


def play(N, indirect, penalty, mean_reward, reward_refunded, deck, flips, rng, seed, env, greedy_refund):
    _RESHUFFLE["ctx"] = None
    g = SkyjoGame(num_players=N, score_penalty=penalty, observe_other_player_indirect=indirect)
    g.players_cards = deck[: 12 * N].reshape(N, 12).astype(np.int8).copy()
    masks = np.full((N, 12), 2, dtype=np.int8)
    for p in range(N):
        masks[p, flips[p, 0]] = 1
        masks[p, flips[p, 1]] = 1
    g.players_masked = masks
    rest = [int(x) for x in deck[12 * N:]]
    g.discard_pile, g.drawpile = [rest[-1]], rest[:-1]
    g._reset_start_player()
    _RESHUFFLE["ctx"] = {"seed": seed, "env": env, "episode": 0, "q": 0, "n": 0}

    rec = {k: [] for k in ("agent", "phase", "action", "obs", "mask", "obs_other", "mask_other", "hand", "top")}
    while not g.is_terminated:
        pid, phase = g.expected_action
        obs, mask = g.collect_observation(pid)
        other = (pid + 1) % N
        obs_o, mask_o = g.collect_observation(other)
        legal = np.flatnonzero(mask)
        action = int(rng.choice(legal))
        if greedy_refund and phase == "place" and rng.random() < 0.7:
            # prefer a swap that completes a column of three equal open cards
            hand = g.hand_card
            for s in legal[legal < 12]:
                col = g.players_cards[pid][3 * (s // 3): 3 * (s // 3) + 3].copy()
                m = g.players_masked[pid][3 * (s // 3): 3 * (s // 3) + 3].copy()
                col[s % 3] = hand
                m[s % 3] = 1
                if col.min() == col.max() and np.all(m == 1):
                    action = int(s)
                    break
        rec["agent"].append(pid)
        rec["phase"].append(0 if phase == "draw" else 1)
        rec["action"].append(action)
        rec["obs"].append(obs)
        rec["mask"].append(mask)
        rec["obs_other"].append(obs_o)
        rec["mask_other"].append(mask_o)
        rec["hand"].append(g.hand_card)
        rec["top"].append(g.discard_pile[-1] if g.discard_pile else -3)
        g.act(pid, action)
    metrics = g.get_game_metrics()
    ns = types.SimpleNamespace(mean_reward=mean_reward, reward_refunded=reward_refunded)
    reward = SimpleSkyjoEnv._calc_final_rewards(ns, **metrics)
    pid, _ = g.expected_action
    final_obs, final_mask = g.collect_observation(pid)
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(
        final_score=np.array(metrics["final_score"], dtype=np.float64),
        reward=np.array(reward, dtype=np.float64),
        num_refunded=np.array(metrics["num_refunded"]),
        num_placed=np.array(metrics["num_placed"]),
        final_cards=g.players_cards.copy(),
        final_masked=g.players_masked.copy(),
        final_obs=final_obs, final_mask=final_mask, final_agent=pid,
        n_reshuffles=_RESHUFFLE["ctx"]["n"],
        discard_sorted_len=len(g.discard_pile), draw_len=len(g.drawpile),
    )
    return out


CONFIGS = [
    # name, N, indirect, penalty, mean_reward, reward_refunded, deck kind, games, greedy
    ("n1_direct", 1, False, 2.0, 1.0, 0.0, "standard", 6, False),
    ("n2_direct", 2, False, 2.0, 1.0, 0.0, "standard", 10, False),
    ("n2_indirect", 2, True, 2.0, 1.0, 0.001, "standard", 8, False),
    ("n3_indirect_default", 3, True, 2.0, 1.0, 0.001, "standard", 8, False),
    ("n3_direct_dense", 3, False, 2.0, 0.0, 0.01, "dense", 10, True),
    ("n4_direct", 4, False, 2.0, 1.0, 0.0, "standard", 12, False),
    ("n4_direct_dense", 4, False, 1.1, -1.0, 0.01, "dense", 12, True),
    ("n4_indirect_dense", 4, True, 3.5, 1.0, 0.0, "dense", 8, True),
    ("n5_direct", 5, False, 1.0, 1.0, 0.0, "standard", 6, False),
    ("n8_direct", 8, False, 2.0, 1.0, 0.0, "standard", 10, False),
    ("n8_direct_dense", 8, False, 0.7, 0.5, 0.01, "dense", 8, True),
    ("n8_indirect", 8, True, 2.0, 1.0, 0.001, "standard", 6, False),
    ("n12_direct", 12, False, 2.0, 1.0, 0.0, "standard", 6, False),
    ("n12_indirect_dense", 12, True, 1.3, 0.0, 0.01, "dense", 6, True),
    # player counts at and above the warp-assist threshold (skyjo_step.cuh SKYJO_ASSIST_MIN_N = 6)
    ("n6_direct", 6, False, 2.0, 1.0, 0.0, "standard", 8, False),
    ("n6_direct_dense", 6, False, 2.5, 1.0, 0.02, "dense", 8, True),
    ("n7_indirect_dense", 7, True, 1.5, 1.0, 0.01, "dense", 6, True),
    ("n10_direct", 10, False, 2.0, 0.0, 0.0, "standard", 6, False),
]


def main():
    total_steps = 0
    for ci, (name, N, ind, pen, mr, rr, kind, games, greedy) in enumerate(CONFIGS):
        rng = np.random.default_rng(1000 + ci)
        seed = 7000 + ci
        per = []
        decks, flipss = [], []
        for gi in range(games):
            deck = make_deck(rng, kind)
            flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
            per.append(play(N, ind, pen, mr, rr, deck, flips, rng, seed, gi, greedy))
            decks.append(deck)
            flipss.append(flips)
        lengths = np.array([len(p["action"]) for p in per], dtype=np.int32)
        total_steps += int(lengths.sum())
        cat = lambda k, dt: np.concatenate([p[k] for p in per]).astype(dt)  # noqa: E731
        stack = lambda k, dt: np.stack([p[k] for p in per]).astype(dt)  # noqa: E731
        np.savez_compressed(
            os.path.join(HERE, f"{name}.npz"),
            num_players=N, indirect=ind, score_penalty=pen, mean_reward=mr, reward_refunded=rr,
            seed=seed, decks=np.stack(decks).astype(np.int8), flips=np.stack(flipss).astype(np.uint8),
            lengths=lengths,
            agent=cat("agent", np.int8), phase=cat("phase", np.int8), action=cat("action", np.int8),
            obs=cat("obs", np.int8), mask=cat("mask", np.int8),
            obs_other=cat("obs_other", np.int8), mask_other=cat("mask_other", np.int8),
            hand=cat("hand", np.int8), top=cat("top", np.int8),
            final_score=stack("final_score", np.float64), reward=stack("reward", np.float64),
            num_refunded=stack("num_refunded", np.int32), num_placed=stack("num_placed", np.int32),
            final_cards=stack("final_cards", np.int8), final_masked=stack("final_masked", np.int8),
            final_obs=stack("final_obs", np.int8), final_mask=stack("final_mask", np.int8),
            final_agent=np.array([p["final_agent"] for p in per], dtype=np.int8),
            n_reshuffles=np.array([p["n_reshuffles"] for p in per], dtype=np.int32),
        )
        print(f"{name}: games={games} steps={int(lengths.sum())} refunds={int(sum(p['num_refunded'].sum() for p in per))} "
              f"reshuffles={int(sum(p['n_reshuffles'] for p in per))}")
    # the reference's own recorded known answer: notebooks/trainpettingzoo.ipynb:52745-52758
    print("total steps", total_steps)


if __name__ == "__main__":
    main()
