"""Generate tests/golden/scoring_kat.npz: known answers of the reference's end-of-game arithmetic on boards built to
hit its corner cases: ties for the best score (penalty forgiven), a finisher with a NEGATIVE score
that is not the best (multiplied all the same), columns of three equal cards (worth nothing, removed or not),
fractional and < 1 penalties.

Run here (the reference does not exist on the GPU box):   python tests/golden/make_scoring_kat.py

The boards are low cards (-2 .. 2, at most ten of each as in a real deck) so that sums collide and go negative.  Each
case is a whole game: every player draws from the pile and flips his first hidden slot, so the boards stay as dealt
(equal columns are removed on the way) and the starter finishes first.  The game is played by the C oracle, which
supplies the final boards, the finisher and the refund counts; the ANSWERS stored here -- final scores and rewards --
are computed from those by the unmodified reference: `SkyjoGame._evaluate_game` (rlskyjo/game/skyjo.py:477-498) and
`SimpleSkyjoEnv._calc_final_rewards` (rlskyjo/environment/skyjo_env.py:293-312).
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))


def flip_action(mask):
    """draw phase: the draw pile; place phase: discard the drawn card and reveal the first hidden slot"""
    if mask[24]:
        return 24
    hidden = np.flatnonzero(mask[12:24])
    return 12 + int(hidden[0]) if len(hidden) else int(np.flatnonzero(mask)[0])


def make_case(rng, N):
    low = np.repeat(np.arange(-2, 3, dtype=np.int8), 10)            # 50 low cards
    rng.shuffle(low)
    boards = low[: 12 * N].copy()
    if rng.random() < 0.5:                                          # plant an equal column somewhere
        p, c, v = int(rng.integers(N)), int(rng.integers(4)), boards[0]
        idx = np.flatnonzero(boards == v)
        if len(idx) >= 3:
            tgt = [12 * p + 3 * c + j for j in range(3)]
            for t, i in zip(tgt, idx[:3]):
                boards[t], boards[i] = boards[i], boards[t]
    rest = np.concatenate([low[12 * N:], np.repeat(np.arange(3, 13, dtype=np.int8), 10)])
    rng.shuffle(rest)
    deck = np.concatenate([boards, rest]).astype(np.int8)
    flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
    return deck, flips


def play_flip_game(game, deck, flips):
    game.reset_injected(deck, flips)
    actions = []
    while True:
        pid = game.expected_action[0]
        _, mask = game.collect_observation(pid)
        a = flip_action(mask)
        actions.append(a)
        if game.act(pid, a):
            return pid, actions


def main():
    spec = importlib.util.spec_from_file_location("make_golden_kat", os.path.join(HERE, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)          # imports the reference over tests/shims
    from oracle import oracle as O
    rng = np.random.default_rng(777)
    cases = []
    tally = {"tie_forgiven": 0, "negative_penalised": 0, "equal_column": 0, "penalised": 0}
    for ci in range(160):
        N = int(rng.choice([2, 3, 4]))
        penalty = float(rng.choice([0.5, 1.5, 2.0, 3.0]))
        mr, rr = float(rng.choice([0.0, 1.0])), float(rng.choice([0.0, 0.01]))
        deck, flips = make_case(rng, N)
        g = O.OracleGame(N, penalty, False)
        finisher, actions = play_flip_game(g, deck, flips)
        cards = g.players_cards.copy()
        refunded = np.array(g.game_metrics["num_refunded"], dtype=np.int64)
        score = mg.SkyjoGame._evaluate_game(cards, finisher, penalty)                      # the reference
        ns = types.SimpleNamespace(mean_reward=mr, reward_refunded=rr)
        reward = mg.SimpleSkyjoEnv._calc_final_rewards(ns, final_score=score, num_refunded=list(refunded))
        raw = mg.SkyjoGame._evaluate_game(cards, finisher, 1.0)
        if raw[finisher] != min(raw):
            tally["penalised"] += 1
            tally["negative_penalised"] += raw[finisher] < 0
        elif sum(1 for x in raw if x == raw[finisher]) > 1:
            tally["tie_forgiven"] += 1
        tally["equal_column"] += int(any(c[3 * k] == c[3 * k + 1] == c[3 * k + 2] for c in cards for k in range(4)))
        cases.append(dict(N=N, penalty=penalty, mr=mr, rr=rr, deck=deck, flips=flips, finisher=finisher,
                          score=np.array(score, np.float64), reward=np.array(reward, np.float64),
                          cards=cards, steps=len(actions)))
    print(tally)
    assert all(v >= 3 for v in tally.values()), tally
    C = len(cases)

    def pad(key, shape, dtype, fill):
        a = np.full((C,) + shape, fill, dtype=dtype)
        for i, c in enumerate(cases):
            v = np.asarray(c[key])
            a[i][tuple(slice(0, n) for n in v.shape)] = v
        return a

    np.savez_compressed(
        os.path.join(HERE, "scoring_kat.npz"),
        N=np.array([c["N"] for c in cases], np.int8), penalty=np.array([c["penalty"] for c in cases]),
        mr=np.array([c["mr"] for c in cases]), rr=np.array([c["rr"] for c in cases]),
        finisher=np.array([c["finisher"] for c in cases], np.int8), steps=np.array([c["steps"] for c in cases], np.int32),
        deck=pad("deck", (150,), np.int8, 0), flips=pad("flips", (4, 2), np.uint8, 0),
        cards=pad("cards", (4, 12), np.int8, 0), score=pad("score", (4,), np.float64, np.nan),
        reward=pad("reward", (4,), np.float64, np.nan))

if __name__ == "__main__":
    main()
