"""Turn the one recorded known-answer episode of the reference into a replayable fixture.

    python tests/golden/make_notebook_trace.py      (needs /root/reference; run in the build container)

/root/reference/notebooks/trainpettingzoo.ipynb (cell 5, output lines 105-2747 of the file) prints,
for one complete 3-player direct-observation game played through `skyjo_env.env()` -- i.e. through
PettingZoo 1.14's AEC machinery and wrapper stack on the author's machine -- every
`last()` tuple (observation dict, cumulative reward, done), the rendered board and the sampled
action, ending with `Results: {0: 46.0, 1: 118.0, 2: 87.0}`, the first done agent's reward and
the remaining `_cumulative_rewards`.  Nothing here is computed by our code except the
reconstruction of the INPUTS (deck order, open slots) from what the trace reveals:
  * a slot's initial card is what the first observation shows (open slots), what a flip reveals,
    what lands on the discard pile when a still-hidden slot is swapped, or the `u<value>` the
    final board prints for slots that stayed hidden;
  * the draw pile's order is the sequence of hand cards after every action 24; cards never drawn
    are filled from the remaining multiset (ten each of -2..12, rlskyjo/game/skyjo.py:80).
The fixture (tests/golden/notebook_trace.npz) holds those inputs, the actions and every
recorded output, for replay through the oracle, the host-compiled kernels and the GPU env.
"""
import json
import os
import re

import numpy as np

NB = "/root/reference/notebooks/trainpettingzoo.ipynb"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "notebook_trace.npz")
N = 3


def ints(s):
    return [int(x) for x in re.findall(r"-?\d+", s)]


def main():
    nb = json.load(open(NB))
    cell = nb["cells"][5]
    txt = "".join("".join(o.get("text", [])) for o in cell["outputs"] if o.get("output_type") == "stream")
    blocks = txt.split("training fct: ")[1:]
    obs, mask, reward, done, agent, action, top, hand, boards = [], [], [], [], [], [], [], [], []
    for b in blocks:
        m = re.match(r"\{'observations': array\((\[.*?\]), dtype=int8\), 'action_mask': array\((\[.*?\]), dtype=int8\)\} "
                     r"(\S+) (True|False) \{\}", b, re.S)
        obs.append(ints(m.group(1)))
        mask.append(ints(m.group(2)))
        reward.append(float(m.group(3)))
        done.append(m.group(4) == "True")
        t = re.search(r"discard pile top: (\S+)", b).group(1)
        h = re.search(r"holding card player \d+: (\S+)", b).group(1)
        top.append(-3 if t == "empty" else int(t))
        hand.append(15 if h == "empty" else int(h))
        grids = re.findall(r"======= Player (\d) =+ \n\[\[(.*?)\]\]", b, re.S)
        board = []
        for _, g in grids:
            rows = [r.replace("[", "").replace("]", "").split() for r in g.split("\n")]
            arr = np.array(rows, dtype=object)          # rendered as reshape(4,-1).T (skyjo.py:554): 3 x 4
            board.append(arr.T.reshape(-1).tolist())    # back to slot order 0..11
        boards.append(board)
        a = re.search(r"sampled action (player_\d): (\d+)", b)
        agent.append(int(re.search(r"next turn: \w+ by Player (\d)", b).group(1)))
        action.append(int(a.group(2)) if a else -1)
    T = len(blocks) - 1                                  # the last block is the post-game last()
    assert all(a >= 0 for a in action[:T]) and action[T] == -1 and done[T] and not any(done[:T])
    results = re.search(r"Results: \{0: ([\d.]+), 1: ([\d.]+), 2: ([\d.]+)\}", txt)
    final_score = [float(results.group(i)) for i in (1, 2, 3)]
    first_done_reward = float(re.search(r"\ndone (\S+)\n", txt).group(1))
    rest = re.search(r"\{'player_1': (\S+), 'player_2': (\S+)\}", txt)
    rewards = [first_done_reward, float(rest.group(1)), float(rest.group(2))]

    # ---- reconstruct the dealt table ---------------------------------------------------------
    o0 = np.array(obs[0][19:]).reshape(N, 12)
    init = np.full((N, 12), 99)
    init[o0 != 15] = o0[o0 != 15]
    flips = np.array([np.flatnonzero(o0[p] != 15) for p in range(N)], dtype=np.uint8)
    assert flips.shape == (N, 2)
    touched = o0 != 15                                   # slots whose initial card is already known
    draws = []
    for t in range(T):
        p, a = agent[t], action[t]
        nxt_cards = np.array(obs[t + 1][19:]).reshape(N, 12)
        if a == 24:
            draws.append(obs[t + 1][18])                 # the hand card after drawing from the pile
        elif a < 12 and not touched[p, a]:
            init[p, a] = obs[t + 1][17]                  # the swapped-out hidden card is the new discard top
            touched[p, a] = True
        elif 12 <= a < 24:
            assert not touched[p, a - 12]
            init[p, a - 12] = nxt_cards[p, a - 12]       # revealed by the flip
            touched[p, a - 12] = True
    for p in range(N):                                   # still hidden at the end: printed as u<value>
        for s, cell_txt in enumerate(boards[T][p]):
            if str(cell_txt).startswith("u"):
                assert not touched[p, s]
                init[p, s] = int(str(cell_txt)[1:])
                touched[p, s] = True
    assert touched.all() and (init != 99).all()
    first_discard = top[0]
    counts = {v: 10 for v in range(-2, 13)}
    for v in list(init.reshape(-1)) + draws + [first_discard]:
        counts[int(v)] -= 1
    assert min(counts.values()) >= 0, counts
    filler = [v for v in range(-2, 13) for _ in range(counts[v])]
    drawpile = filler + draws[::-1]                      # python list, top = last (skyjo.py:366)
    deck = list(init.reshape(-1)) + drawpile + [first_discard]
    assert len(deck) == 150
    np.savez_compressed(
        OUT, num_players=N, indirect=False, score_penalty=2.0, mean_reward=1.0, reward_refunded=0.0,
        deck=np.array(deck, np.int8), flips=flips, agent=np.array(agent, np.int8), action=np.array(action, np.int8),
        obs=np.array(obs, np.int8), mask=np.array(mask, np.int8), reward=np.array(reward), done=np.array(done),
        top=np.array(top, np.int8), hand=np.array(hand, np.int8), final_score=np.array(final_score),
        rewards=np.array(rewards), steps=T)
    print(f"wrote {OUT}: {T} actions ({action[:T].count(24)} draws, {sum(a == 25 for a in action[:T])} takes, "
          f"{sum(a < 12 for a in action[:T])} swaps, {sum(12 <= a < 24 for a in action[:T])} flips), "
          f"results {final_score}, rewards {rewards}")


if __name__ == "__main__":
    main()
