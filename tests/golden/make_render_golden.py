"""Generate tests/golden/render_n3.json with the UNMODIFIED reference (run here; the reference does
not exist on the GPU box):   python tests/golden/make_render_golden.py

One injected 3-player game (deck order, open slots and actions are inputs); the recorded outputs are
the reference's own SkyjoGame.render_table() strings (rlskyjo/game/skyjo.py:507-564) at a few
steps and after game over, render_action_explainer(a) for every action (:566-590) and
render_actions() (:592-602)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from rlskyjo.game.skyjo import SkyjoGame  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    N = 3
    deck = np.repeat(np.arange(-2, 13, dtype=np.int8), 10)
    rng.shuffle(deck)
    flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
    g = SkyjoGame(num_players=N, score_penalty=2.0, observe_other_player_indirect=False)
    g.players_cards = deck[: 12 * N].reshape(N, 12).astype(np.int8).copy()
    masks = np.full((N, 12), 2, dtype=np.int8)
    for p in range(N):
        masks[p, flips[p, 0]] = 1
        masks[p, flips[p, 1]] = 1
    g.players_masked = masks
    rest = [int(x) for x in deck[12 * N:]]
    g.discard_pile, g.drawpile = [rest[-1]], rest[:-1]
    g._reset_start_player()
    actions, renders = [], {}
    t = 0
    while not g.is_terminated:
        if t in (0, 1, 7, 30):
            renders[str(t)] = g.render_table()
        pid, _ = g.expected_action
        _, mask = g.collect_observation(pid)
        a = int(rng.choice(np.flatnonzero(mask)))
        actions.append(a)
        g.act(pid, a)
        t += 1
    renders["final"] = g.render_table()
    out = {"num_players": N, "deck": deck.tolist(), "flips": flips.tolist(), "actions": actions, "renders": renders,
           "explainer": [SkyjoGame.render_action_explainer(a) for a in range(26)],
           "render_actions": SkyjoGame.render_actions()}
    with open(os.path.join(HERE, "render_n3.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("steps", len(actions), "renders", list(renders))


if __name__ == "__main__":
    main()
