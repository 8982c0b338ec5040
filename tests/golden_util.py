"""Helpers to iterate the reference-generated fixtures in tests/golden/."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    # n<players>_*.npz: per-config fixtures from make_golden.py (notebook_trace.npz has its own test)
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f[0] == "n" and f[1].isdigit())


class Golden:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.N = int(z["num_players"])
        self.indirect = bool(z["indirect"])
        self.score_penalty = float(z["score_penalty"])
        self.mean_reward = float(z["mean_reward"])
        self.reward_refunded = float(z["reward_refunded"])
        self.seed = int(z["seed"])
        self.z = {k: z[k] for k in z.files}
        self.lengths = self.z["lengths"]
        self.offsets = np.concatenate([[0], np.cumsum(self.lengths)])
        self.games = len(self.lengths)

    def game(self, gi):
        a, b = self.offsets[gi], self.offsets[gi + 1]
        per_step = ("agent", "phase", "action", "obs", "mask", "obs_other", "mask_other", "hand", "top")
        per_game = ("decks", "flips", "final_score", "reward", "num_refunded", "num_placed", "final_cards",
                    "final_masked", "final_obs", "final_mask", "final_agent", "n_reshuffles")
        out = {k: self.z[k][a:b] for k in per_step}
        out.update({k: self.z[k][gi] for k in per_game})
        return out


def fingerprint(N):
    """Per-game statistics of the unmodified reference under policy_ra (make_fingerprint.py)."""
    import json
    with open(os.path.join(GOLDEN_DIR, "fingerprint_policy_ra.json")) as f:
        return json.load(f)["configs"][str(N)]


def fingerprint_close(ref, key, value, n, sigmas=5.0):
    """|value - reference mean| within `sigmas` standard errors of the difference of two sample means
    (the reference's sd is used for both samples)."""
    r = ref[key]
    tol = sigmas * r["sd"] * (1.0 / ref["games"] + 1.0 / n) ** 0.5
    assert abs(value - r["mean"]) <= tol, f"{key}: {value:.4f} vs reference {r['mean']:.4f} +- {tol:.4f}"
