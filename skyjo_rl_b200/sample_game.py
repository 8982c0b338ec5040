"""Batched counterpart of the reference's driver loop `sample_run`
(rlskyjo/game/sample_game.py:5-28): play at least `games` full games with the uniform random
admissible policy (`policy_ra`, random_admissible_policy.py:6-28).  All envs of the batch play in
lockstep; the policy is drawn in-kernel, finished envs restart at once.

    python -m skyjo_rl_b200.sample_game [--games 5000] [--players 2] [--envs 4096] [--verbose]
"""
import argparse

from .env import BatchedSkyjoEnv, render_action_explainer


def sample_run(games=5000, verbose=0, config={"num_players": 2}, num_envs=4096, device="cuda:0", seed=0):  # noqa: B006
    """Returns the episode statistics (dict) after >= `games` finished games."""
    env = BatchedSkyjoEnv(num_envs=min(num_envs, max(games, 1)), device=device, seed=seed, **config)
    env.reset()
    chunk = 8
    while True:
        if verbose:  # follow env 0 the way the reference prints its single game
            for _ in range(chunk):
                print(env.game_view(0).render_table())
                env.step_random(1)
        else:
            env.step_random(chunk)
        st = env.stats()
        if st["episodes"] >= games:
            env.check()
            env.close()
            return st


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=5000)
    ap.add_argument("--players", type=int, default=2)
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    out = sample_run(a.games, int(a.verbose), {"num_players": a.players}, a.envs)
    print({k: v for k, v in out.items() if not k.startswith("wins_seat") or v})
    print("legend:", render_action_explainer(24), "/", render_action_explainer(0))
