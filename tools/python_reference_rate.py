"""BASELINE config 1 on the UNMODIFIED Python reference: N = 2, policy_ra, 10 000 games over
multiprocessing.Pool(os.cpu_count()), each worker running the loop of rlskyjo/game/sample_game.py:10-21 with a step
counter (seeded with set_seed(100 + rank)); plus single-process figures for N = 2 / 4 / 8.

The reference is imported from --ref: baseline/_ref (the offline `pip install --target baseline/_ref` of the
reference that __graft_entry__.build() performs in the build container; git-ignored, it travels to the GPU box with
the tree) or /root/reference where that is mounted.  numba JIT when numba imports, otherwise a no-op `njit` stand-in
(the reference's own fallback, skyjo.py:13-16, cannot decorate `@njit()`), and the output says which.

    python tools/python_reference_rate.py [--ref DIR] [--games 10000] [--json]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_ref(explicit=None):
    for d in ([explicit] if explicit else []) + [os.path.join(ROOT, "baseline", "_ref"), "/root/reference"]:
        if d and os.path.isdir(os.path.join(d, "rlskyjo", "game")):
            return d
    return None


def ensure_numba():
    """True if numba's JIT is used; otherwise installs a no-op `numba.njit` so that skyjo.py imports."""
    try:
        import numba  # noqa: F401
        return True
    except Exception:  # noqa: BLE001
        m = types.ModuleType("numba")

        def njit(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f
        m.njit = njit
        sys.modules["numba"] = m
        return False


def worker(args):
    rank, games, N, ref = args
    if ref not in sys.path:
        sys.path.insert(0, ref)
    ensure_numba()
    import numpy as np
    from rlskyjo.game.skyjo import SkyjoGame
    from rlskyjo.models.random_admissible_policy import policy_ra
    g = SkyjoGame(num_players=N)
    g.set_seed(100 + rank)
    rng = np.random.default_rng(rank)

    def play(n):
        steps = 0
        for _ in range(n):
            g.reset()
            while not g.is_terminated:
                pid, _ = g.expected_action
                obs, mask = g.collect_observation(pid)
                g.act(pid, policy_ra(obs, mask, rng))
                steps += 1
        return steps
    play(3)                                   # numba JIT warm-up
    t0 = time.perf_counter()
    steps = play(games)
    return steps, time.perf_counter() - t0


def measure(ref, games=10000, players=2, procs=None, singles=((2, 300), (4, 200), (8, 100))):
    P = procs or os.cpu_count() or 1
    jit = ensure_numba()
    per = max(1, games // P)
    with mp.get_context("spawn").Pool(P) as pool:   # spawn: the caller may hold a CUDA context
        t0 = time.perf_counter()
        res = pool.map(worker, [(r, per, players, ref) for r in range(P)])
        wall = time.perf_counter() - t0
    steps = sum(s for s, _ in res)
    slowest = max(t for _, t in res)
    out = {"value": steps / slowest, "unit": "env-steps/s", "cores": P, "jit": "numba" if jit else "no-op njit stand-in",
           "games": per * P, "num_players": players, "env_steps": steps, "seconds_slowest_worker": slowest,
           "wall_seconds_with_jit_warmup": wall, "source": os.path.relpath(ref, ROOT) if ref.startswith(ROOT) else ref,
           "loop": "rlskyjo/game/sample_game.py:10-21 (SkyjoGame + policy_ra), multiprocessing.Pool, set_seed(100 + rank)"}
    single = {}
    for N, g in singles:
        s, t = worker((0, g, N, ref))
        single[f"N{N}"] = s / t
    out["single_process"] = single
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=None)
    ap.add_argument("--games", type=int, default=10000)
    ap.add_argument("--players", type=int, default=2)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    ref = find_ref(a.ref)
    if ref is None:
        print(json.dumps({"unavailable": "no rlskyjo under baseline/_ref or /root/reference"}) if a.json
              else "reference not found")
        return
    out = measure(ref, a.games, a.players, a.procs or None)
    if a.json:
        print(json.dumps(out))
        return
    cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    print(f"# host: {out['cores']} cores, {cpu[0] if cpu else '?'}; reference from {out['source']}, {out['jit']}")
    print(f"config 1: N={out['num_players']}, {out['games']} games, Pool({out['cores']}): {out['env_steps']} env-steps, "
          f"slowest worker {out['seconds_slowest_worker']:.2f} s (wall incl. JIT warm-up "
          f"{out['wall_seconds_with_jit_warmup']:.1f} s) -> {out['value']:.0f} env-steps/s on {out['cores']} cores")
    for k, v in out["single_process"].items():
        print(f"single process, {k}: {v:.0f} env-steps/s")


if __name__ == "__main__":
    main()
