"""BASELINE config 1 on the UNMODIFIED Python reference (rlskyjo imported from /root/reference, numba JIT): N = 2,
policy_ra, 10 000 games over multiprocessing.Pool(os.cpu_count()), each worker running the loop of
rlskyjo/game/sample_game.py:10-21 with a step counter; plus single-process figures for N = 2 / 4 / 8.
Runs only where the reference is mounted (the build container -- it cannot travel to the GPU box); the
figures it prints are recorded in profiles/ next to the C-port baseline bench.py times on the GPU box's host.

    python tools/python_reference_rate.py > profiles/r1_v7_python_reference_container.txt
"""
import multiprocessing as mp
import os
import sys
import time

REF = "/root/reference"


def worker(args):
    rank, games, N = args
    sys.path.insert(0, REF)
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


def main():
    if not os.path.isdir(os.path.join(REF, "rlskyjo")):
        print("reference not mounted")
        return
    P = os.cpu_count()
    cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][:1]
    print(f"# host: {P} cores, {cpu[0] if cpu else '?'} (build container, not the GPU box)")
    with mp.Pool(P) as pool:
        t0 = time.perf_counter()
        res = pool.map(worker, [(r, 10000 // P, 2) for r in range(P)])
        wall = time.perf_counter() - t0
    steps = sum(s for s, _ in res)
    slowest = max(t for _, t in res)
    print(f"config 1: N=2, {10000 // P * P} games, Pool({P}): {steps} env-steps, slowest worker {slowest:.2f} s "
          f"(wall incl. JIT warm-up {wall:.1f} s) -> {steps / slowest:.0f} env-steps/s on {P} cores")
    for N, games in ((2, 400), (4, 250), (8, 120)):
        s, t = worker((0, games, N))
        print(f"single process, N={N}: {s} env-steps in {t:.2f} s -> {s / t:.0f} env-steps/s")


if __name__ == "__main__":
    main()
