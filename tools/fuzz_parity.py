#!/usr/bin/env python
"""Randomised differential runs beyond the committed tests (development aid; CPU only).

    python tools/fuzz_parity.py hostsim   SECONDS    # host-compiled kernels vs the C oracle: random player counts,
                                                     # observation / reset modes, penalties, batch sizes, lockstep
                                                     # rollouts with the in-kernel policy and external action streams
                                                     # with illegal actions and truncation (tests/parity_util.py)
    python tools/fuzz_parity.py reference SECONDS    # the LIVE reference (/root/reference, build container only) vs
                                                     # the C oracle: random injected decks (standard and dense), N = 1..12

    python tools/fuzz_parity.py gpu       SECONDS    # the same as hostsim, on the CUDA build (under gpurun)
    python tools/fuzz_parity.py gpu-chunked SECONDS  # skyjo_step_random(n >= 2) on the CUDA build: env-range streams and
                                                     # refill windows, every env against the oracle at chunk boundaries
    python tools/fuzz_parity.py reference-strategy SECONDS   # the live reference vs the oracle on strategy-played games
    python tools/fuzz_parity.py strategy  SECONDS    # host-compiled kernels vs the C oracle on games played by the
                                                     # hoarder / hunter / closer strategies (tests/test_strategy_games.py)

Prints every mismatch with its parameters and a final count.  Round-1 totals (no mismatch): reference 36 178 games
(7.1 M steps, 65 872 reshuffles), reference-strategy 3 076 games (3.3 M steps, 13 402 reshuffles, 5 067 removals),
hostsim 3 676 runs, strategy 3 836 games (N = 1 .. 12; 7.2 M steps, 94 966 reshuffles, 8 638 removals), gpu 45 runs.
Round 2 (no mismatch): gpu-chunked 501 runs (83.3 M env-steps; the last 203 with the graph-replayed calls), gpu 418 runs.
Round 2, CPU modes after the last refactor of skyjo_core.cuh (no mismatch): hostsim 1 054 runs, reference 16 605 games
(3.3 M steps, 30 378 reshuffles), strategy 839 games (1.8 M steps, 28 674 reshuffles, 2 166 removals),
reference-strategy 1 233 games (1.3 M steps, 5 343 reshuffles, 2 110 removals)."""
import importlib.util
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402


def fuzz_hostsim(seconds, gpu=False):
    import parity_util as P
    if gpu:                                   # the CUDA build itself (needs a B200): warp-level staging, warp assist
        from skyjo_rl_b200 import BatchedSkyjoEnv as HostSimEnv
    else:
        from hostsim.sim import HostSimEnv
    rng = np.random.default_rng(int(time.time()))
    t_end, runs, bad = time.time() + seconds, 0, 0
    while time.time() < t_end:
        N, ind, mode = int(rng.integers(1, 13)), bool(rng.integers(2)), int(rng.choice([1, 2]))
        pen, mr = float(rng.choice([0.5, 1.0, 2.0, 3.3])), float(rng.choice([-1.0, 0.0, 1.0]))
        rr, B = float(rng.choice([0.0, 0.01, 0.25])), int(rng.integers(1, 200 if gpu else 40))
        T = int(rng.integers(60 * N + 100, 60 * N + 600))          # long enough for games to end
        params = dict(N=N, indirect=ind, mode=mode, penalty=pen, mr=mr, rr=rr, B=B, T=T)
        try:
            if rng.integers(3) < 2:
                P.rng_rollout(HostSimEnv, N, ind, pen, mr, rr, B, T, mode)
            else:
                params.update(mode=int(rng.choice([0, 1, 2])), max_steps=int(rng.choice([0, 0, 40, 200])),
                              p_illegal=float(rng.choice([0.0, 0.02, 0.1])), seed=int(rng.integers(1, 10 ** 6)))
                P.external_actions_rollout(HostSimEnv, N, ind, params["mode"], params["max_steps"], B, min(T, 300),
                                           p_illegal=params["p_illegal"], seed=params["seed"])
            runs += 1
        except AssertionError:
            bad += 1
            print("MISMATCH", params, flush=True)
            traceback.print_exc()
    print(f"{'GPU' if gpu else 'hostsim'} vs oracle: {runs} runs ok, {bad} mismatches")
    return bad


def fuzz_chunked(seconds):
    """The multi-stream / refill-window path of skyjo_step_random on the CUDA build (under gpurun): random player
    counts, env-range counts, chunk lengths (the window length depends on N), reset modes -- every env against the
    oracle at every chunk boundary, plus the exported state (tests/parity_util.chunked_rollout)."""
    import parity_util as P
    from skyjo_rl_b200 import BatchedSkyjoEnv
    rng = np.random.default_rng(int(time.time()))
    t_end, runs, bad, steps = time.time() + seconds, 0, 0, 0
    while time.time() < t_end:
        N, ind, mode = int(rng.integers(1, 13)), bool(rng.integers(2)), int(rng.choice([1, 2]))
        B, ranges = int(rng.integers(1, 1500)), int(rng.integers(1, 9))
        chunks = [int(rng.choice([2, 3, 7, 8, 9, 16, 31, 32, 33, 40, 64, 70])) for _ in range(int(rng.integers(4, 14)))]
        params = dict(N=N, indirect=ind, mode=mode, B=B, ranges=ranges, chunks=chunks)
        try:
            s, _, _ = P.chunked_rollout(lambda **kw: BatchedSkyjoEnv(**kw), N, ind, float(rng.choice([1.0, 2.0, 3.3])), 1.0,
                                        float(rng.choice([0.0, 0.01])), B, chunks, reset_mode=mode, ranges=ranges,
                                        seed=int(rng.integers(1, 10 ** 6)), first_env=int(rng.integers(0, 10 ** 9)))
            runs += 1
            steps += s
        except AssertionError:
            bad += 1
            print("MISMATCH", params, flush=True)
            traceback.print_exc()
    print(f"GPU chunked step_random vs oracle: {runs} runs ok ({steps} env-steps), {bad} mismatches")
    return bad


def fuzz_reference(seconds):
    from oracle import oracle as O
    spec = importlib.util.spec_from_file_location("make_golden_fuzz", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rng = np.random.default_rng(int(time.time()) ^ 0x5A5A)
    t_end, games, steps, resh, bad = time.time() + seconds, 0, 0, 0, 0
    while time.time() < t_end:
        N, indirect = int(rng.integers(1, 13)), bool(rng.integers(2))
        kind = str(rng.choice(["standard", "dense"]))
        penalty, mr = float(rng.choice([0.5, 1.0, 1.5, 2.0, 3.0])), float(rng.choice([-1.0, 0.0, 1.0]))
        rr, seed, gi = float(rng.choice([0.0, 0.01])), int(rng.integers(1, 2 ** 40)), int(rng.integers(0, 1000))
        deck = mg.make_deck(rng, kind)
        flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
        try:
            g = mg.play(N, indirect, penalty, mr, rr, deck, flips, rng, seed, gi, kind == "dense")
            og = O.OracleGame(N, penalty, indirect)
            og.reset_injected(deck, flips)
            og.set_rng_reshuffle(seed, gi, 0)
            for t in range(len(g["action"])):
                pid = og.expected_action[0]
                assert pid == g["agent"][t]
                obs, mask = og.collect_observation(pid)
                assert np.array_equal(obs, g["obs"][t]) and np.array_equal(mask, g["mask"][t]), t
                oo, mo = og.collect_observation((pid + 1) % N)
                assert np.array_equal(oo, g["obs_other"][t]) and np.array_equal(mo, g["mask_other"][t]), t
                assert og.act(pid, int(g["action"][t])) == (t == len(g["action"]) - 1)
            assert np.array(og.game_metrics["final_score"]).tobytes() == g["final_score"].tobytes()
            assert og.final_rewards(mr, rr).tobytes() == g["reward"].tobytes()
            assert og.n_reshuffles == g["n_reshuffles"]
            games, steps, resh = games + 1, steps + len(g["action"]), resh + int(g["n_reshuffles"])
        except AssertionError:
            bad += 1
            print("MISMATCH", dict(N=N, indirect=indirect, kind=kind, penalty=penalty, seed=seed, game=gi), flush=True)
            traceback.print_exc()
    print(f"live reference vs oracle: {games} games ok ({steps} steps, {resh} reshuffles), {bad} mismatches")
    return bad


def fuzz_strategy(seconds):
    from hostsim.sim import HostSimEnv
    import test_strategy_games as S
    rng = np.random.default_rng(int(time.time()) ^ 0xC3C3)
    t_end, games, steps, resh, removed, bad = time.time() + seconds, 0, 0, 0, 0, 0
    while time.time() < t_end:
        N, ind = int(rng.integers(1, 13)), bool(rng.integers(2))
        pen, mr, rr = float(rng.choice([0.5, 2.0, 3.0])), float(rng.choice([0.0, 1.0])), float(rng.choice([0.0, 0.01]))
        seed, env_id = int(rng.integers(1, 10 ** 9)), int(rng.integers(0, 10 ** 6))
        try:
            t, r, f = S.play_against_oracle(HostSimEnv, N, ind, pen, mr, rr, seed, env_id, np.random.default_rng(seed))
            games, steps, resh, removed = games + 1, steps + t, resh + r, removed + f
        except AssertionError:
            bad += 1
            print("MISMATCH", dict(N=N, indirect=ind, penalty=pen, seed=seed, env=env_id), flush=True)
            traceback.print_exc()
    print(f"strategy games, hostsim vs oracle: {games} games ok ({steps} steps, {resh} reshuffles, {removed} removals), "
          f"{bad} mismatches")
    return bad


def fuzz_reference_strategy(seconds):
    import test_strategy_games as S
    spec = importlib.util.spec_from_file_location("make_golden_fuzz2", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    rng = np.random.default_rng(int(time.time()) ^ 0x1234)
    t_end, games, steps, resh, removed, bad = time.time() + seconds, 0, 0, 0, 0, 0
    while time.time() < t_end:
        N, ind = int(rng.integers(1, 7)), bool(rng.integers(2))
        pen, mr, rr = float(rng.choice([0.5, 2.0, 3.0])), float(rng.choice([0.0, 1.0])), float(rng.choice([0.0, 0.01]))
        seed, env_id = int(rng.integers(1, 2 ** 40)), int(rng.integers(0, 1000))
        kind = str(rng.choice(["standard", "dense"]))
        deck = mg.make_deck(rng, kind)
        flips = np.stack([rng.choice(12, 2, replace=False) for _ in range(N)]).astype(np.uint8)
        try:
            t, r, f = S.play_reference_vs_oracle(mg, N, ind, pen, mr, rr, deck, flips, seed, env_id, rng)
            games, steps, resh, removed = games + 1, steps + t, resh + r, removed + f
        except AssertionError:
            bad += 1
            print("MISMATCH", dict(N=N, indirect=ind, kind=kind, penalty=pen, seed=seed, env=env_id), flush=True)
            traceback.print_exc()
    print(f"strategy games, live reference vs oracle: {games} games ok ({steps} steps, {resh} reshuffles, "
          f"{removed} removals), {bad} mismatches")
    return bad


if __name__ == "__main__":
    which, seconds = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 60.0
    fn = {"hostsim": fuzz_hostsim, "reference": fuzz_reference, "strategy": fuzz_strategy,
          "gpu": lambda sec: fuzz_hostsim(sec, gpu=True), "gpu-chunked": fuzz_chunked,
          "reference-strategy": fuzz_reference_strategy}[which]
    sys.exit(1 if fn(seconds) else 0)
