"""Bit parity of the BENCHMARKED code path with the oracle, at bench scale.

bench.py drives `skyjo_step_random(h, 64)` on 2^20 (config 2) and 2^24 (config 3) envs.  From 2^18 envs that entry
is `step_random_ranges` (csrc/skyjo_capi.cu): the batch stepped as 4 independent env ranges on 4 CUDA streams, each
with its own flagged refill deals, in the phase-locked next-step reset mode.  The per-step oracle comparisons of
tests/test_gpu_parity.py use small batches and step_random(1), i.e. the single-stream path; here the multi-stream
path itself is compared with the oracle:

  * at small batches with the range count forced (skyjo_set_env_ranges), every env, chunks of odd and even sizes;
  * at BASELINE config 2 (N=4, 2^20 envs) and config 3 (N=8, 2^24 envs) at FULL size: >= 512 sampled envs -- runs of
    32 consecutive envs spread over the batch, the envs either side of every range boundary, the first and the last
    32 -- replayed on the oracle by global env id (an env's games depend only on (seed, global id)); at every chunk
    boundary obs / mask / agent / done / float64 rewards / final scores, at the end the complete exported state
    (hidden cards, both piles, episode index, metrics).

Reference loop being reproduced: /root/reference/rlskyjo/game/sample_game.py:10-21."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from parity_util import chunked_rollout, sample_blocks  # noqa: E402

pytestmark = pytest.mark.gpu


def _env(**kw):
    from skyjo_rl_b200 import BatchedSkyjoEnv
    return BatchedSkyjoEnv(**kw)


@pytest.mark.parametrize("N,indirect,B,ranges,chunks,mode", [
    (4, False, 700, 4, [64, 7, 13, 64, 2, 64, 64, 33, 64], 2),
    (4, False, 700, 4, [64, 64, 5, 64, 64, 64], 1),
    (2, True, 1000, 3, [9, 64, 64, 64, 2, 31], 2),
    (8, False, 520, 4, [64] * 9 + [3, 64], 2),
    (8, False, 1030, 8, [16, 64, 64, 64, 64, 64, 64, 64, 64, 64], 1),
    (12, False, 300, 2, [64] * 12, 2),
    (3, False, 130, 1, [2, 64, 64, 64, 64], 2),       # one range, n >= 2: the single-stream window path
])
def test_env_range_streams_match_oracle_on_every_env(N, indirect, B, ranges, chunks, mode):
    steps, ended, st = chunked_rollout(_env, N, indirect, 2.0, 1.0, 0.01, B, chunks, reset_mode=mode, ranges=ranges,
                                       first_env=12345)
    assert ended > B // 2 and st["steps"] == steps
    if N >= 8:
        assert st["reshuffles"] > 0


@pytest.mark.parametrize("N,B,chunks,mode,min_sample", [
    (4, 1 << 20, [64] * 5, 2, 512),            # BASELINE config 2, the bench's reset mode: 320 steps
    (4, 1 << 20, [64] * 3 + [63, 64], 1, 512),  # same-step reset (BatchedSkyjoEnv's default)
    (8, 1 << 24, [64] * 8, 2, 512),            # BASELINE config 3 at full size: 512 steps
])
def test_bench_configs_match_oracle_on_sampled_envs_at_full_size(N, B, chunks, mode, min_sample):
    ids = sample_blocks(B, n_blocks=16, width=32, ranges=4)
    assert len(ids) >= min_sample
    env = _env(num_envs=B, num_players=N, score_penalty=2.0, mean_reward=1.0, reward_refunded=0.0, seed=N,
               auto_reset=mode, first_global_env_id=0)
    steps, ended, st = chunked_rollout(None, N, False, 2.0, 1.0, 0.0, B, chunks, reset_mode=mode, ids=ids, seed=N,
                                       env=env)
    T = sum(chunks)
    assert ended > len(ids) and steps > 0.98 * len(ids) * T
    # whole-batch bookkeeping: every lockstep slot of every env is an env-step or (mode 2) a counted reset slot
    if mode == 1:
        assert st["steps"] == B * T
    else:
        waiting = int((env.done_code != 0).sum())
        assert st["steps"] == B * T - (st["episodes"] - waiting)
    if N == 8:
        assert st["reshuffles"] > 0 and st["refunds"] > 0
    assert st["illegal"] == 0 and st["truncated"] == 0


def test_seed_after_random_stepping_restarts_the_same_games():
    # skyjo_seed must wait for the refill deal still running on the library's side stream before it zeroes the
    # episode counters: seed(s) after random stepping == a fresh env seeded with s, bit for bit
    kw = dict(num_envs=5000, num_players=4, auto_reset=2)
    a = _env(seed=1, **kw)
    a.reset()
    for n in (64, 3, 64):
        a.step_random(n)
        a.seed(99)
        b = _env(seed=99, **kw)
        b.reset()
        for _ in range(3):
            assert torch.equal(a.observations, b.observations) and torch.equal(a.agent_selection, b.agent_selection)
            a.step_random(50)
            b.step_random(50)
        assert torch.equal(a.observations, b.observations) and torch.equal(a.rewards, b.rewards)
        va, vb = a.export(0, 64), b.export(0, 64)
        assert all(x.episode == y.episode and np.array_equal(x.players_cards, y.players_cards) for x, y in zip(va, vb))
        b.close()
    a.check()


def test_side_stream_stats_allreduce_snapshots_iteration_boundaries():
    # skyjo_stats_allreduce_async (one rank: the collective degenerates to a copy on the side stream): the vector
    # that lands is the snapshot at the call, whatever is queued behind it, and the two buffers alternate
    env = _env(num_envs=1 << 16, num_players=4, seed=3, auto_reset=2)
    env.reset()
    want, got = [], []
    for it in range(6):
        env.step_random(64)
        want.append(env.stats_tensor().clone())
        got.append(env.stats_allreduce_async(None))
        env.step_random(17)                       # queued behind the snapshot: must not leak into it
        if it % 2 == 1:
            env.stats_allreduce_wait()
            torch.cuda.synchronize()
            assert torch.equal(got[-2], want[-2]) and torch.equal(got[-1], want[-1])
            assert got[-1].data_ptr() != got[-2].data_ptr()
    assert int(want[-1][16]) > int(want[0][16]) > 0
    env.check()


@pytest.mark.parametrize("N,B,mode,T", [(4, 1 << 18, 1, 300), (4, 1 << 18, 2, 300), (3, 1 << 17, 1, 260)])
def test_external_action_path_matches_oracle_on_sampled_envs_at_scale(N, B, mode, T):
    """The path of BASELINE config 4 and of skyjo_step_host -- skyjo_step with actions from outside, one refill deal
    per step -- at 2^18 envs: actions drawn on the device from the mask (as a policy would), 0.3 % of the sampled
    envs given an illegal action per step; every step obs / mask / agent / done / float64 rewards of 384 sampled envs
    against OracleBatch (SkyjoGame.act + TerminateIllegalWrapper semantics + the reset schedule)."""
    from parity_util import OracleBatch, to_np
    ids = sample_blocks(B, n_blocks=8, width=32, ranges=4)[:384]
    seed, first = 77 + N, 10_000_000
    env = _env(num_envs=B, num_players=N, score_penalty=2.0, mean_reward=1.0, reward_refunded=0.25, seed=seed,
               auto_reset=mode, first_global_env_id=first)
    env.reset()
    ref = OracleBatch(N, False, 2.0, 1.0, 0.25, seed, len(ids), first, mode, 0, env_ids=[first + int(i) for i in ids])
    tid = torch.as_tensor(ids, device=env.device)
    rng = np.random.default_rng(5)
    g = torch.Generator(device=env.device)
    g.manual_seed(3)
    ended = illegal = 0
    for t in range(T):
        obs, mask, agent = ref.publish()
        np.testing.assert_array_equal(to_np(env.observations[tid]), obs, err_msg=f"obs at step {t}")
        np.testing.assert_array_equal(to_np(env.action_mask[tid]), mask, err_msg=f"mask at step {t}")
        np.testing.assert_array_equal(to_np(env.agent_selection[tid]), agent, err_msg=f"agent at step {t}")
        a = torch.multinomial(env.action_mask.float(), 1, generator=g).squeeze(1).to(torch.uint8)
        bad = np.flatnonzero(rng.random(len(ids)) < 0.003)
        for i in bad:                                  # an action of the other phase, or out of range
            a[int(ids[i])] = int(rng.choice(np.flatnonzero(mask[i] == 0))) if rng.random() < 0.7 else int(rng.integers(26, 200))
        act = to_np(a[tid]).astype(np.int64)
        done, reward, score = ref.step(act, mask, agent)
        env.step(a)
        np.testing.assert_array_equal(to_np(env.done_code[tid]), done, err_msg=f"done at step {t}")
        assert to_np(env.rewards[tid]).tobytes() == reward.tobytes(), f"rewards at step {t}"
        scored = (done == 1) & ~np.isnan(score[:, 0])
        assert to_np(env.final_scores[tid])[scored].tobytes() == score[scored].tobytes(), f"scores at step {t}"
        ended += int((done == 1).sum())
        illegal += int((done == 2).sum())
    env.check()
    assert ended > len(ids) // 2 and illegal > 0, (ended, illegal)
    st = env.stats()
    assert st["illegal"] >= illegal and st["episodes"] > B // 2


@pytest.mark.parametrize("N,B,ranges,mode", [(4, 5000, 4, "next_step"), (2, 3000, 1, True), (8, 2500, 3, "next_step")])
def test_graph_replayed_calls_play_the_same_games_as_direct_launches(N, B, ranges, mode, monkeypatch):
    """skyjo_step_random replays a call shape from a CUDA graph from its third occurrence on (seen, captured,
    replayed; the kernels then read the lockstep counter from device memory).  An env created with SKYJO_NO_GRAPH=1
    launches directly: both must publish identical buffers after every call of a sequence that mixes replayed
    shapes, shapes that cannot be replayed (an odd number of refill windows), single steps, a reseed and a restored
    checkpoint."""
    import torch
    from skyjo_rl_b200 import BatchedSkyjoEnv
    monkeypatch.setenv("SKYJO_NO_GRAPH", "1")
    direct = BatchedSkyjoEnv(num_envs=B, num_players=N, seed=4, device="cuda:0", auto_reset=mode)
    monkeypatch.delenv("SKYJO_NO_GRAPH")
    graph = BatchedSkyjoEnv(num_envs=B, num_players=N, seed=4, device="cuda:0", auto_reset=mode)
    for e in (direct, graph):
        e.set_env_ranges(ranges)
        e.reset()

    def same(tag):
        for name in ("observations", "action_mask", "agent_selection", "done_code", "rewards"):
            assert torch.equal(getattr(direct, name), getattr(graph, name)), (tag, name)
        assert direct.step_count == graph.step_count

    saved = None
    for i, n in enumerate([64, 64, 64, 64, 5, 64, 16, 1, 64, 128, 128, 128, 64]):
        direct.step_random(n)
        graph.step_random(n)
        same((i, n))
        if i == 5:
            saved = graph.state_dict()
    assert direct.stats() == graph.stats()
    graph.load_state_dict(saved)            # back to the state after call 5: the device counter is stale now
    direct.load_state_dict(saved)
    for n in (64, 64, 7, 64):
        direct.step_random(n)
        graph.step_random(n)
        same(("restored", n))
    for e in (direct, graph):
        e.seed(99)                          # drops the graphs (the seed is baked into their kernel parameters)
    for n in (64, 64, 64, 64):
        direct.step_random(n)
        graph.step_random(n)
        same(("reseeded", n))
    direct.check()
    graph.check()
    assert torch.equal(direct.state_dict()["state"], graph.state_dict()["state"])
