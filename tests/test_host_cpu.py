"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the header
declares, validates configs, fails loudly without a device, and its host RNG twins agree with
the oracle's independent restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from skyjo_rl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    header = open(os.path.join(ROOT, "include", "skyjo_b200.h")).read()
    declared = set(re.findall(r"\b(skyjo_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.skyjo_abi_version() == 1


def test_struct_sizes_match_header():
    assert C.sizeof(_lib.SkyjoConfig) == 40
    assert C.sizeof(_lib.SkyjoOutputs) == 48
    assert C.sizeof(_lib.SkyjoEnvDebug) == 536


def test_obs_len_and_state_bytes():
    L = _lib.load()
    for N in range(1, 13):
        cfg = _lib.SkyjoConfig(N, 0, 2.0, 1.0, 0.0, 1, 0)
        assert L.skyjo_obs_len(C.byref(cfg)) == 19 + 12 * N          # reference skyjo.py:43-45
        cfg.observe_other_player_indirect = 1
        assert L.skyjo_obs_len(C.byref(cfg)) == 31
        nbytes = L.skyjo_state_bytes(C.byref(cfg), 1000)
        assert nbytes > 0 and nbytes % 256 == 0
    for bad in (0, 13, -1):
        cfg = _lib.SkyjoConfig(bad, 0, 2.0, 1.0, 0.0, 1, 0)
        assert L.skyjo_obs_len(C.byref(cfg)) == -1                  # skyjo.py:24-26
        assert L.skyjo_state_bytes(C.byref(cfg), 10) == -1


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    L = _lib.load()
    cfg = _lib.SkyjoConfig(4, 0, 2.0, 1.0, 0.0, 1, 0)
    h = C.c_void_p()
    rc = L.skyjo_create(C.byref(cfg), 0, 128, 0, 0, None, 0, C.byref(h))
    assert rc == 4 and b"no CPU fallback" in L.skyjo_last_error()
    from skyjo_rl_b200 import BatchedSkyjoEnv
    with pytest.raises(RuntimeError):
        BatchedSkyjoEnv(num_envs=8, num_players=2)


def test_host_rng_twins_match_oracle_restatement():
    L = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(20):
        seed, env, ep = int(rng.integers(2**62)), int(rng.integers(2**40)), int(rng.integers(2**31))
        d = (C.c_int8 * 150)()
        L.skyjo_host_deck(seed, env, ep, d)
        assert list(d) == O.rng_deck(seed, env, ep).tolist()
        f = (C.c_uint8 * 24)()
        L.skyjo_host_flips(seed, env, ep, 12, f)
        assert list(f) == O.rng_flips(seed, env, ep, 12).reshape(-1).tolist()
        mask = (rng.random(26) < 0.5).astype(np.int8)
        mask[int(rng.integers(26))] = 1
        bits = int(sum(1 << i for i in range(26) if mask[i]))
        t = int(rng.integers(2**40))
        assert L.skyjo_host_policy(seed, env, t, bits) == O.rng_policy(seed, env, t, mask)
    out = (C.c_uint32 * 4)()
    L.skyjo_host_philox4x32_10((C.c_uint32 * 4)(0, 0, 0, 0), (C.c_uint32 * 2)(0, 0), out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_spaces_match_reference_bounds():
    from skyjo_rl_b200.spaces import Box, Discrete
    b = Box(low=-24, high=127, shape=(67,), dtype=np.int8)   # reference skyjo_env.py:129-134
    assert b.shape == (67,) and b.low.min() == -24 and b.high.max() == 127
    assert Discrete(26).n == 26                               # skyjo_env.py:146-151


def test_host_expand_packed_matches_numpy_unpack():
    """Host half of skyjo_step_host's wire format (csrc/skyjo_hostio.cuh): one uint32 per env ->
    mask[26] + agent + done."""
    L = _lib.load()
    rng = np.random.default_rng(5)
    n = 1000
    packed = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    packed[:4] = [0, 0xFFFFFFFF, 3 << 24, (1 << 26) - 1]
    mask = np.full((n, 26), 7, np.int8)
    agent = np.full(n, 7, np.int8)
    done = np.full(n, 7, np.uint8)
    L.skyjo_host_expand_packed(packed.ctypes.data, n, mask.ctypes.data, agent.ctypes.data, done.ctypes.data)
    exp_mask = ((packed[:, None] >> np.arange(26, dtype=np.uint32)[None, :]) & 1).astype(np.int8)
    np.testing.assert_array_equal(mask, exp_mask)
    np.testing.assert_array_equal(agent, (packed >> 28).astype(np.int8))
    np.testing.assert_array_equal(done, ((packed >> 26) & 3).astype(np.uint8))
    L.skyjo_host_expand_packed(packed.ctypes.data, n, None, None, done.ctypes.data)   # null outputs are skipped


def test_integration_md_stub_matches_the_binding():
    # the reference-side ctypes stub documented in INTEGRATION.md loads the library and declares the same
    # structures and signatures as the shipped binding (no compute call: there is no GPU here)
    import ctypes as C
    from integration_stub import load_stub
    from skyjo_rl_b200 import _lib
    m = load_stub()
    for name in ("SkyjoConfig", "SkyjoOutputs"):
        a, b = getattr(m, name), getattr(_lib, name)
        assert C.sizeof(a) == C.sizeof(b)
        assert [(f[0], f[1]) for f in a._fields_] == [(f[0], f[1]) for f in b._fields_]
    L = _lib.load()
    for fn in ("skyjo_state_bytes", "skyjo_create", "skyjo_bind_outputs", "skyjo_reset", "skyjo_step"):
        got, exp = getattr(m.L, fn).argtypes, getattr(L, fn).argtypes
        assert len(got) == len(exp) and all(C.sizeof(x) == C.sizeof(y) for x, y in zip(got, exp)), fn
    cfg = m.SkyjoConfig(4, 0, 2.0, 1.0, 0.0, 1, 0)
    assert m.L.skyjo_state_bytes(C.byref(cfg), 1 << 10) == L.skyjo_state_bytes(C.byref(_lib.SkyjoConfig(4, 0, 2.0, 1.0, 0.0, 1, 0)), 1 << 10)


def _rec_buf(n, RB):
    """record array with the 256 readable bytes behind it that the gather version of the expansion may touch"""
    raw = np.zeros(n * RB + 256, dtype=np.uint8)
    return raw[:n * RB].reshape(n, RB)


def _reference_observation_rows():
    """every observation row the unmodified reference produced for the fixtures, grouped by row length"""
    import os
    from conftest import GOLDEN_DIR, golden_files
    rows = {}
    for f in golden_files():
        z = np.load(os.path.join(GOLDEN_DIR, f))
        for k in ("obs", "obs_other", "final_obs"):
            a = z[k].astype(np.int8)
            rows.setdefault(a.shape[1], []).append(a)
    z = np.load(os.path.join(GOLDEN_DIR, "env_trace.npz"))
    for k in z.files:
        if k.endswith("/obs"):
            rows.setdefault(z[k].shape[1], []).append(z[k].astype(np.int8))
    return {D: np.ascontiguousarray(np.concatenate(v)) for D, v in rows.items()}


def test_compact_observation_records_round_trip_on_every_reference_observation():
    # wire format of skyjo_step_host (csrc/skyjo_hostio.cuh): expand(pack(row)) == row for every observation the
    # reference produced in the fixtures (removed columns, dense decks, N = 1..12, both observation modes),
    # through the 16-byte-shuffle expansion and through the portable one
    from skyjo_rl_b200 import _lib
    L = _lib.load()
    total = removed = 0
    for D, obs in sorted(_reference_observation_rows().items()):
        R = (D - 19) // 12
        RB = L.skyjo_host_obs_record_bytes(D)
        assert RB == 12 + 6 * R + (R + 1) // 2 and RB < D
        n = obs.shape[0]
        rec = _rec_buf(n, RB)
        assert L.skyjo_host_pack_obs(obs.ctypes.data, n, D, rec.ctypes.data) == 0
        for portable in (0, 1, 2, 3):
            out = np.full((n + 1, D), 99, dtype=np.int8)           # one guard row behind the last one
            L.skyjo_host_expand_obs(rec.ctypes.data, n, D, out.ctypes.data, portable)
            np.testing.assert_array_equal(out[:n], obs)
            assert (out[n] == 99).all()
        total += n
        removed += int((obs == -14).sum())
    assert total > 20000 and removed > 3000          # the fixtures hold removed columns
    assert L.skyjo_host_obs_record_bytes(67) == 38 and L.skyjo_host_obs_record_bytes(30) == -1


def test_compact_observation_records_flag_rows_they_cannot_hold():
    from skyjo_rl_b200 import _lib
    L = _lib.load()
    D = 67
    good = np.zeros((1, D), dtype=np.int8)
    good[0, 17] = -3
    good[0, 18] = 15
    good[0, 19:] = 15
    rec = _rec_buf(1, L.skyjo_host_obs_record_bytes(D))
    assert L.skyjo_host_pack_obs(good.ctypes.data, 1, D, rec.ctypes.data) == 0
    for pos, val in ((19, 13), (25, -3), (2, 16), (5, -1), (17, 13), (18, 14), (1, 16), (40, -14)):
        bad = good.copy()
        bad[0, pos] = val           # a card / bin / top / hand outside the symbol sets, a lone -14
        assert L.skyjo_host_pack_obs(bad.ctypes.data, 1, D, rec.ctypes.data) == 1, (pos, val)
    ok = good.copy()
    ok[0, 4] = 100                  # the value-0 bin travels as a whole byte
    ok[0, 22:25] = -14              # a removed column
    assert L.skyjo_host_pack_obs(ok.ctypes.data, 1, D, rec.ctypes.data) == 0
    out = np.zeros_like(ok)
    L.skyjo_host_expand_obs(rec.ctypes.data, 1, D, out.ctypes.data, 0)
    np.testing.assert_array_equal(out, ok)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 7, 8, 12])
def test_compact_observation_records_round_trip_on_random_rows(R):
    # every row length 19 + 12 R (R = 1 is also the indirect mode's row), random symbols from the sets a row can hold
    from skyjo_rl_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(R)
    D, n = 19 + 12 * R, 3001
    obs = np.zeros((n, D), dtype=np.int8)
    obs[:, 0] = rng.integers(-24, 128, n)
    obs[:, 1] = rng.integers(0, 13, n)
    obs[:, 2:17] = rng.integers(0, 16, (n, 15))
    obs[:, 4] = rng.integers(0, 128, n)                      # the value-0 bin is a whole byte
    obs[:, 17] = rng.integers(-3, 13, n)
    obs[:, 18] = rng.choice(np.r_[np.arange(-2, 13), 15], n)
    cards = rng.choice(np.r_[np.arange(-2, 13), 15, 15, 15], (n, R, 4, 3)).astype(np.int8)
    cards[rng.random((n, R, 4)) < 0.15] = -14                # removed columns
    obs[:, 19:] = cards.reshape(n, 12 * R)
    RB = L.skyjo_host_obs_record_bytes(D)
    rec = _rec_buf(n, RB)
    assert L.skyjo_host_pack_obs(obs.ctypes.data, n, D, rec.ctypes.data) == 0
    for portable in (0, 1, 2, 3):
        out = np.full((n + 1, D), 77, dtype=np.int8)
        L.skyjo_host_expand_obs(rec.ctypes.data, n, D, out.ctypes.data, portable)
        np.testing.assert_array_equal(out[:n], obs)
        assert (out[n] == 77).all()


def test_empty_and_invalid_inputs_are_errors_not_crashes():
    # an empty batch, a null config, zero-length host helpers: error codes / no-ops, never a crash (no GPU needed:
    # these are rejected before the device is touched)
    L = _lib.load()
    cfg = _lib.SkyjoConfig(4, 0, 2.0, 1.0, 0.0, 1, 0)
    h = C.c_void_p()
    assert L.skyjo_state_bytes(C.byref(cfg), 0) == -1
    assert L.skyjo_create(C.byref(cfg), 0, 0, 0, 0, None, 0, C.byref(h)) == 1 and b"num_envs" in L.skyjo_last_error()
    assert L.skyjo_create(C.byref(cfg), 0, 8, 0, -5, None, 0, C.byref(h)) == 1
    assert L.skyjo_create(None, 0, 8, 0, 0, None, 0, C.byref(h)) == 1
    bad = _lib.SkyjoConfig(4, 0, 2.0, 1.0, 0.0, 7, 0)                       # auto_reset outside SKYJO_RESET_*
    assert L.skyjo_create(C.byref(bad), 0, 8, 0, 0, None, 0, C.byref(h)) == 1
    assert L.skyjo_step(None, None, 0, None) != 0 and L.skyjo_reset(None, None) != 0
    assert L.skyjo_stats_allreduce(None, None, None, None) == 1
    assert L.skyjo_host_wire_bytes(None) == -1 and L.skyjo_set_host_wire(None, 0) == 1
    assert L.skyjo_host_pack_obs(None, 0, 67, None) == 0                    # zero rows: nothing touched
    L.skyjo_host_expand_obs(None, 0, 67, None, 0)
    L.skyjo_host_expand_packed(None, 0, None, None, None)
    assert L.skyjo_host_obs_record_bytes(19) == -1 and L.skyjo_host_obs_record_bytes(31) == 19
    assert L.skyjo_host_policy(0, 0, 0, 0) == -1                            # no legal action


def _host_reshuffle(L, seed, env, ep, q, pile):
    buf = np.ascontiguousarray(pile, dtype=np.int8).copy()
    rc = L.skyjo_host_reshuffle(seed, env, ep, q, buf.ctypes.data, len(buf))
    assert rc == 0, L.skyjo_last_error()
    return buf


def test_host_reshuffle_twin_matches_oracle_and_python_restatement():
    # the exported twin of the in-game reshuffle rule (skyjo.py:127-138 stand-in) against the oracle's and the
    # pure-Python statement: same permutation, multiset preserved, 7-bit reshuffle index, bad piles rejected
    import rng_twin
    L = _lib.load()
    rng = np.random.default_rng(5)
    for _ in range(120):
        n = int(rng.integers(1, 150))
        pile = rng.integers(-2, 13, n).astype(np.int8)
        seed, env, ep = int(rng.integers(2**62)), int(rng.integers(2**40)), int(rng.integers(2**31))
        q = int(rng.integers(300))
        out = _host_reshuffle(L, seed, env, ep, q, pile)
        assert out.tolist() == O.rng_reshuffle(seed, env, ep, q, pile).tolist()
        assert out.tolist() == rng_twin.reshuffle(seed, env, ep, q, pile.tolist())
        assert sorted(out.tolist()) == sorted(pile.tolist())
    bad = np.array([1, 2, 13], dtype=np.int8)
    assert L.skyjo_host_reshuffle(1, 2, 3, 0, bad.ctypes.data, 3) != 0
    assert L.skyjo_host_reshuffle(1, 2, 3, 0, bad.ctypes.data, 0) != 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/rlskyjo"), reason="the reference is only mounted in the build container")
def test_live_reference_patched_with_host_reshuffle_equals_oracle():
    """What a third party does: the UNMODIFIED reference with SkyjoGame._reshuffle_discard_pile
    replaced by a function that calls skyjo_host_reshuffle plays the games of the oracle driven by its own
    restatement of the rule -- long 9- and 11-player games with several in-game reshuffles each."""
    pytest.importorskip("numba")
    import sys
    sys.path.insert(0, "/root/reference")
    try:
        from rlskyjo.game.skyjo import SkyjoGame
    finally:
        sys.path.remove("/root/reference")
    L = _lib.load()
    ctx = {}

    def patched(old_pile):          # (int8[L]) -> (list[L-1], list[1]), skyjo.py:127-138
        if not ctx:
            lst = [int(x) for x in old_pile]
        else:
            lst = _host_reshuffle(L, ctx["seed"], ctx["env"], 0, ctx["q"], np.asarray(old_pile, np.int8)).tolist()
            ctx["q"] += 1
        top = lst.pop()
        return lst, [top]

    orig = SkyjoGame.__dict__["_reshuffle_discard_pile"]
    SkyjoGame._reshuffle_discard_pile = staticmethod(patched)
    try:
        rng = np.random.default_rng(77)
        total_reshuffles = 0
        for env, N in enumerate([9, 11, 8, 12]):
            seed = 4242 + env
            deck = O.rng_deck(seed, env, 0)
            flips = O.rng_flips(seed, env, 0, N)
            ctx.clear()
            g = SkyjoGame(num_players=N, score_penalty=2.0, observe_other_player_indirect=False)
            g.players_cards = deck[:12 * N].reshape(N, 12).astype(np.int8).copy()        # SURVEY 9.8 injection
            masked = np.full((N, 12), 2, np.int8)
            for p in range(N):
                masked[p, flips[p]] = 1
            g.players_masked = masked
            rest = [int(x) for x in deck[12 * N:]]
            g.discard_pile, g.drawpile = [rest[-1]], rest[:-1]
            g._reset_start_player()
            ctx.update(seed=seed, env=env, q=0)
            og = O.OracleGame(N, 2.0, False)
            og.reset_rng(seed, env, 0)
            while not g.is_terminated:
                pid = g.expected_action[0]
                assert pid == og.expected_action[0]
                obs, mask = g.collect_observation(pid)
                o2, m2 = og.collect_observation(pid)
                np.testing.assert_array_equal(obs, o2)
                np.testing.assert_array_equal(mask, m2)
                legal = np.flatnonzero(mask)
                a = 24 if mask[24] and rng.random() < 0.8 else int(rng.choice(legal))   # favour the draw pile
                assert g.act(pid, a) == og.act(pid, a)
            assert [float(x) for x in g.game_metrics["final_score"]] == og.game_metrics["final_score"]
            total_reshuffles += og.n_reshuffles
            assert ctx["q"] == og.n_reshuffles
        assert total_reshuffles >= 4
    finally:
        SkyjoGame._reshuffle_discard_pile = orig


def _aligned(nbytes, dtype, fill, align=64, offset=0):
    """numpy view of `nbytes` bytes whose address is `offset` past a multiple of `align`, inside a guarded buffer"""
    raw = np.full(nbytes + 2 * align + 64, fill, dtype=np.uint8)
    start = (-raw.ctypes.data) % align + offset
    return raw, raw[start:start + nbytes].view(dtype)


@pytest.mark.parametrize("n,offset", [(64, 0), (1000, 0), (4096 + 77, 0), (130, 16), (5000, 1), (63, 0), (1, 0)])
def test_wide_host_expansions_equal_the_portable_ones(n, offset):
    """The AVX-512 / streaming-store expansions of skyjo_step_host's wire format (csrc/skyjo_hostsimd.cpp) on aligned
    and misaligned caller buffers, whole and partial 64-env groups, against the scalar code and numpy; nothing
    outside the destination is touched."""
    L = _lib.load()
    assert L.skyjo_host_simd_level() in (0, 2, 3)
    rng = np.random.default_rng(n + offset)
    packed = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
    raw_m, mask = _aligned(n * 26, np.int8, 7, offset=offset)
    raw_a, agent = _aligned(n, np.int8, 7, offset=offset)
    raw_d, done = _aligned(n, np.uint8, 7, offset=offset)
    L.skyjo_host_expand_packed(packed.ctypes.data, n, mask.ctypes.data, agent.ctypes.data, done.ctypes.data)
    exp_mask = ((packed[:, None] >> np.arange(26, dtype=np.uint32)[None, :]) & 1).astype(np.int8)
    np.testing.assert_array_equal(mask.reshape(n, 26), exp_mask)
    np.testing.assert_array_equal(agent, (packed >> 28).astype(np.int8))
    np.testing.assert_array_equal(done, ((packed >> 26) & 3).astype(np.uint8))
    for raw, view in ((raw_m, mask), (raw_a, agent), (raw_d, done)):
        lo = view.ctypes.data - raw.ctypes.data
        assert (raw[:lo] == 7).all() and (raw[lo + view.nbytes:] == 7).all()      # guards intact
    for R in (1, 2, 3, 4, 8, 12):
        D = 19 + 12 * R
        obs = np.zeros((n, D), dtype=np.int8)
        obs[:, 0] = rng.integers(-24, 128, n)
        obs[:, 1] = rng.integers(0, 13, n)
        obs[:, 2:17] = rng.integers(0, 16, (n, 15))
        obs[:, 4] = rng.integers(0, 128, n)
        obs[:, 17] = rng.integers(-3, 13, n)
        obs[:, 18] = rng.choice(np.r_[np.arange(-2, 13), 15], n)
        cards = rng.choice(np.r_[np.arange(-2, 13), 15, 15], (n, R, 4, 3)).astype(np.int8)
        cards[rng.random((n, R, 4)) < 0.1] = -14
        obs[:, 19:] = cards.reshape(n, 12 * R)
        RB = L.skyjo_host_obs_record_bytes(D)
        rec = _rec_buf(n, RB)
        assert L.skyjo_host_pack_obs(obs.ctypes.data, n, D, rec.ctypes.data) == 0
        for mode in (0, 1, 2, 3):
            raw_o, out = _aligned(n * D, np.int8, 99, offset=offset)
            L.skyjo_host_expand_obs(rec.ctypes.data, n, D, out.ctypes.data, mode)
            np.testing.assert_array_equal(out.reshape(n, D), obs, err_msg=f"mode {mode} R {R}")
            lo = out.ctypes.data - raw_o.ctypes.data
            assert (raw_o[:lo] == 99).all()
            if mode in (0, 3):   # the streaming versions write whole groups only: the bytes behind the last row stay untouched
                assert (raw_o[lo + out.nbytes + 16:] == 99).all()


def test_round2_entry_points_validate_their_arguments():
    # the policy kernel's weight image, the side-stream all-reduce, the env-range knob, the host twins added in round 2:
    # sizes and argument checks that need no device
    L = _lib.load()
    assert L.skyjo_policy_packed_bytes() == 96 * 256 * 2 + 256 * 256 * 2 + 256 * 32 * 2 + 2 * 256 * 4 + 32 * 4
    assert L.skyjo_policy_trace_len() == 2048
    one = C.c_void_p(16)
    assert L.skyjo_policy_pack(67, 26, None, None, None, None, None, None, None, None) == 1
    assert L.skyjo_policy_pack(97, 26, one, one, one, one, one, one, one, None) == 1 and b"96 bytes" in L.skyjo_last_error()
    assert L.skyjo_policy_pack(67, 27, one, one, one, one, one, one, one, None) == 1
    assert L.skyjo_policy_pack(67, 26, one, one, one, one, one, one, C.c_void_p(8), None) == 1      # misaligned image
    assert L.skyjo_policy_sample(None, one, 0, one, None, None, None, None) == 1
    assert L.skyjo_policy_value(None, one, one, None) == 1 and L.skyjo_policy_value(None, one, None, None) == 1
    assert L.skyjo_policy_debug(None, one, None, None, None, None) == 1
    assert L.skyjo_stats_allreduce_async(None, None, None, None) == 1 and L.skyjo_stats_allreduce_wait(None, None) == 1
    assert L.skyjo_set_env_ranges(None, 2) == 1 and L.skyjo_host_wire_share(None) == -1
    assert L.skyjo_set_host_wire(None, 2) == 1
    assert L.skyjo_graph_replay_count(None) == -1 and L.skyjo_launch_count(None) == -1
    assert L.skyjo_host_simd_level() in (0, 2, 3)
