"""ctypes binding of oracle/libskyjo_oracle.so -- TEST INFRASTRUCTURE ONLY.

`OracleGame` mirrors the attribute names of the reference's `SkyjoGame`
(/root/reference/rlskyjo/game/skyjo.py:19-49) so that parity tests read like the
reference's own loops (`expected_action`, `collect_observation`, `act`, ...).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libskyjo_oracle.so")

SK_MAX_PLAYERS = 12
SK_PILE_CAP = 512

_SHUFFLE_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int8), C.c_int)


class _Game(C.Structure):
    _fields_ = [
        ("num_players", C.c_int),
        ("score_penalty", C.c_double),
        ("indirect", C.c_int),
        ("cards", (C.c_int8 * 12) * SK_MAX_PLAYERS),
        ("masked", (C.c_int8 * 12) * SK_MAX_PLAYERS),
        ("drawpile", C.c_int8 * SK_PILE_CAP),
        ("n_draw", C.c_int),
        ("discard", C.c_int8 * SK_PILE_CAP),
        ("n_disc", C.c_int),
        ("hand", C.c_int),
        ("is_terminated", C.c_int),
        ("exp_player", C.c_int),
        ("exp_phase", C.c_int),
        ("num_refunded", C.c_int * SK_MAX_PLAYERS),
        ("num_placed", C.c_int * SK_MAX_PLAYERS),
        ("has_final_score", C.c_int),
        ("final_score", C.c_double * SK_MAX_PLAYERS),
        ("n_reshuffles", C.c_int),
        ("shuffle", _SHUFFLE_FN),
        ("shuffle_ctx", C.c_void_p),
    ]


class _ShuffleCtx(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("env", C.c_uint64), ("episode", C.c_uint32), ("q", C.c_uint32)]


def build_oracle(force=False):
    """Compile the C restatement (gcc) if the shared object is missing or stale."""
    src = os.path.join(_HERE, "skyjo_oracle.c")
    hdr = os.path.join(_HERE, "skyjo_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libskyjo_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_SO)
        L.sk_obs_len.restype = C.c_int
        L.sk_obs_len.argtypes = [C.c_int, C.c_int]
        L.sk_init.argtypes = [C.POINTER(_Game), C.c_int, C.c_double, C.c_int]
        L.sk_reset_injected.argtypes = [C.POINTER(_Game), C.c_void_p, C.c_void_p]
        L.sk_collect_observation.argtypes = [C.POINTER(_Game), C.c_int, C.c_void_p, C.c_void_p]
        L.sk_act.argtypes = [C.POINTER(_Game), C.c_int, C.c_int]
        L.sk_evaluate_game.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p]
        L.sk_evaluate_game.restype = None
        L.sk_calc_final_rewards.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.sk_calc_final_rewards.restype = None
        L.sk_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sk_philox4x32_10.restype = None
        L.sk_rng_deck.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.sk_rng_deck.restype = None
        L.sk_rng_flips.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p]
        L.sk_rng_flips.restype = None
        L.sk_rng_policy.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.sk_rng_policy.restype = C.c_int
        L.sk_rng_reshuffle.argtypes = [C.c_void_p, C.POINTER(C.c_int8), C.c_int]
        L.sk_rng_reshuffle.restype = None
        L.sk_reset_rng.argtypes = [C.POINTER(_Game), C.POINTER(_ShuffleCtx)]
        L.sk_rollout.argtypes = [C.c_int, C.c_double, C.c_int, C.c_uint64, C.c_uint64, C.c_int64,
                                 C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
        L.sk_rollout.restype = C.c_int64
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleGame:
    """One game on the C oracle, with the reference's attribute names."""

    def __init__(self, num_players=3, score_penalty=2.0, observe_other_player_indirect=False):
        self._L = lib()
        self._g = _Game()
        rc = self._L.sk_init(C.byref(self._g), num_players, float(score_penalty), int(observe_other_player_indirect))
        if rc != 0:
            raise AssertionError("Skyjo can be played from 1 up to 12 players")
        self.num_players = num_players
        self.score_penalty = score_penalty
        self.observe_other_player_indirect = observe_other_player_indirect
        self.obs_shape = (self._L.sk_obs_len(num_players, int(observe_other_player_indirect)),)
        self.action_mask_shape = (26,)
        self._ctx = None
        self._py_shuffle = None

    # -- dealing ------------------------------------------------------------------------
    def reset_injected(self, deck, flips):
        deck = np.ascontiguousarray(deck, dtype=np.int8)
        flips = np.ascontiguousarray(flips, dtype=np.uint8)
        assert deck.shape == (150,) and flips.shape == (self.num_players, 2)
        rc = self._L.sk_reset_injected(C.byref(self._g), _ptr(deck), _ptr(flips))
        assert rc == 0, f"bad injected deal ({rc})"

    def reset_rng(self, seed, env, episode):
        """Deal from the twins of the product RNG and wire the twin in-game reshuffle."""
        self._ctx = _ShuffleCtx(seed, env, episode, 0)
        rc = self._L.sk_reset_rng(C.byref(self._g), C.byref(self._ctx))
        assert rc == 0

    def set_rng_reshuffle(self, seed, env, episode):
        """Keep the injected deal but use the product's keyed rule for in-game reshuffles."""
        self._ctx = _ShuffleCtx(seed, env, episode, 0)
        fn = C.cast(self._L.sk_rng_reshuffle, _SHUFFLE_FN)
        self._g.shuffle = fn
        self._g.shuffle_ctx = C.cast(C.pointer(self._ctx), C.c_void_p)

    def set_shuffle(self, fn):
        """fn(np.int8[len]) -> permuted np.int8[len]; stands in for np.random.shuffle."""
        def tramp(_ctx, pile, n):
            arr = np.ctypeslib.as_array(pile, shape=(n,))
            arr[:] = fn(arr.copy())
        self._py_shuffle = _SHUFFLE_FN(tramp)
        self._g.shuffle = self._py_shuffle
        self._g.shuffle_ctx = None

    # -- reference surface ----------------------------------------------------------------
    @property
    def expected_action(self):
        return [self._g.exp_player, "draw" if self._g.exp_phase == 0 else "place"]

    @property
    def is_terminated(self):
        return bool(self._g.is_terminated)

    @property
    def hand_card(self):
        return int(self._g.hand)

    @property
    def players_cards(self):
        return np.array([list(self._g.cards[p]) for p in range(self.num_players)], dtype=np.int8)

    @property
    def players_masked(self):
        return np.array([list(self._g.masked[p]) for p in range(self.num_players)], dtype=np.int8)

    @property
    def drawpile(self):
        return [int(x) for x in self._g.drawpile[: self._g.n_draw]]

    @property
    def discard_pile(self):
        return [int(x) for x in self._g.discard[: self._g.n_disc]]

    @property
    def n_reshuffles(self):
        return int(self._g.n_reshuffles)

    @property
    def game_metrics(self):
        N = self.num_players
        return {
            "num_refunded": [int(x) for x in self._g.num_refunded[:N]],
            "num_placed": [int(x) for x in self._g.num_placed[:N]],
            "final_score": [float(x) for x in self._g.final_score[:N]] if self._g.has_final_score else False,
        }

    def collect_observation(self, player_id):
        obs = np.empty(self.obs_shape, dtype=np.int8)
        mask = np.empty(26, dtype=np.int8)
        rc = self._L.sk_collect_observation(C.byref(self._g), player_id, _ptr(obs), _ptr(mask))
        assert rc == 0
        return obs, mask

    def act(self, player_id, action_int):
        rc = self._L.sk_act(C.byref(self._g), int(player_id), int(action_int))
        if rc < 0:
            raise AssertionError(f"ILLEGAL ACTION (oracle code {rc})")
        return rc == 1

    def final_rewards(self, mean_reward=1.0, reward_refunded=0.0):
        N = self.num_players
        fs = np.array(self.game_metrics["final_score"], dtype=np.float64)
        nr = np.array(self.game_metrics["num_refunded"], dtype=np.int32)
        out = np.empty(N, dtype=np.float64)
        self._L.sk_calc_final_rewards(_ptr(fs), _ptr(nr), N, float(mean_reward), float(reward_refunded), _ptr(out))
        return out


def evaluate_game(cards, finisher, score_penalty):
    cards = np.ascontiguousarray(cards, dtype=np.int8)
    N = cards.shape[0]
    out = np.empty(N, dtype=np.float64)
    lib().sk_evaluate_game(_ptr(cards), N, int(finisher), float(score_penalty), _ptr(out))
    return out


def calc_final_rewards(final_score, num_refunded, mean_reward, reward_refunded):
    fs = np.ascontiguousarray(final_score, dtype=np.float64)
    nr = np.ascontiguousarray(num_refunded, dtype=np.int32)
    out = np.empty(len(fs), dtype=np.float64)
    lib().sk_calc_final_rewards(_ptr(fs), _ptr(nr), len(fs), float(mean_reward), float(reward_refunded), _ptr(out))
    return out


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(4, dtype=np.uint32)
    lib().sk_philox4x32_10(_ptr(c), _ptr(k), _ptr(out))
    return out


def rng_deck(seed, env, episode):
    out = np.empty(150, dtype=np.int8)
    lib().sk_rng_deck(seed, env, episode, _ptr(out))
    return out


def rng_flips(seed, env, episode, num_players):
    out = np.empty((num_players, 2), dtype=np.uint8)
    lib().sk_rng_flips(seed, env, episode, num_players, _ptr(out))
    return out


def rng_policy(seed, env, t, mask):
    m = np.ascontiguousarray(mask, dtype=np.int8)
    return int(lib().sk_rng_policy(seed, env, t, _ptr(m)))


def rng_reshuffle(seed, env, episode, q, pile):
    """Twin in-game reshuffle on a numpy pile; returns the permuted copy."""
    ctx = _ShuffleCtx(seed, env, episode, q)
    arr = np.ascontiguousarray(pile, dtype=np.int8).copy()
    lib().sk_rng_reshuffle(C.byref(ctx), arr.ctypes.data_as(C.POINTER(C.c_int8)), len(arr))
    return arr


def rollout(num_players, score_penalty, indirect, seed, env0, games):
    """CPU baseline driver: returns (act() calls, checksum, sum of final scores)."""
    cs = C.c_uint64(0)
    ss = C.c_double(0.0)
    steps = lib().sk_rollout(num_players, float(score_penalty), int(indirect), seed, env0, games,
                             C.byref(cs), C.byref(ss))
    if steps < 0:
        raise RuntimeError(f"oracle rollout failed ({steps})")
    return int(steps), int(cs.value), float(ss.value)
