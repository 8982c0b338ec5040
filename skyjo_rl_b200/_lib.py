"""ctypes binding of libskyjo_b200.so (the C ABI declared in include/skyjo_b200.h).

There is no CPU fallback: if the shared library is missing this module raises, and every
compute entry point fails without a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SKYJO_LIB: development override used by tools/variants.py to time experimental builds of the same
# library side by side; the product always loads the in-tree libskyjo_b200.so
LIB_PATH = os.environ.get("SKYJO_LIB") or os.path.join(HERE, "libskyjo_b200.so")

MAX_PLAYERS = 12
NUM_ACTIONS = 26
DECK = 150
NUM_STATS = 32

ACT_U8, ACT_I8, ACT_I32, ACT_I64 = 0, 1, 2, 3
RUNNING, DONE_GAME_OVER, DONE_ILLEGAL, DONE_TRUNCATED = 0, 1, 2, 3

STAT_NAMES = [
    "episodes", "episode_steps", "score_raw_sum", "winner_raw_sum", "finisher_raw_sum", "penalised",
    "penalised_raw_sum", "refunds", "reshuffles", "illegal", "truncated", "act_draw_pile",
    "act_take_discard", "act_swap", "act_flip", "starter_seat0", "steps",
] + [f"wins_seat{i}" for i in range(12)]

# every symbol include/skyjo_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "skyjo_abi_version", "skyjo_last_error", "skyjo_obs_len", "skyjo_state_bytes", "skyjo_create",
    "skyjo_destroy", "skyjo_bind_outputs", "skyjo_reset", "skyjo_reset_injected", "skyjo_seed",
    "skyjo_step", "skyjo_step_random", "skyjo_rollout_random", "skyjo_profile_begin", "skyjo_profile_end",
    "skyjo_step_random_profile", "skyjo_step_host", "skyjo_set_host_threads", "skyjo_observe", "skyjo_stats_device",
    "skyjo_stats_host", "skyjo_stats_clear", "skyjo_sample_actions", "skyjo_quiesce", "skyjo_export_debug", "skyjo_check", "skyjo_step_count",
    "skyjo_set_step_count", "skyjo_launch_count", "skyjo_graph_replay_count", "skyjo_host_philox4x32_10", "skyjo_host_deck", "skyjo_host_flips",
    "skyjo_host_policy", "skyjo_host_expand_packed", "skyjo_set_host_wire", "skyjo_host_wire_bytes",
    "skyjo_host_obs_record_bytes", "skyjo_host_pack_obs", "skyjo_host_expand_obs", "skyjo_stats_allreduce",
    "skyjo_host_reshuffle", "skyjo_set_env_ranges", "skyjo_stats_allreduce_async", "skyjo_stats_allreduce_wait",
    "skyjo_host_simd_level", "skyjo_host_wire_share", "skyjo_policy_packed_bytes", "skyjo_policy_pack",
    "skyjo_policy_sample", "skyjo_policy_value", "skyjo_policy_debug", "skyjo_policy_trace", "skyjo_policy_trace_len",
]


class SkyjoConfig(C.Structure):
    _fields_ = [
        ("num_players", C.c_int32),
        ("observe_other_player_indirect", C.c_int32),
        ("score_penalty", C.c_double),
        ("mean_reward", C.c_double),
        ("reward_refunded", C.c_double),
        ("auto_reset", C.c_int32),
        ("max_episode_steps", C.c_int32),
    ]


class SkyjoOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("obs_dev", "action_mask_dev", "agent_dev", "done_dev", "reward_dev", "final_score_dev")]


class SkyjoRollout(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("obs_dev", "action_mask_dev", "agent_dev", "done_dev")]


class SkyjoEnvDebug(C.Structure):
    _fields_ = [
        ("players_cards", (C.c_int8 * 12) * MAX_PLAYERS),
        ("players_masked", (C.c_int8 * 12) * MAX_PLAYERS),
        ("discard_hist", C.c_int8 * 16),
        ("draw_hist", C.c_int8 * 16),
        ("drawpile", C.c_int8 * DECK),
        ("n_draw", C.c_int16),
        ("n_discard", C.c_int16),
        ("hand_card", C.c_int8),
        ("discard_top", C.c_int8),
        ("expected_player", C.c_int8),
        ("expected_phase", C.c_int8),
        ("starter", C.c_int8),
        ("is_terminated", C.c_int8),
        ("draw_is_multiset", C.c_int8),
        ("n_reshuffles", C.c_int8),
        ("step_in_episode", C.c_int32),
        ("episode", C.c_uint32),
        ("num_refunded", C.c_int8 * MAX_PLAYERS),
        ("num_placed", C.c_int16 * MAX_PLAYERS),
        ("pad", C.c_int8 * 8),
    ]


class SkyjoError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libskyjo_b200 error {code}: {message}")
        self.code = code


_lib = None


def load():
    """Load the shared library (building is __graft_entry__.build() / skyjo_rl_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m skyjo_rl_b200.build` "
            "(the CUDA library is the product; there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i64, u64, i32, u32 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_uint32
    cfgp = C.POINTER(SkyjoConfig)
    sig = {
        "skyjo_abi_version": (i32, []),
        "skyjo_last_error": (C.c_char_p, []),
        "skyjo_obs_len": (i32, [cfgp]),
        "skyjo_state_bytes": (i64, [cfgp, i64]),
        "skyjo_create": (i32, [cfgp, i32, i64, u64, i64, vp, i64, C.POINTER(vp)]),
        "skyjo_destroy": (i32, [vp]),
        "skyjo_bind_outputs": (i32, [vp, C.POINTER(SkyjoOutputs)]),
        "skyjo_reset": (i32, [vp, vp]),
        "skyjo_reset_injected": (i32, [vp, vp, vp, vp]),
        "skyjo_seed": (i32, [vp, u64, vp]),
        "skyjo_step": (i32, [vp, vp, i32, vp]),
        "skyjo_step_random": (i32, [vp, i32, vp]),
        "skyjo_set_env_ranges": (i32, [vp, i32]),
        "skyjo_rollout_random": (i32, [vp, i32, C.POINTER(SkyjoRollout), vp]),
        "skyjo_profile_begin": (i32, [vp]),
        "skyjo_profile_end": (i32, [vp, vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64), C.POINTER(i64)]),
        "skyjo_step_random_profile": (i32, [vp, i32, vp, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                             C.POINTER(i64), C.POINTER(i64)]),
        "skyjo_step_host": (i32, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "skyjo_set_host_threads": (i32, [vp, i32]),
        "skyjo_set_host_wire": (i32, [vp, i32]),
        "skyjo_host_wire_bytes": (i64, [vp]),
        "skyjo_host_wire_share": (i32, [vp]),
        "skyjo_observe": (i32, [vp, i32, vp, vp, vp]),
        "skyjo_stats_device": (i32, [vp, vp, vp]),
        "skyjo_stats_host": (i32, [vp, vp, vp]),
        "skyjo_stats_allreduce": (i32, [vp, vp, vp, vp]),
        "skyjo_stats_allreduce_async": (i32, [vp, vp, vp, vp]),
        "skyjo_stats_allreduce_wait": (i32, [vp, vp]),
        "skyjo_stats_clear": (i32, [vp, vp]),
        "skyjo_sample_actions": (i32, [vp, vp, vp, u64, vp, vp, vp, vp]),
        "skyjo_policy_packed_bytes": (i64, []),
        "skyjo_policy_pack": (i32, [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]),
        "skyjo_policy_sample": (i32, [vp, vp, u64, vp, vp, vp, vp, vp]),
        "skyjo_policy_value": (i32, [vp, vp, vp, vp]),
        "skyjo_policy_debug": (i32, [vp, vp, vp, vp, vp, vp]),
        "skyjo_policy_trace": (i32, [vp, vp, vp, vp, vp]),
        "skyjo_policy_trace_len": (i32, []),
        "skyjo_quiesce": (i32, [vp, vp]),
        "skyjo_export_debug": (i32, [vp, i64, i64, vp, vp]),
        "skyjo_check": (i32, [vp, vp]),
        "skyjo_step_count": (i64, [vp]),
        "skyjo_set_step_count": (i32, [vp, i64]),
        "skyjo_launch_count": (i64, [vp]),
        "skyjo_graph_replay_count": (i64, [vp]),
        "skyjo_host_philox4x32_10": (None, [vp, vp, vp]),
        "skyjo_host_deck": (None, [u64, u64, u32, vp]),
        "skyjo_host_flips": (None, [u64, u64, u32, i32, vp]),
        "skyjo_host_policy": (i32, [u64, u64, u64, u32]),
        "skyjo_host_reshuffle": (i32, [u64, u64, u32, u32, vp, i32]),
        "skyjo_host_expand_packed": (None, [vp, i64, vp, vp, vp]),
        "skyjo_host_obs_record_bytes": (i32, [i32]),
        "skyjo_host_simd_level": (i32, []),
        "skyjo_host_pack_obs": (i64, [vp, i64, i32, vp]),
        "skyjo_host_expand_obs": (None, [vp, i64, i32, vp, i32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SkyjoError(rc, load().skyjo_last_error().decode())
