"""ctypes binding of tests/hostsim/libskyjo_hostsim.so -- TEST INFRASTRUCTURE ONLY.

The library is the g++ compilation of the CUDA library's per-env __host__ __device__ functions
(skyjo_rl_b200/csrc/skyjo_core.cuh, skyjo_deal.cuh).  It lets the CPU test-suite check the
kernel logic (transition, scoring, auto-reset, observation streams, warp-shuffle staging) against
the oracle without a GPU.  Nothing under skyjo_rl_b200/ imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from skyjo_rl_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libskyjo_hostsim.so")
CSRC = os.path.join(HERE, "..", "..", "skyjo_rl_b200", "csrc")


def build(force=False):
    deps = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "..", "include", "skyjo_b200.h"))
    stale = (not os.path.exists(SO)) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps)
    if force or stale:
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-Wno-unknown-pragmas", "-o", SO, os.path.join(HERE, "hostsim.cpp")])
    return SO


_L = None


def lib():
    global _L
    if _L is None:
        build()
        L = C.CDLL(SO)
        vp = C.c_void_p
        L.hs_create.restype = vp
        L.hs_create.argtypes = [C.POINTER(_lib.SkyjoConfig), C.c_longlong, C.c_ulonglong, C.c_longlong]
        L.hs_destroy.argtypes = [vp]
        L.hs_reset.argtypes = [vp]
        L.hs_reset_injected.argtypes = [vp, vp, vp]
        L.hs_step.argtypes = [vp, vp]
        L.hs_observe.argtypes = [vp, C.c_int, vp, vp]
        L.hs_export.argtypes = [vp, C.c_longlong, vp]
        for name, typ in (("hs_obs", C.c_int8), ("hs_mask", C.c_int8), ("hs_agent", C.c_int8),
                          ("hs_done", C.c_uint8), ("hs_reward", C.c_double), ("hs_score", C.c_double),
                          ("hs_stats", C.c_longlong)):
            getattr(L, name).restype = C.POINTER(typ)
            getattr(L, name).argtypes = [vp]
        L.hs_errflag.restype = C.c_uint32
        L.hs_errflag.argtypes = [vp]
        _L = L
    return _L


class HostSimEnv:
    """numpy twin of BatchedSkyjoEnv's buffers, driven by the host-compiled device functions."""

    def __init__(self, num_envs, num_players=2, score_penalty=2.0, observe_other_player_indirect=False,
                 mean_reward=1.0, reward_refunded=0.0, seed=0, auto_reset=True, max_episode_steps=0,
                 first_global_env_id=0):
        self.L = lib()
        self.num_envs, self.num_players = num_envs, num_players
        self.cfg = _lib.SkyjoConfig(num_players, int(observe_other_player_indirect), score_penalty, mean_reward,
                                    reward_refunded, int(auto_reset), max_episode_steps)
        self.h = self.L.hs_create(C.byref(self.cfg), num_envs, seed, first_global_env_id)
        self.obs_len = 31 if observe_other_player_indirect else 19 + 12 * num_players
        B, N, D = num_envs, num_players, self.obs_len
        as_np = np.ctypeslib.as_array
        self.observations = as_np(self.L.hs_obs(self.h), shape=(B, D))
        self.action_mask = as_np(self.L.hs_mask(self.h), shape=(B, 26))
        self.agent_selection = as_np(self.L.hs_agent(self.h), shape=(B,))
        self.done_code = as_np(self.L.hs_done(self.h), shape=(B,))
        self.rewards = as_np(self.L.hs_reward(self.h), shape=(B, N))
        self.final_scores = as_np(self.L.hs_score(self.h), shape=(B, N))
        self._stats = as_np(self.L.hs_stats(self.h), shape=(_lib.NUM_STATS,))

    def __del__(self):
        try:
            self.L.hs_destroy(self.h)
        except Exception:  # noqa: BLE001
            pass

    def reset(self):
        self.L.hs_reset(self.h)

    def reset_injected(self, decks, flips):
        decks = np.ascontiguousarray(decks, dtype=np.int8)
        flips = np.ascontiguousarray(flips, dtype=np.uint8)
        self.L.hs_reset_injected(self.h, decks.ctypes.data, flips.ctypes.data)
        assert self.L.hs_errflag(self.h) == 0

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.int32)
        self.L.hs_step(self.h, a.ctypes.data)

    def step_random(self, n=1):
        for _ in range(n):
            self.L.hs_step(self.h, None)

    def set_env_ranges(self, n=0):
        pass    # a launch-scheduling knob of the CUDA library; the host build steps env by env

    def observe(self, agent):
        obs = np.empty_like(self.observations)
        mask = np.empty_like(self.action_mask)
        self.L.hs_observe(self.h, int(agent), obs.ctypes.data, mask.ctypes.data)
        return {"observations": obs, "action_mask": mask}

    def stats(self):
        return dict(zip(_lib.STAT_NAMES, [int(x) for x in self._stats]))

    def check(self):
        assert self.L.hs_errflag(self.h) == 0, hex(self.L.hs_errflag(self.h))

    def export(self, env0=0, count=None):
        from skyjo_rl_b200.env import GameView
        count = self.num_envs - env0 if count is None else count
        out = []
        for i in range(count):
            d = _lib.SkyjoEnvDebug()
            self.L.hs_export(self.h, env0 + i, C.byref(d))
            out.append(GameView(d, self.num_players))
        return out
