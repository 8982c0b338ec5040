"""BatchedSkyjoEnv -- the host-side mirror of the reference's PettingZoo env for a batch of
lockstep games on one B200.

It keeps the surface of `SimpleSkyjoEnv` (reference rlskyjo/environment/skyjo_env.py:29-332):
same constructor keywords (:38-45), `reset / step / observe / last / seed / render / close`,
`agent_selection`, `agents`, `possible_agents`, `rewards`, `_cumulative_rewards`, `dones`
(+ `terminations` / `truncations`), `infos`, `observation_space(agent)`, `action_space(agent)`.
Everything that was a scalar or a small numpy array per game is a torch tensor over the env
batch, living in device buffers that the CUDA library writes in place (zero copy):

    obs["observations"]  int8   [B, D]     obs["action_mask"]  int8 [B, 26]
    agent_selection      int8   [B]        (agent name = f"player_{i}", skyjo_env.py:116,324)
    rewards              float64[B, N]     dones               bool [B]

All game logic runs in libskyjo_b200.so (hand-written sm_100a kernels behind the C ABI of
include/skyjo_b200.h).  PyTorch only owns the memory and the stream.  There is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .spaces import Box, Dict, Discrete

DEFAULT_CONFIG = {  # reference skyjo_env.py:10-16
    "num_players": 3,
    "score_penalty": 2.0,
    "observe_other_player_indirect": True,
    "mean_reward": 1.0,
    "reward_refunded": 0.001,
}

_ACT_DTYPES = {torch.uint8: _lib.ACT_U8, torch.int8: _lib.ACT_I8, torch.int32: _lib.ACT_I32,
               torch.int64: _lib.ACT_I64}


class BatchedSkyjoEnv:
    metadata = {"render.modes": ["human"], "name": "skyjo", "is_parallelizable": False,
                "video.frames_per_second": 1}  # skyjo_env.py:31-36

    def __init__(self, num_envs, num_players=2, score_penalty=2.0, observe_other_player_indirect=False,
                 mean_reward=1.0, reward_refunded=0.0, device="cuda:0", seed=0, auto_reset=True,
                 max_episode_steps=0, first_global_env_id=0):
        assert 0 < num_players <= 12, \
            "Skyjo can be played from 1 up to 8 (recommended) / 12 (theoretical) players"  # skyjo.py:24-26
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedSkyjoEnv needs a CUDA device (B200); there is no CPU fallback")
        self._L = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedSkyjoEnv runs on CUDA devices only")
        if self.device.index is None:   # "cuda" = the current device, with an explicit index from here on
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.num_players = int(num_players)
        self.mean_reward = float(mean_reward)
        self.reward_refunded = float(reward_refunded)
        self.score_penalty = float(score_penalty)
        self.observe_other_player_indirect = bool(observe_other_player_indirect)
        # auto_reset: False (freeze), True / "same_step" (the step that ends an episode installs the next
        # one), "next_step" (phase-locked: terminal observation first, the reset takes the env's next
        # lockstep slot -- include/skyjo_b200.h SKYJO_RESET_*)
        modes = {False: 0, True: 1, "off": 0, "same_step": 1, "next_step": 2, 0: 0, 1: 1, 2: 2}
        if auto_reset not in modes:
            raise ValueError(f"auto_reset must be False, True, 'same_step' or 'next_step', not {auto_reset!r}")
        self.reset_mode = modes[auto_reset]
        self.auto_reset = self.reset_mode != 0
        self._cfg = _lib.SkyjoConfig(self.num_players, int(self.observe_other_player_indirect),
                                     self.score_penalty, self.mean_reward, self.reward_refunded,
                                     self.reset_mode, int(max_episode_steps))
        self.obs_len = self._L.skyjo_obs_len(C.byref(self._cfg))
        self.obs_shape = (self.obs_len,)          # skyjo.py:43-45
        self.action_mask_shape = (_lib.NUM_ACTIONS,)  # skyjo.py:46
        B, N, D = self.num_envs, self.num_players, self.obs_len

        nbytes = self._L.skyjo_state_bytes(C.byref(self._cfg), B)
        with torch.cuda.device(self.device):
            self._state = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.observations = torch.empty((B, D), dtype=torch.int8, device=self.device)
            self.action_mask = torch.empty((B, 26), dtype=torch.int8, device=self.device)
            self.agent_selection = torch.zeros(B, dtype=torch.int8, device=self.device)
            self.done_code = torch.zeros(B, dtype=torch.uint8, device=self.device)
            self.rewards = torch.zeros((B, N), dtype=torch.float64, device=self.device)
            self.final_scores = torch.zeros((B, N), dtype=torch.float64, device=self.device)
        handle = C.c_void_p()
        _lib.check(self._L.skyjo_create(C.byref(self._cfg), self.device.index, B, int(seed),
                                        int(first_global_env_id), self._state.data_ptr(), nbytes,
                                        C.byref(handle)))
        self._h = handle
        outs = _lib.SkyjoOutputs(self.observations.data_ptr(), self.action_mask.data_ptr(),
                                 self.agent_selection.data_ptr(), self.done_code.data_ptr(),
                                 self.rewards.data_ptr(), self.final_scores.data_ptr())
        _lib.check(self._L.skyjo_bind_outputs(self._h, C.byref(outs)))
        self._seed = int(seed)
        self.first_global_env_id = int(first_global_env_id)

        # PettingZoo API surface (skyjo_env.py:116-151)
        self.agents = [f"player_{i}" for i in range(N)]
        self.possible_agents = self.agents[:]
        self.infos = {a: {} for a in self.agents}
        obs_space = Dict({
            "observations": Box(low=-24, high=127, shape=self.obs_shape, dtype=np.int8),
            "action_mask": Box(low=0, high=1, shape=self.action_mask_shape, dtype=np.int8),
        })
        self._observation_spaces = {a: obs_space for a in self.possible_agents}
        self._action_spaces = {a: Discrete(_lib.NUM_ACTIONS) for a in self.possible_agents}
        self._has_reset = False
        self.stream = None

    # ---- plumbing ---------------------------------------------------------------------
    def _stream(self):
        # `stream` (a torch.cuda.Stream, default None = torch's current stream at the time of each call) pins every
        # launch of this env -- and of a FusedPolicy built on it -- to one stream, so that two envs can be driven
        # from one thread on two streams without a stream context per call.  Tensors handed in from other streams
        # need the caller's own wait_stream / record_stream, and buffers should be allocated up front (torch's
        # caching allocator associates a block with the stream that was current when it was allocated).  Methods
        # that convert or allocate tensors themselves (reset_injected, step with host-side or mistyped actions,
        # rollout_random / sample_actions without output buffers, stats, export) do that on torch's CURRENT
        # stream: with a pinned stream call them under `with torch.cuda.stream(env.stream):`.
        st = self.stream
        return st.cuda_stream if st is not None else torch.cuda.current_stream(self.device).cuda_stream

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def close(self):
        """part of the PettingZoo API (skyjo_env.py:276-278)"""
        h, self._h = getattr(self, "_h", None), None
        if h:
            torch.cuda.synchronize(self.device)
            self._L.skyjo_destroy(h)

    @property
    def num_agents(self):
        return len(self.agents)

    def observation_space(self, agent):
        return self._observation_spaces[agent]

    def action_space(self, agent):
        return self._action_spaces[agent]

    @property
    def observation_spaces(self):
        return self._observation_spaces

    @property
    def action_spaces(self):
        return self._action_spaces

    @staticmethod
    def _name_to_player_id(name):  # skyjo_env.py:314-324
        return int(name.split("_")[-1])

    # ---- episodes -----------------------------------------------------------------------
    def reset(self, seed=None):
        """SimpleSkyjoEnv.reset (skyjo_env.py:254-267) for every env."""
        if seed is not None:
            return self.seed(seed)
        _lib.check(self._L.skyjo_reset(self._h, self._stream()))
        self._has_reset = True

    def seed(self, seed=None):
        """SimpleSkyjoEnv.seed (skyjo_env.py:280-290): reseed and redeal."""
        if seed is not None:
            self._seed = int(seed)
            _lib.check(self._L.skyjo_seed(self._h, self._seed, self._stream()))
            self._has_reset = True

    def reset_injected(self, decks, flips):
        """Deal given deck orders (int8 [B,150]) and open slots (uint8 [B,N,2]) instead of the
        Philox shuffle -- the injection convention of SURVEY.md 9.1, for replaying games."""
        decks = torch.as_tensor(decks, dtype=torch.int8).to(self.device).contiguous()
        flips = torch.as_tensor(flips, dtype=torch.uint8).to(self.device).contiguous()
        assert decks.shape == (self.num_envs, 150) and flips.shape == (self.num_envs, self.num_players, 2)
        _lib.check(self._L.skyjo_reset_injected(self._h, decks.data_ptr(), flips.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # decks / flips must outlive the launch
        _lib.check(self._L.skyjo_check(self._h, self._stream()))
        self._has_reset = True

    def step(self, actions):
        """SimpleSkyjoEnv.step (skyjo_env.py:216-252) for every env: `actions[i]` in 0..25 is
        played by env i's agent_selection.  Afterwards observations / action_mask /
        agent_selection describe the next turn, done_code / rewards the step's outcome."""
        assert self._has_reset, "reset() needs to be called before step"  # OrderEnforcingWrapper
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions))
        if actions.device != self.device:
            actions = actions.to(self.device)
        if actions.dtype not in _ACT_DTYPES:
            actions = actions.to(torch.int64)
        actions = actions.contiguous()
        assert actions.shape == (self.num_envs,)
        _lib.check(self._L.skyjo_step(self._h, actions.data_ptr(), _ACT_DTYPES[actions.dtype], self._stream()))

    def step_random(self, n_steps=1):
        """n_steps lockstep steps with the uniform legal policy drawn in-kernel
        (the loop of rlskyjo/game/sample_game.py:10-21)."""
        assert self._has_reset, "reset() needs to be called before step"
        _lib.check(self._L.skyjo_step_random(self._h, int(n_steps), self._stream()))

    def set_env_ranges(self, n=0):
        """step_random steps the batch as n independent env ranges on n CUDA streams (0 = default: 4 from 2^18
        envs, else 1).  Same games for every n."""
        _lib.check(self._L.skyjo_set_env_ranges(self._h, int(n)))

    def rollout_random(self, n_steps, out=None):
        """The env-steps of step_random(n_steps) as multi-step launches (state in registers across
        up to 32 steps per kernel); what each step publishes goes to slice t of time-major tensors
        {"observations": int8[T,B,D], "action_mask": int8[T,B,26], "agent_selection": int8[T,B],
        "done_code": uint8[T,B]} (allocated here unless `out` is given)."""
        assert self._has_reset, "reset() needs to be called before step"
        T, B = int(n_steps), self.num_envs
        if out is None:
            with torch.cuda.device(self.device):
                out = {"observations": torch.empty((T, B, self.obs_len), dtype=torch.int8, device=self.device),
                       "action_mask": torch.empty((T, B, 26), dtype=torch.int8, device=self.device),
                       "agent_selection": torch.empty((T, B), dtype=torch.int8, device=self.device),
                       "done_code": torch.empty((T, B), dtype=torch.uint8, device=self.device)}
        for k, shape in (("observations", (B, self.obs_len)), ("action_mask", (B, 26)), ("agent_selection", (B,)),
                         ("done_code", (B,))):
            t = out[k]
            assert t.is_contiguous() and t.device == self.device and t.shape[0] >= T and tuple(t.shape[1:]) == shape
        ro = _lib.SkyjoRollout(out["observations"].data_ptr(), out["action_mask"].data_ptr(),
                               out["agent_selection"].data_ptr(), out["done_code"].data_ptr())
        _lib.check(self._L.skyjo_rollout_random(self._h, T, C.byref(ro), self._stream()))
        return out

    def profile_begin(self):
        _lib.check(self._L.skyjo_profile_begin(self._h))

    def profile_end(self):
        sm, dm = C.c_double(0), C.c_double(0)
        ns, nd = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.skyjo_profile_end(self._h, self._stream(), C.byref(sm), C.byref(dm), C.byref(ns),
                                             C.byref(nd)))
        return {"step_ms": sm.value, "deal_ms": dm.value, "step_launches": ns.value, "deal_launches": nd.value}

    def step_random_profile(self, n_steps):
        """step_random on one stream with CUDA-event timing (one pair per window of back-to-back step launches,
        one per refill deal); returns the summed device times and the launch counts."""
        sm, dm = C.c_double(0), C.c_double(0)
        ns, nd = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.skyjo_step_random_profile(self._h, int(n_steps), self._stream(), C.byref(sm),
                                                     C.byref(dm), C.byref(ns), C.byref(nd)))
        return {"step_ms": sm.value, "deal_ms": dm.value, "step_launches": ns.value, "deal_launches": nd.value}

    def step_host(self, actions, obs=None, mask=None, agent=None, done=None, reward=None):
        """End-to-end host entry: numpy/pinned uint8 actions in, numpy outputs back
        (host-to-device and device-to-host copies included; synchronises)."""
        assert self._has_reset, "reset() needs to be called before step"

        def ptr(t):
            if t is None:
                return None
            return t.data_ptr() if torch.is_tensor(t) else t.ctypes.data
        _lib.check(self._L.skyjo_step_host(self._h, ptr(actions), ptr(obs), ptr(mask), ptr(agent), ptr(done),
                                           ptr(reward), self._stream()))

    def sample_actions(self, logits, mask=None, seed=0, actions=None, logp=None, entropy=None):
        """Fused masked softmax + categorical sample (csrc/skyjo_sample.cuh): float32 logits [B,26]
        -> uint8 actions [B] for step(), float32 logp [B] (and entropy [B] if a tensor is given).
        `mask` defaults to the live action_mask buffer (zero copy)."""
        B = self.num_envs
        assert logits.dtype == torch.float32 and logits.is_contiguous() and tuple(logits.shape) == (B, 26)
        mask = self.action_mask if mask is None else mask
        assert mask.dtype == torch.int8 and mask.is_contiguous() and tuple(mask.shape) == (B, 26)
        if actions is None:
            actions = torch.empty(B, dtype=torch.uint8, device=self.device)
        if logp is None:
            logp = torch.empty(B, dtype=torch.float32, device=self.device)
        _lib.check(self._L.skyjo_sample_actions(self._h, logits.data_ptr(), mask.data_ptr(), int(seed),
                                                actions.data_ptr(), logp.data_ptr(),
                                                entropy.data_ptr() if entropy is not None else None, self._stream()))
        return actions, logp

    def bind_outputs(self, observations=None, action_mask=None, agent_selection=None, done_code=None, rewards=None,
                     final_scores=None):
        """Redirect the kernels' outputs (e.g. to slice t of a rollout storage: the env then writes
        the learner's tensors in place).  Omitted tensors keep their current buffer."""
        B, N, D = self.num_envs, self.num_players, self.obs_len
        spec = {"observations": (observations, torch.int8, (B, D)), "action_mask": (action_mask, torch.int8, (B, 26)),
                "agent_selection": (agent_selection, torch.int8, (B,)), "done_code": (done_code, torch.uint8, (B,)),
                "rewards": (rewards, torch.float64, (B, N)), "final_scores": (final_scores, torch.float64, (B, N))}
        for name, (t, dt, shape) in spec.items():
            if t is not None:
                assert t.dtype == dt and tuple(t.shape) == shape and t.is_contiguous() and t.device == self.device, name
                setattr(self, name, t)
        outs = _lib.SkyjoOutputs(self.observations.data_ptr(), self.action_mask.data_ptr(),
                                 self.agent_selection.data_ptr(), self.done_code.data_ptr(),
                                 self.rewards.data_ptr(), self.final_scores.data_ptr())
        _lib.check(self._L.skyjo_bind_outputs(self._h, C.byref(outs)))

    def set_host_threads(self, n=0):
        """Worker threads step_host uses to expand the packed mask / agent / done words (0 = default)."""
        _lib.check(self._L.skyjo_set_host_threads(self._h, int(n)))

    def set_host_wire(self, mode="raw"):
        """How step_host moves observation rows over the link: "raw" rows, "compact" records (packed on the device,
        expanded on the host) or "mixed" (an adaptive share of the env ranges compact; the default where the CPU
        has AVX-512) -- csrc/skyjo_hostio.cuh."""
        _lib.check(self._L.skyjo_set_host_wire(self._h, {"raw": 0, "compact": 1, "mixed": 2}[mode]))

    @property
    def host_wire_bytes(self):
        """bytes the last step_host call moved device -> host"""
        return int(self._L.skyjo_host_wire_bytes(self._h))

    @property
    def host_wire_share(self):
        """of 8 env ranges, how many step_host currently sends as compact records"""
        return int(self._L.skyjo_host_wire_share(self._h))

    def observe(self, agent=None):
        """SimpleSkyjoEnv.observe (skyjo_env.py:199-214).  agent=None returns the live buffers
        (view of each env's agent_selection); a name or index encodes that seat's view."""
        if agent is None:
            return {"observations": self.observations, "action_mask": self.action_mask}
        idx = self._name_to_player_id(agent) if isinstance(agent, str) else int(agent)
        obs = torch.empty_like(self.observations)
        mask = torch.empty_like(self.action_mask)
        _lib.check(self._L.skyjo_observe(self._h, idx, obs.data_ptr(), mask.data_ptr(), self._stream()))
        return {"observations": obs, "action_mask": mask}

    def last(self):
        """AECEnv.last(): (observation, cumulative reward, done, info) of agent_selection, batched.
        Rewards are zero until an episode ends, so the cumulative reward equals `rewards`.
        AEC-faithful with auto_reset=False or "next_step": the step that ends a game leaves the finisher selected,
        with the terminal observation, as the reference's env does (skyjo_env.py:239-247).  In the default same-step
        mode that step has already installed the next episode, so agent_selection / the observation belong to the
        NEW game while `rewards` / `dones` still describe the finished one: read `rewards[:, seat]` per seat there
        (what ppo.py does) instead of pairing last()'s reward with its observation."""
        sel = self.agent_selection.long().unsqueeze(1)
        reward = self.rewards.gather(1, sel).squeeze(1)
        return self.observe(), reward, self.dones, {}

    # ---- outcome views ------------------------------------------------------------------
    @property
    def dones(self):
        return self.done_code != 0

    @property
    def terminations(self):
        return (self.done_code == _lib.DONE_GAME_OVER) | (self.done_code == _lib.DONE_ILLEGAL)

    @property
    def truncations(self):
        return self.done_code == _lib.DONE_TRUNCATED

    @property
    def _cumulative_rewards(self):
        return self.rewards

    def agent_name(self, index):
        return f"player_{int(index)}"

    # ---- statistics / debugging -------------------------------------------------------
    def stats_tensor(self, comm=None):
        """int64[32] statistics vector on the device (input of the NCCL all-reduce); with `comm` (a
        skyjo_rl_b200.nccl.StatsComm) already summed over its ranks by the library's own ncclAllReduce
        (skyjo_stats_allreduce), asynchronously on the env's stream."""
        out = torch.empty(_lib.NUM_STATS, dtype=torch.int64, device=self.device)
        if comm is not None:
            _lib.check(self._L.skyjo_stats_allreduce(self._h, comm.handle, out.data_ptr(), self._stream()))
        else:
            _lib.check(self._L.skyjo_stats_device(self._h, out.data_ptr(), self._stream()))
        return out

    def stats_allreduce_async(self, comm=None):
        """Per-iteration statistics all-reduce off the step path (skyjo_stats_allreduce_async): the env's stream only
        runs the one-CTA local reduction, the ncclAllReduce over `comm` (a StatsComm; None = this rank alone) runs on
        a side stream the library owns.  Returns the int64[32] device tensor the sum lands in -- two tensors
        alternate -- valid after stats_allreduce_wait()."""
        if not hasattr(self, "_ar_out"):
            self._ar_out = [torch.zeros(_lib.NUM_STATS, dtype=torch.int64, device=self.device) for _ in range(2)]
            self._ar_k = 0
        out = self._ar_out[self._ar_k]
        self._ar_k ^= 1
        _lib.check(self._L.skyjo_stats_allreduce_async(self._h, comm.handle if comm is not None else None,
                                                       out.data_ptr(), self._stream()))
        return out

    def stats_allreduce_wait(self):
        """The env's stream waits for every statistics all-reduce still in flight on the side stream."""
        _lib.check(self._L.skyjo_stats_allreduce_wait(self._h, self._stream()))

    def stats(self, all_reduce=False, group=None, comm=None):
        """Episode statistics as a dict.  With all_reduce=True the vector is summed over the
        ranks of `group` with torch.distributed (NCCL) first -- the only collective of the env,
        never on the step path.  With `comm` (a skyjo_rl_b200.nccl.StatsComm) the sum is the
        library's own ncclAllReduce (skyjo_stats_allreduce) on the env's stream."""
        if comm is not None:
            return dict(zip(_lib.STAT_NAMES, self.stats_tensor(comm).tolist()))
        vec = self.stats_tensor()
        if all_reduce:
            torch.distributed.all_reduce(vec, group=group)
        return dict(zip(_lib.STAT_NAMES, vec.tolist()))

    def clear_stats(self):
        _lib.check(self._L.skyjo_stats_clear(self._h, self._stream()))

    def check(self):
        """Synchronise and raise if a kernel flagged an inconsistency."""
        _lib.check(self._L.skyjo_check(self._h, self._stream()))

    @property
    def step_count(self):
        return int(self._L.skyjo_step_count(self._h))

    @property
    def launch_count(self):
        return int(self._L.skyjo_launch_count(self._h))

    @property
    def graph_replay_count(self):
        """step_random calls served by a CUDA-graph replay so far"""
        return int(self._L.skyjo_graph_replay_count(self._h))

    def export(self, env0=0, count=None):
        """SkyjoGame-shaped dump of envs [env0, env0+count) as a list of GameView."""
        count = self.num_envs - env0 if count is None else count
        nbytes = C.sizeof(_lib.SkyjoEnvDebug) * count
        buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        _lib.check(self._L.skyjo_export_debug(self._h, env0, count, buf.data_ptr(), self._stream()))
        host = buf.cpu().numpy().tobytes()
        arr = (_lib.SkyjoEnvDebug * count).from_buffer_copy(host)
        return [GameView(arr[i], self.num_players) for i in range(count)]

    def game_view(self, env_index):
        v = self.export(env_index, 1)[0]
        if v.is_terminated:  # frozen env (auto_reset=False): the scores of the finished game
            v.game_metrics["final_score"] = [float(x) for x in self.final_scores[env_index].tolist()]
        return v

    def render(self, env_index=0, mode="human"):
        """SimpleSkyjoEnv.render (skyjo_env.py:269-274) for one env of the batch."""
        if mode == "human":
            print(self.game_view(env_index).render_table())

    # ---- checkpoint ---------------------------------------------------------------------
    def state_dict(self):
        """Everything needed to resume: the flat device state plus the published outputs."""
        _lib.check(self._L.skyjo_quiesce(self._h, self._stream()))
        return {
            "config": {k: getattr(self._cfg, k) for k, _ in _lib.SkyjoConfig._fields_},
            "num_envs": self.num_envs, "seed": self._seed, "first_global_env_id": self.first_global_env_id,
            "step_count": self.step_count,
            "state": self._state.clone(), "observations": self.observations.clone(),
            "action_mask": self.action_mask.clone(), "agent_selection": self.agent_selection.clone(),
            "done_code": self.done_code.clone(), "rewards": self.rewards.clone(),
            "final_scores": self.final_scores.clone(),
        }

    def load_state_dict(self, sd):
        """Resume from state_dict().  Every later Philox draw (deal, reshuffle, in-kernel policy, sample) is keyed
        by (seed, global env id) and the rewards by the config, all of which live in the handle: a checkpoint only
        continues the same games in an env built with the same values, so a mismatch raises."""
        mine = {k: getattr(self._cfg, k) for k, _ in _lib.SkyjoConfig._fields_}
        diff = {k: (sd["config"].get(k), v) for k, v in mine.items() if sd["config"].get(k) != v}
        for k, v in (("num_envs", self.num_envs), ("seed", self._seed), ("first_global_env_id", self.first_global_env_id)):
            if sd[k] != v:
                diff[k] = (sd[k], v)
        if diff:
            raise ValueError(f"checkpoint does not belong to this env (saved, this env): {diff}")
        _lib.check(self._L.skyjo_quiesce(self._h, self._stream()))
        for name in ("observations", "action_mask", "agent_selection", "done_code", "rewards", "final_scores"):
            getattr(self, name).copy_(sd[name])
        self._state.copy_(sd["state"])
        _lib.check(self._L.skyjo_set_step_count(self._h, int(sd["step_count"])))
        self._has_reset = True


class GameView:
    """Read-only `SkyjoGame`-shaped snapshot of one env (reference skyjo.py attribute names)."""

    def __init__(self, d, num_players):
        N = num_players
        self.num_players = N
        self.players_cards = np.array([list(d.players_cards[p]) for p in range(N)], dtype=np.int8)
        self.players_masked = np.array([list(d.players_masked[p]) for p in range(N)], dtype=np.int8)
        self.hand_card = int(d.hand_card)
        self.discard_top = int(d.discard_top)
        self.expected_action = [int(d.expected_player), "draw" if d.expected_phase == 0 else "place"]
        self.is_terminated = bool(d.is_terminated)
        self.starter = int(d.starter)
        self.step_in_episode = int(d.step_in_episode)
        self.episode = int(d.episode)
        self.n_reshuffles = int(d.n_reshuffles)
        self.draw_is_multiset = bool(d.draw_is_multiset)
        self.n_draw = int(d.n_draw)
        self.n_discard = int(d.n_discard)
        self.discard_hist = np.array(list(d.discard_hist[:15]), dtype=np.int64) & 0xFF
        self.draw_hist = np.array(list(d.draw_hist[:15]), dtype=np.int64) & 0xFF
        self.drawpile = None if self.draw_is_multiset else [int(x) for x in d.drawpile[: self.n_draw]]
        self.game_metrics = {
            "num_refunded": [int(x) for x in d.num_refunded[:N]],
            "num_placed": [int(x) for x in d.num_placed[:N]],
        }

    @property
    def discard_pile_sorted(self):
        """The discard pile as a sorted multiset (its order below the top is not kept on the GPU)."""
        return [v - 2 for v in range(15) for _ in range(int(self.discard_hist[v]))]

    def render_table(self):
        """Board in the format of SkyjoGame.render_table (skyjo.py:507-564)."""
        hand = self.hand_card if -2 <= self.hand_card <= 12 else "empty"
        top = self.discard_top if self.discard_top != -3 else "empty"
        s = f"{'='*7} render board: {'='*5} \n{'='*7} stats {'='*12} \n"
        s += f"next turn: {self.expected_action[1]} by Player {self.expected_action[0]} \n"
        s += f"holding card player {self.expected_action[0]}: {hand} \ndiscard pile top: {top} \n"
        if self.is_terminated:  # skyjo.py:516-521
            res = dict(zip(range(self.num_players), self.game_metrics.get("final_score", [])))
            s += f"{'='*7} GAME DONE {'='*8} \nResults: {res} \n"
        for p in range(self.num_players):
            arr = self.players_cards[p].astype(np.str_)
            hid = self.players_masked[p] == 2
            arr[hid] = np.char.add("u", arr[hid]) if self.is_terminated else "u"
            arr[self.players_masked[p] == 0] = "d"
            arr = arr.reshape(4, -1).T
            s += f"{'='*7} Player {p} {'='*10} \n"
            s += np.array2string(arr, separator="\t ", formatter={"str_kind": lambda x: str(x)}) + "\n"
        return s


def render_action_explainer(action_int):
    """Text for an action id, format of SkyjoGame.render_action_explainer (skyjo.py:566-590; the
    row is printed as place_id % 4 there, kept)."""
    assert action_int in range(0, 26), "action not valid action int {action_int}"
    if action_int == 24:
        return "draw from drawpile"
    if action_int == 25:
        return "draw from discard pile"
    if action_int < 12:
        place_id, head = action_int, f"place card ({action_int}) - "
    else:
        place_id, head = action_int - 12, f"handcard discard & reveal card ({action_int}) - "
    return head + f"col:{place_id // 3} row:{place_id % 4}"


def render_actions():
    """Legend of the 26 action ids, format of SkyjoGame.render_actions (skyjo.py:592-602)."""
    grid = [[f"{3 * c + r}/{12 + 3 * c + r}" for c in range(4)] for r in range(3)]
    body = "\n ".join("[" + "\t ".join(row) + "]" for row in grid)
    return ("action ids 0-25: \n(put handcard here / reveal this card) \n [" + body + "] \n"
            "24: draw from drawpile \n 25: draw from discard pile")
