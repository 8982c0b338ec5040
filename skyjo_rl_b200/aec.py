"""Single-game PettingZoo-style AEC view over one env of a batch.

`SkyjoAECView` exposes the agent-environment-cycle surface the reference's consumers use
(reference rlskyjo/environment/vanilla_env_example.py:14-35 and RLlib's PettingZooEnv):
`reset / agent_iter / last / step(action | None) / observe / agent_selection / agents / rewards /
_cumulative_rewards / dones / infos`, on top of any backend with the buffers of
`BatchedSkyjoEnv` (created with `auto_reset=False`).

The bookkeeping restates PettingZoo 1.14.0's `AECEnv` helpers as used by
`SimpleSkyjoEnv.step` (reference rlskyjo/environment/skyjo_env.py:216-252): `_accumulate_rewards`,
`_clear_rewards`, `_dones_step_first`, `_was_done_step`, and the wrapper stack of
`skyjo_env.env()` (:19-26): `TerminateIllegalWrapper(illegal_reward=-1)` (done in-kernel: done
code 2, offender -1, others 0), `AssertOutOfBoundsWrapper`, `OrderEnforcingWrapper`.
PettingZoo itself is not installed here: this is a restatement from its documented behaviour
and the notebook trace, functional rather than bit-pinned parity.
"""
import numpy as np

from . import _lib


def _np(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


class SkyjoAECView:
    metadata = {"render.modes": ["human"], "name": "skyjo", "is_parallelizable": False,
                "video.frames_per_second": 1}  # skyjo_env.py:31-36

    def __init__(self, batched_env, index=0):
        self.env = batched_env
        self.index = int(index)
        N = batched_env.num_players
        self.possible_agents = [f"player_{i}" for i in range(N)]  # skyjo_env.py:116
        self.agents = []
        self.agent_selection = None
        self._has_reset = False
        self._skip_agent_selection = None
        # One-env backends with a host entry (BatchedSkyjoEnv.step_host -> skyjo_step_host) step through it: one
        # library call returns the next agent's observation, mask, agent, done code and rewards in pinned host
        # buffers, instead of a launch plus five small device-to-host copies; last() then reads the cached copy.
        self._host = None
        self._cached = None                  # (agent name, observations, action_mask) as published by the last step
        if batched_env.num_envs == 1 and hasattr(batched_env, "step_host"):
            import torch
            pin = lambda shape, dt: torch.zeros(shape, dtype=dt).pin_memory()  # noqa: E731
            self._host = {"act": pin((1,), torch.uint8), "obs": pin((1, batched_env.obs_len), torch.int8),
                          "mask": pin((1, _lib.NUM_ACTIONS), torch.int8), "agent": pin((1,), torch.int8),
                          "done": pin((1,), torch.uint8), "reward": pin((1, N), torch.float64)}
            self._host_np = {k: v.numpy() for k, v in self._host.items()}

    # ---- spaces (delegated) ---------------------------------------------------------------
    def observation_space(self, agent):
        return self.env.observation_space(agent)

    def action_space(self, agent):
        return self.env.action_space(agent)

    @property
    def num_agents(self):
        return len(self.agents)

    # ---- AEC API -----------------------------------------------------------------------------
    def reset(self):
        """skyjo_env.py:254-267"""
        self.env.reset()
        self._cached = None
        self.agents = self.possible_agents[:]
        self.rewards = {a: 0 for a in self.agents}
        self._cumulative_rewards = {a: 0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        self.infos = {a: {} for a in self.agents}
        self.agent_selection = self._expected_agent()
        self._skip_agent_selection = None
        self._has_reset = True

    def seed(self, seed=None):
        if seed is not None:
            self.env.seed(seed)
            self._begin_episode()

    def reset_injected(self, deck, flips):
        """reset() with the deck order (int8[150]) and the open slots (uint8[N,2]) supplied instead
        of shuffled: replays a recorded game.  Single-env backends only."""
        assert self.env.num_envs == 1
        self.env.reset_injected(np.asarray(deck, np.int8)[None], np.asarray(flips, np.uint8)[None])
        self._begin_episode()

    def _begin_episode(self):
        self._cached = None
        self.agents = self.possible_agents[:]
        self.rewards = {a: 0 for a in self.agents}
        self._cumulative_rewards = {a: 0 for a in self.agents}
        self.dones = {a: False for a in self.agents}
        self.infos = {a: {} for a in self.agents}
        self.agent_selection = self._expected_agent()
        self._skip_agent_selection = None
        self._has_reset = True

    def _expected_agent(self):
        return f"player_{int(_np(self.env.agent_selection)[self.index])}"

    def observe(self, agent):
        """skyjo_env.py:199-214: observation dict of any named agent"""
        assert self._has_reset, "reset() needs to be called before observe"  # OrderEnforcingWrapper
        if self._cached is not None and self._cached[0] == agent:
            return {"observations": self._cached[1].copy(), "action_mask": self._cached[2].copy()}
        o = self.env.observe(int(agent.split("_")[-1]))
        return {"observations": _np(o["observations"])[self.index].copy(),
                "action_mask": _np(o["action_mask"])[self.index].copy()}

    def last(self, observe=True):
        agent = self.agent_selection
        obs = self.observe(agent) if observe else None
        return obs, self._cumulative_rewards[agent], self.dones[agent], self.infos[agent]

    def agent_iter(self, max_iter=2 ** 63):
        it = max_iter
        while self.agents and it > 0:
            it -= 1
            yield self.agent_selection

    def step(self, action):
        """skyjo_env.py:216-252 under the wrapper stack of skyjo_env.env() (:19-26)"""
        assert self._has_reset, "reset() needs to be called before step"
        agent = self.agent_selection
        if self.dones[agent]:
            return self._was_done_step(action)
        assert action is not None and 0 <= int(action) < _lib.NUM_ACTIONS, \
            "action is not in action space"                     # AssertOutOfBoundsWrapper
        if self._host is not None:
            h, hn = self._host, self._host_np
            hn["act"][0] = int(action)
            self.env.step_host(h["act"], h["obs"], h["mask"], h["agent"], h["done"], h["reward"])
            code = int(hn["done"][0])
            self.agent_selection = f"player_{int(hn['agent'][0])}"
            self._cached = (self.agent_selection, hn["obs"][0].copy(), hn["mask"][0].copy())
            rw = hn["reward"][0]
        else:
            # the batch steps in lockstep: every other env of the backend plays its first legal action
            actions = np.argmax(_np(self.env.action_mask), axis=1).astype(np.uint8)
            actions[self.index] = int(action)
            self.env.step(actions)
            code = int(_np(self.env.done_code)[self.index])
            self.agent_selection = self._expected_agent()
            rw = _np(self.env.rewards)[self.index] if code != _lib.RUNNING else None
        if code != _lib.RUNNING:
            if code == _lib.DONE_ILLEGAL:
                self._cumulative_rewards[agent] = 0              # TerminateIllegalWrapper
            self.rewards = {a: float(rw[int(a.split("_")[-1])]) for a in self.agents}
            self.dones = {a: True for a in self.agents}
        self._accumulate_rewards()
        self._clear_rewards()
        self._dones_step_first()

    # ---- PettingZoo 1.14 AECEnv helpers ---------------------------------------------------
    def _accumulate_rewards(self):
        for a, r in self.rewards.items():
            self._cumulative_rewards[a] += r

    def _clear_rewards(self):
        for a in self.rewards:
            self.rewards[a] = 0

    def _dones_step_first(self):
        order = [a for a in self.agents if self.dones[a]]
        if order:
            self._skip_agent_selection = self.agent_selection
            self.agent_selection = order[0]
        return self.agent_selection

    def _was_done_step(self, action):
        if action is not None:
            raise ValueError("when an agent is done, the only valid action is None")
        agent = self.agent_selection
        assert self.dones[agent], "an agent that was not done as attempted to be removed"
        del self.dones[agent], self.rewards[agent], self._cumulative_rewards[agent], self.infos[agent]
        self.agents.remove(agent)
        order = [a for a in self.agents if self.dones[a]]
        if order:
            if self._skip_agent_selection is None:
                self._skip_agent_selection = self.agent_selection
            self.agent_selection = order[0]
        else:
            if self._skip_agent_selection is not None:
                self.agent_selection = self._skip_agent_selection
            self._skip_agent_selection = None
        self._clear_rewards()

    def render(self, mode="human"):
        if mode == "human":
            print(self.env.export(self.index, 1)[0].render_table())

    def close(self):
        pass


class RLlibDictEnv:
    """Multi-agent dict view of one AEC env: the protocol RLlib's `PettingZooEnv` wrapper gives the
    reference (`rlskyjo/models/train_model_simple_rllib.py:30-33` registers
    `PettingZooEnv(skyjo_env.env(**config))`).  `reset()` returns {agent on turn: observation};
    `step({agent: action})` plays that agent's action and returns (obs, reward, done, info) dicts for
    the agent that moves next -- or, when the game is over, for every seat at once (the wrapper steps
    the done agents with None itself) with `done["__all__"] = True`.  Restated from ray 1.9.2's
    wrapper (ray is not installed here: functional, not pinned, parity)."""

    def __init__(self, aec_env):
        self.env = aec_env
        self.agents = list(aec_env.possible_agents)
        self.observation_space = aec_env.observation_space(self.agents[0])
        self.action_space = aec_env.action_space(self.agents[0])

    def reset(self):
        self.env.reset()
        return {self.env.agent_selection: self.env.observe(self.env.agent_selection)}

    def step(self, action_dict):
        self.env.step(action_dict[self.env.agent_selection])
        obs_d, rew_d, done_d, info_d = {}, {}, {}, {}
        while self.env.agents:
            obs, rew, done, info = self.env.last()
            a = self.env.agent_selection
            obs_d[a], rew_d[a], done_d[a], info_d[a] = obs, rew, done, info
            if self.env.dones[a]:
                self.env.step(None)
            else:
                break
        done_d["__all__"] = not self.env.agents
        return obs_d, rew_d, done_d, info_d

    def seed(self, seed=None):
        self.env.seed(seed)

    def render(self, mode="human"):
        return self.env.render(mode)

    def close(self):
        self.env.close()


def env(**config):
    """Drop-in for `rlskyjo.environment.skyjo_env.env(**config)` (skyjo_env.py:19-26): one game on
    cuda:0 with the reference's keyword arguments (`num_players`, `score_penalty`,
    `observe_other_player_indirect`, `mean_reward`, `reward_refunded`)."""
    from .env import BatchedSkyjoEnv
    device = config.pop("device", "cuda:0")
    seed = config.pop("seed", 0)
    return SkyjoAECView(BatchedSkyjoEnv(num_envs=1, auto_reset=False, device=device, seed=seed, **config), 0)


def simple_episode(config, policy=None, rng=None, verbose=0):
    """The reference's canonical consumer loop (vanilla_env_example.py:6-41) on the GPU env.
    Returns the list of (agent, cumulative reward) seen with done=True."""
    from .policy import policy_ra
    e = env(**config)
    e.reset()
    finished = []
    for agent in e.agent_iter(max_iter=300 * config["num_players"]):
        obs, reward, done, info = e.last()
        if not done:
            action = (policy or policy_ra)(obs["observations"], obs["action_mask"], rng)
            e.step(action)
            if verbose:
                e.render()
        else:
            e.step(None)
            finished.append((agent, reward))
    return finished
