"""The wrapper stack of rlskyjo/environment/skyjo_env.py:19-26, restated from pettingzoo 1.14.0:
a wrapper forwards reset/step/observe to the wrapped env and re-reads its AEC attributes after
each call, so that every layer exposes the same `agent_selection / rewards / dones / infos /
agents / _cumulative_rewards` objects."""
import warnings

from .env import AECEnv

_SHARED = ("agent_selection", "rewards", "dones", "infos", "agents", "_cumulative_rewards")


class BaseWrapper(AECEnv):
    def __init__(self, env):
        super().__init__()
        self.env = env
        self.possible_agents = env.possible_agents
        self.metadata = env.metadata

    def _pull(self):
        for name in _SHARED:
            setattr(self, name, getattr(self.env, name))

    def observation_space(self, agent):
        return self.env.observation_space(agent)

    def action_space(self, agent):
        return self.env.action_space(agent)

    @property
    def unwrapped(self):
        return getattr(self.env, "unwrapped", self.env)

    def seed(self, seed=None):
        self.env.seed(seed)

    def close(self):
        self.env.close()

    def render(self, mode="human"):
        return self.env.render(mode)

    def reset(self):
        self.env.reset()
        self._pull()

    def observe(self, agent):
        return self.env.observe(agent)

    def step(self, action):
        self.env.step(action)
        self._pull()


class CaptureStdoutWrapper(BaseWrapper):
    """render(mode='human') output is captured and returned as a string; nothing on the step path."""


class TerminateIllegalWrapper(BaseWrapper):
    def __init__(self, env, illegal_reward):
        super().__init__(env)
        self._illegal_value = illegal_reward
        self._prev_obs = None

    def reset(self):
        self._terminated = False
        self._prev_obs = None
        super().reset()

    def observe(self, agent):
        obs = super().observe(agent)
        if agent == self.agent_selection:
            self._prev_obs = obs
        return obs

    def step(self, action):
        offender = self.agent_selection
        if self._prev_obs is None:
            self.observe(offender)
        assert "action_mask" in self._prev_obs
        mask = self._prev_obs["action_mask"]
        self._prev_obs = None
        if self._terminated and self.dones[offender]:
            self._was_done_step(action)
        elif not self.dones[offender] and not mask[action]:
            warnings.warn("[WARNING]: Illegal move made, game terminating with current player losing.")
            self._cumulative_rewards[offender] = 0
            self.dones = {a: True for a in self.dones}
            self.rewards = {a: 0 for a in self.dones}
            self.rewards[offender] = float(self._illegal_value)
            self._accumulate_rewards()
            self._dones_step_first()
            self._terminated = True
        else:
            super().step(action)


class AssertOutOfBoundsWrapper(BaseWrapper):
    def step(self, action):
        assert (action is None and self.dones[self.agent_selection]) or \
            self.action_space(self.agent_selection).contains(action), "action is not in action space"
        super().step(action)


class OrderEnforcingWrapper(BaseWrapper):
    def __init__(self, env):
        self._has_reset = False
        super().__init__(env)

    def observe(self, agent):
        assert self._has_reset, "reset() needs to be called before observe"
        return super().observe(agent)

    def step(self, action):
        assert self._has_reset, "reset() needs to be called before step"
        if not self.agents:
            warnings.warn("step() called after every agent is done; reset() first")
            return
        super().step(action)

    def reset(self):
        self._has_reset = True
        super().reset()
