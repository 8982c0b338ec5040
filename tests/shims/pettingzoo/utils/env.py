"""AECEnv of pettingzoo 1.14.0, restated (agent-environment-cycle bookkeeping only)."""


class AECEnv:
    def __init__(self):
        pass

    @property
    def num_agents(self):
        return len(self.agents)

    @property
    def max_num_agents(self):
        return len(self.possible_agents)

    def _dones_step_first(self):
        # the first done agent takes the turn; the interrupted selection is remembered
        done_agents = [a for a in self.agents if self.dones[a]]
        if done_agents:
            self._skip_agent_selection = self.agent_selection
            self.agent_selection = done_agents[0]
        return self.agent_selection

    def _clear_rewards(self):
        for a in self.rewards:
            self.rewards[a] = 0

    def _accumulate_rewards(self):
        for a, r in self.rewards.items():
            self._cumulative_rewards[a] += r

    def agent_iter(self, max_iter=2 ** 63):
        left = max_iter
        while self.agents and left > 0:
            left -= 1
            yield self.agent_selection

    def last(self, observe=True):
        a = self.agent_selection
        obs = self.observe(a) if observe else None
        return obs, self._cumulative_rewards[a], self.dones[a], self.infos[a]

    def _was_done_step(self, action):
        if action is not None:
            raise ValueError("when an agent is done, the only valid action is None")
        a = self.agent_selection
        assert self.dones[a], "an agent that was not done as attempted to be removed"
        del self.dones[a]
        del self.rewards[a]
        del self._cumulative_rewards[a]
        del self.infos[a]
        self.agents.remove(a)
        done_agents = [x for x in self.agents if self.dones[x]]
        if done_agents:
            if getattr(self, "_skip_agent_selection", None) is None:
                self._skip_agent_selection = self.agent_selection
            self.agent_selection = done_agents[0]
        else:
            if getattr(self, "_skip_agent_selection", None) is not None:
                self.agent_selection = self._skip_agent_selection
            self._skip_agent_selection = None
        self._clear_rewards()
