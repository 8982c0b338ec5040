"""Policies that consume the env's device buffers in place (BASELINE config 4).

* `policy_ra` -- the reference's uniform random admissible policy for one observation
  (rlskyjo/models/random_admissible_policy.py:6-28), numpy, used by the AEC view.
* `ActionMaskPolicy` -- plain-torch restatement of `TorchActionMaskModel`
  (rlskyjo/models/action_mask_model.py:13-77): RLlib's `TorchFC` defaults (two tanh layers of
  256, separate value branch) on `obs["observations"].float()`, logits + clamp(log(mask), FLOAT_MIN)
  (:63-71), value head (:76-77).  ray is not installed: functional, not bit, parity.
* `sample_actions` / `rollout` -- masked categorical sampling straight from `env.observations`
  and `env.action_mask` (zero copy) into a uint8 action tensor that `env.step` hands to the
  fused kernel, and a `[T, B, ...]` rollout recorder for a PPO learner.
"""
import numpy as np
import torch
from torch import nn

FLOAT_MIN = torch.finfo(torch.float32).min  # ray.rllib.utils.torch_ops.FLOAT_MIN


def policy_ra(observation, action_mask, rng=None):
    """random_admissible_policy.py:6-28"""
    p = np.asarray(action_mask, dtype=np.float64)
    p = p / p.sum()
    if rng is None:
        return int(np.random.choice(np.arange(26), p=p))
    return int(rng.choice(np.arange(26), p=p))


class ActionMaskPolicy(nn.Module):
    def __init__(self, obs_len, num_actions=26, hiddens=(256, 256)):
        super().__init__()

        def mlp(out):
            layers, d = [], obs_len
            for h in hiddens:
                layers += [nn.Linear(d, h), nn.Tanh()]
                d = h
            layers.append(nn.Linear(d, out))
            return nn.Sequential(*layers)
        self.logits_net = mlp(num_actions)
        self.value_net = mlp(1)          # vf_share_layers=False
        self._features = None

    def forward(self, obs):
        """obs: {"observations": int8/float [B, D], "action_mask": int8/float [B, 26]} -> masked logits"""
        x = obs["observations"].float()
        self._features = x
        logits = self.logits_net(x)
        inf_mask = torch.clamp(torch.log(obs["action_mask"].float()), min=FLOAT_MIN)  # action_mask_model.py:70
        return logits + inf_mask

    def value_function(self):
        return self.value_net(self._features).squeeze(-1)


@torch.no_grad()
def sample_actions(policy, env, generator=None):
    """Masked categorical sample for every env, reading the env's live buffers in place."""
    logits = policy({"observations": env.observations, "action_mask": env.action_mask})
    probs = torch.softmax(logits, dim=-1)
    a = torch.multinomial(probs, 1, generator=generator).squeeze(1)
    logp = torch.log(probs.gather(1, a.unsqueeze(1)).squeeze(1))
    return a.to(torch.uint8), logp, policy.value_function()


@torch.no_grad()
def rollout(policy, env, T, generator=None):
    """T lockstep steps; returns a dict of [T, B, ...] tensors (obs/mask are copied out of the
    env's buffers, which the kernel overwrites every step)."""
    B, D, N = env.num_envs, env.obs_len, env.num_players
    dev = env.device
    buf = {
        "obs": torch.empty((T, B, D), dtype=torch.int8, device=dev),
        "mask": torch.empty((T, B, 26), dtype=torch.int8, device=dev),
        "agent": torch.empty((T, B), dtype=torch.int8, device=dev),
        "action": torch.empty((T, B), dtype=torch.uint8, device=dev),
        "logp": torch.empty((T, B), dtype=torch.float32, device=dev),
        "value": torch.empty((T, B), dtype=torch.float32, device=dev),
        "done": torch.empty((T, B), dtype=torch.uint8, device=dev),
        "reward": torch.empty((T, B, N), dtype=torch.float64, device=dev),
    }
    for t in range(T):
        buf["obs"][t].copy_(env.observations)
        buf["mask"][t].copy_(env.action_mask)
        buf["agent"][t].copy_(env.agent_selection)
        a, logp, v = sample_actions(policy, env, generator)
        env.step(a)
        buf["action"][t], buf["logp"][t], buf["value"][t] = a, logp, v
        buf["done"][t].copy_(env.done_code)
        buf["reward"][t].copy_(env.rewards)
    return buf
