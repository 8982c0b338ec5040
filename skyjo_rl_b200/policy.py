"""Policies that consume the env's device buffers in place (BASELINE config 4).

* `policy_ra` -- the reference's uniform random admissible policy for one observation
  (rlskyjo/models/random_admissible_policy.py:6-28), numpy, used by the AEC view.
* `ActionMaskPolicy` -- plain-torch restatement of `TorchActionMaskModel`
  (rlskyjo/models/action_mask_model.py:13-77): RLlib's `TorchFC` defaults (two tanh layers of
  256, separate value branch) on `obs["observations"].float()`, logits + clamp(log(mask), FLOAT_MIN)
  (:63-71), value head (:76-77).  ray is not installed: functional, not bit, parity.
* `FusedPolicy` -- the same network as ONE sm_100a kernel (csrc/skyjo_policy.cu: tcgen05.mma on bf16 copies of the
  weights held in shared memory, activations in tensor memory, tanh + masked softmax + Philox sample in the
  epilogues): int8 observations in, uint8 actions out, nothing in between touches HBM.
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


class FusedPolicy:
    """`ActionMaskPolicy` evaluated by the library's fused tensor-core kernel on an env's live buffers.

        fused = FusedPolicy(policy, env)          # packs bf16 copies of the weights (repack() after an update)
        actions, logp = fused.sample(seed)        # uint8 [B] for env.step, float32 [B]
        values = fused.value()                    # float32 [B] from the value branch

    bf16 operands with fp32 accumulation: functional parity with the fp32 module (the tolerance is measured in
    tests/test_gpu_fused_policy.py), not bit parity.  Observation rows of at most 96 bytes."""

    def __init__(self, policy, env, with_value=True):
        from . import _lib
        self._lib, self._L = _lib, _lib.load()
        self.policy, self.env = policy, env
        n = int(self._L.skyjo_policy_packed_bytes())
        self._packed = torch.empty(n, dtype=torch.uint8, device=env.device)
        self._packed_v = torch.empty(n, dtype=torch.uint8, device=env.device) if with_value else None
        self.repack()

    def _pack(self, net, n_out, dst):
        lin = [m for m in net if isinstance(m, nn.Linear)]
        assert len(lin) == 3 and lin[0].out_features == 256 and lin[1].out_features == 256 and lin[2].out_features == n_out
        ws = [t.detach().to(self.env.device, torch.float32).contiguous()
              for m in lin for t in (m.weight, m.bias)]
        self._lib.check(self._L.skyjo_policy_pack(self.env.obs_len, n_out, *[w.data_ptr() for w in ws],
                                                  dst.data_ptr(), self.env._stream()))
        st = self.env.stream if self.env.stream is not None else torch.cuda.current_stream(self.env.device)
        st.synchronize()                                               # the fp32 copies die with this frame

    def repack(self):
        self._pack(self.policy.logits_net, 26, self._packed)
        if self._packed_v is not None:
            self._pack(self.policy.value_net, 1, self._packed_v)

    def sample(self, seed=0, actions=None, logp=None, entropy=None, logits=None):
        env, B = self.env, self.env.num_envs
        actions = torch.empty(B, dtype=torch.uint8, device=env.device) if actions is None else actions
        logp = torch.empty(B, dtype=torch.float32, device=env.device) if logp is None else logp
        self._lib.check(self._L.skyjo_policy_sample(
            env._h, self._packed.data_ptr(), int(seed), actions.data_ptr(), logp.data_ptr(),
            entropy.data_ptr() if entropy is not None else None, logits.data_ptr() if logits is not None else None,
            env._stream()))
        return actions, logp

    def value(self, out=None):
        env = self.env
        out = torch.empty(env.num_envs, dtype=torch.float32, device=env.device) if out is None else out
        self._lib.check(self._L.skyjo_policy_value(env._h, self._packed_v.data_ptr(), out.data_ptr(), env._stream()))
        return out

    def debug(self):
        """(pre-activations of layer 1, of layer 2, unmasked logits) as float32 tensors -- parity tests"""
        env, B = self.env, self.env.num_envs
        pre1 = torch.empty((B, 256), dtype=torch.float32, device=env.device)
        pre2 = torch.empty((B, 256), dtype=torch.float32, device=env.device)
        logits = torch.empty((B, 26), dtype=torch.float32, device=env.device)
        self._lib.check(self._L.skyjo_policy_debug(env._h, self._packed.data_ptr(), pre1.data_ptr(), pre2.data_ptr(),
                                                   logits.data_ptr(), env._stream()))
        return pre1, pre2, logits
