"""Plain-torch PPO learner on top of the batched env (BASELINE config 4).

It replaces what `rlskyjo/models/train_model_simple_rllib.py:34-78` delegates to RLlib's PPOTrainer
(one shared `TorchActionMaskModel` policy for every seat, `action_mask_model.py:58-77`):

* `RolloutStorage` -- time-major tensors `[T(+1), B, ...]` that the env's kernels write IN PLACE:
  before step t the env's outputs are re-bound (`skyjo_bind_outputs`) to slice t + 1, so the fused
  step kernel stores the next observation / mask / agent, the done code and the rewards straight
  into the learner's memory (no copy, `obs[t].data_ptr()` never changes).
* actions are drawn by the fused masked-softmax-sample kernel (`csrc/skyjo_sample.cuh`) from the
  policy's logits and the mask slice, as uint8 that `skyjo_step` consumes.
* `gae_turn_based` -- generalized advantage estimation per (env, seat): SkyJo is turn based (a seat
  acts twice in a row: draw, then place) and every seat is paid only when the game ends
  (`skyjo_env.py:242-247`), so the successor of a transition is the SAME seat's next decision and
  the final rewards of all seats are credited to their last decisions.
* `PPOTrainer` -- clipped-surrogate PPO with value clipping, entropy bonus, minibatch epochs
  (defaults follow the hyper-parameters visible in `notebooks/trainpettingzoo.ipynb:2917,3047`:
  lr 5e-5, clip 0.3, vf_clip 10, lambda 1.0, gamma 0.99) and a gradient all-reduce over
  `torch.distributed` when initialised (one process per GPU, envs sharded by global env id).

ray / RLlib are not installed here, so this is a functional replacement, not a bit-parity one.
"""
import torch

from .policy import FLOAT_MIN, ActionMaskPolicy


class RolloutStorage:
    def __init__(self, env, T):
        B, N, D, dev = env.num_envs, env.num_players, env.obs_len, env.device
        self.T, self.B, self.N = int(T), B, N
        z = dict(device=dev)
        self.obs = torch.empty((T + 1, B, D), dtype=torch.int8, **z)
        self.mask = torch.empty((T + 1, B, 26), dtype=torch.int8, **z)
        self.agent = torch.empty((T + 1, B), dtype=torch.int8, **z)
        self.action = torch.empty((T, B), dtype=torch.uint8, **z)
        self.logp = torch.empty((T, B), dtype=torch.float32, **z)
        self.value = torch.empty((T + 1, B), dtype=torch.float32, **z)
        self.done = torch.zeros((T, B), dtype=torch.uint8, **z)
        self.reward = torch.zeros((T, B, N), dtype=torch.float64, **z)
        self._final = torch.zeros((B, N), dtype=torch.float64, **z)
        self._primed = False

    def prime(self, env):
        """Slot 0 <- the env's current turn (first use), or the last slot of the previous rollout."""
        if not self._primed:
            self.obs[0].copy_(env.observations)
            self.mask[0].copy_(env.action_mask)
            self.agent[0].copy_(env.agent_selection)
            self._primed = True
        else:
            self.obs[0].copy_(self.obs[self.T])
            self.mask[0].copy_(self.mask[self.T])
            self.agent[0].copy_(self.agent[self.T])
        # reward rows are written only for envs whose episode ends in that step (and cleared one
        # step later through whatever slice is bound then): start every rollout from zeros
        self.reward.zero_()


@torch.no_grad()
def collect(policy, env, storage, sample_seed=0, fused=None):
    """T lockstep steps of every env under `policy`; the env writes slices 1..T of the storage.
    With `fused` (a skyjo_rl_b200.policy.FusedPolicy of the same policy) the whole policy side of a step -- forward,
    masked softmax, sample, log-probability, and the value branch -- runs in the library's tensor-core kernels on
    the bound observation slice; otherwise in ATen."""
    st = storage
    st.prime(env)
    for t in range(st.T):
        if fused is not None:
            # the env's bound outputs ARE slice t (slice T of the previous rollout, copied to slot 0, for t = 0)
            fused.sample(sample_seed, actions=st.action[t], logp=st.logp[t])
            fused.value(out=st.value[t])
        else:
            logits = policy({"observations": st.obs[t], "action_mask": st.mask[t]})
            st.value[t].copy_(policy.value_function())
            env.sample_actions(logits.contiguous(), st.mask[t], seed=sample_seed, actions=st.action[t], logp=st.logp[t])
        env.bind_outputs(observations=st.obs[t + 1], action_mask=st.mask[t + 1], agent_selection=st.agent[t + 1],
                         done_code=st.done[t], rewards=st.reward[t], final_scores=st._final)
        env.step(st.action[t])
    if fused is not None:
        fused.value(out=st.value[st.T])
    else:
        policy({"observations": st.obs[st.T], "action_mask": st.mask[st.T]})
        st.value[st.T].copy_(policy.value_function())
    return st


def gae_turn_based(value, agent, done, reward, gamma=0.99, lam=1.0):
    """Advantages / returns / validity of the T transitions of every env.

    value f32[T+1,B] (value[t] = V of the acting seat's observation at t, value[T] bootstraps the
    seat that acts next), agent int[T+1,B], done uint8[T,B], reward [T,B,N] (non-zero only where an
    episode ended).  The successor of seat q's decision is q's NEXT decision in the same episode;
    when the episode ends every seat's pending decision receives its final reward with successor
    value 0.  Decisions whose successor lies beyond the horizon (the last decision of the seats
    that are not on turn at T) are marked invalid."""
    T, B = done.shape
    N = reward.shape[2]
    dev = value.device
    idx = torch.arange(B, device=dev)
    nv = torch.zeros((B, N), dtype=torch.float32, device=dev)      # value of the seat's next decision
    nadv = torch.zeros((B, N), dtype=torch.float32, device=dev)    # its advantage (for the lambda chain)
    known = torch.zeros((B, N), dtype=torch.bool, device=dev)
    carry = torch.zeros((B, N), dtype=torch.float32, device=dev)   # reward received since the seat's decision
    pT = agent[T].long()
    nv[idx, pT] = value[T]
    known[idx, pT] = True
    adv = torch.empty((T, B), dtype=torch.float32, device=dev)
    valid = torch.empty((T, B), dtype=torch.bool, device=dev)
    for t in range(T - 1, -1, -1):
        d = done[t] != 0
        dm = d.unsqueeze(1)
        carry = torch.where(dm, reward[t].to(torch.float32), carry)
        nv = torch.where(dm, torch.zeros_like(nv), nv)
        nadv = torch.where(dm, torch.zeros_like(nadv), nadv)
        known = known | dm
        p = agent[t].long()
        r = carry[idx, p]
        ok = known[idx, p]
        delta = r + gamma * nv[idx, p] - value[t]
        a = delta + gamma * lam * nadv[idx, p]
        a = torch.where(ok, a, torch.zeros_like(a))
        adv[t] = a
        valid[t] = ok
        carry[idx, p] = 0.0
        nv[idx, p] = value[t]
        nadv[idx, p] = a
        known[idx, p] = True
    ret = adv + value[:T]
    return adv, ret, valid


class PPOTrainer:
    def __init__(self, env, policy=None, rollout_len=64, lr=5e-5, gamma=0.99, lam=1.0, clip=0.3, vf_clip=10.0,
                 vf_coef=1.0, ent_coef=0.0, epochs=4, minibatches=8, max_grad_norm=None, seed=0, fused=False):
        self.env = env
        self.policy = policy if policy is not None else ActionMaskPolicy(env.obs_len).to(env.device)
        # fused=True: rollouts through the library's tcgen05 policy kernel (bf16 copies of the weights, repacked after
        # every update; rows of at most 96 bytes); the PPO epochs themselves stay in fp32 autograd
        self.fused = None
        if fused:
            from .policy import FusedPolicy
            self.fused = FusedPolicy(self.policy, env, with_value=True)
        self.storage = RolloutStorage(env, rollout_len)
        self.opt = torch.optim.Adam(self.policy.parameters(), lr=lr)
        self.gamma, self.lam, self.clip, self.vf_clip = gamma, lam, clip, vf_clip
        self.vf_coef, self.ent_coef, self.epochs, self.minibatches = vf_coef, ent_coef, epochs, minibatches
        self.max_grad_norm = max_grad_norm
        self.seed = int(seed)
        self.iteration = 0
        self._gen = torch.Generator(device=env.device)
        self._gen.manual_seed(self.seed)

    def _sync_grads(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            w = dist.get_world_size()
            for p in self.policy.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad.div_(w)

    def evaluate(self, obs, mask, action):
        logits = self.policy({"observations": obs, "action_mask": mask})
        logp_all = torch.log_softmax(logits, dim=-1)
        logp = logp_all.gather(1, action.long().unsqueeze(1)).squeeze(1)
        legal = mask != 0
        p = torch.where(legal, logp_all.exp(), torch.zeros_like(logp_all))
        ent = -(p * torch.where(legal, logp_all, torch.zeros_like(logp_all))).sum(-1)
        return logp, ent, self.policy.value_function()

    def train_iteration(self):
        env, st = self.env, self.storage
        s0 = env.stats()
        if self.fused is not None:
            self.fused.repack()
        collect(self.policy, env, st, sample_seed=self.seed, fused=self.fused)
        s1 = env.stats()
        adv, ret, valid = gae_turn_based(st.value, st.agent, st.done, st.reward, self.gamma, self.lam)
        T, B = st.T, st.B
        sel = valid.reshape(-1).nonzero(as_tuple=False).squeeze(1)
        obs = st.obs[:T].reshape(T * B, -1)
        mask = st.mask[:T].reshape(T * B, 26)
        act = st.action.reshape(-1)
        logp_old = st.logp.reshape(-1)
        v_old = st.value[:T].reshape(-1)
        adv_f, ret_f = adv.reshape(-1), ret.reshape(-1)
        a_sel = adv_f[sel]
        # advantage moments over ALL ranks (every rank optimises the same objective): sum, sum of squares, count
        mom = torch.stack([a_sel.sum(), (a_sel * a_sel).sum(), a_sel.new_tensor(float(a_sel.numel()))]).double()
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(mom)
        n_all = mom[2].clamp(min=1.0)
        a_mean = mom[0] / n_all
        a_std = ((mom[1] / n_all - a_mean * a_mean).clamp(min=0.0) * n_all / (n_all - 1.0).clamp(min=1.0)).sqrt()
        adv_n = (adv_f - a_mean.to(adv_f.dtype)) / (a_std.to(adv_f.dtype) + 1e-8)
        out = {"policy_loss": 0.0, "vf_loss": 0.0, "entropy": 0.0, "kl": 0.0}
        n_upd = 0
        for _ in range(self.epochs):
            perm = sel[torch.randperm(sel.numel(), device=sel.device, generator=self._gen)]
            # exactly `minibatches` chunks on every rank, whatever its count of valid transitions: the gradient
            # all-reduces of the ranks must pair up one to one (an empty chunk contributes a zero gradient)
            for i in torch.tensor_split(perm, self.minibatches):
                if i.numel() == 0:
                    self.opt.zero_grad(set_to_none=True)
                    for prm in self.policy.parameters():
                        prm.grad = torch.zeros_like(prm)
                    self._sync_grads()
                    self.opt.step()
                    continue
                logp, ent, v = self.evaluate(obs[i], mask[i], act[i])
                ratio = torch.exp(logp - logp_old[i])
                a = adv_n[i]
                pl = -torch.min(ratio * a, torch.clamp(ratio, 1 - self.clip, 1 + self.clip) * a).mean()
                verr = torch.clamp((v - ret_f[i]) ** 2, max=self.vf_clip)   # RLlib's vf_clip_param
                vl = verr.mean()
                loss = pl + self.vf_coef * vl - self.ent_coef * ent.mean()
                self.opt.zero_grad(set_to_none=True)
                loss.backward()
                self._sync_grads()
                if self.max_grad_norm:
                    torch.nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
                self.opt.step()
                with torch.no_grad():
                    out["policy_loss"] += float(pl)
                    out["vf_loss"] += float(vl)
                    out["entropy"] += float(ent.mean())
                    out["kl"] += float((logp_old[i] - logp).mean())
                n_upd += 1
        for k in out:
            out[k] /= max(n_upd, 1)
        eps = s1["episodes"] - s0["episodes"]
        out.update({
            "iteration": self.iteration, "env_steps": T * B, "transitions_used": int(sel.numel()),
            "episodes": eps, "illegal": s1["illegal"] - s0["illegal"],
            "mean_episode_len": (s1["episode_steps"] - s0["episode_steps"]) / max(eps, 1),
            # mean unpenalised score per seat of the finished games (lower = better play, skyjo.py:477-498)
            "mean_raw_score": (s1["score_raw_sum"] - s0["score_raw_sum"]) / max(eps * env.num_players, 1),
            "mean_winner_score": (s1["winner_raw_sum"] - s0["winner_raw_sum"]) / max(eps, 1),
        })
        self.iteration += 1
        return out

    def state_dict(self):
        """Policy + optimiser + env checkpoint (resumable rollouts)."""
        import copy
        return {"policy": {k: v.detach().clone() for k, v in self.policy.state_dict().items()},
                "opt": copy.deepcopy(self.opt.state_dict()), "iteration": self.iteration,
                "env": self.env.state_dict(), "gen": self._gen.get_state().clone()}

    def load_state_dict(self, sd):
        self.policy.load_state_dict(sd["policy"])
        self.opt.load_state_dict(sd["opt"])
        self.iteration = sd["iteration"]
        self.env.load_state_dict(sd["env"])
        self._gen.set_state(sd["gen"])
        self.storage._primed = False


__all__ = ["RolloutStorage", "collect", "gae_turn_based", "PPOTrainer", "FLOAT_MIN"]
