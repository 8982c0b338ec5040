"""The fused tensor-core policy kernel (csrc/skyjo_policy.cu) against the plain-torch fp32 restatement of the
reference's TorchActionMaskModel (skyjo_rl_b200/policy.py, reference rlskyjo/models/action_mask_model.py:58-77).

The kernel computes in bf16 x bf16 -> fp32 (tcgen05.mma), tanh.approx.f32, activations rounded to bf16 between
layers: FUNCTIONAL parity, with these tolerances (measured on B200, stated here as the contract):
  * layer-1 pre-activations: the observations are exact in bf16, so the only error is the weights' rounding:
    |d| <= 2^-8 * sum_k |x_k w_k| per element (checked against that bound), in practice < 0.05 at |x| <= 127;
  * against an fp32 evaluation of the SAME bf16-rounded weights and bf16-rounded activations (what the kernel is
    meant to compute): logits within 2e-2 absolute;
  * against the unrounded fp32 module: logits within 0.15 absolute, mean 0.02 (bf16 has 8 bits of mantissa);
  * log-probability of the drawn action == log_softmax of the kernel's own masked logits to 1e-4; the action is
    always legal; equal logits draw the action skyjo_sample_actions draws.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _setup(N, B, indirect=False, seed=0, steps=37):
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.policy import ActionMaskPolicy, FusedPolicy
    env = BatchedSkyjoEnv(num_envs=B, num_players=N, observe_other_player_indirect=indirect, seed=seed)
    env.reset()
    env.step_random(steps)              # a mix of draw and place turns, open cards, some removed columns
    torch.manual_seed(seed)
    policy = ActionMaskPolicy(env.obs_len).to(env.device)
    return env, policy, FusedPolicy(policy, env)


def _bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)


def _reference_bf16(policy, x):
    """fp32 evaluation of what the kernel is meant to compute: bf16-rounded weights and activations, fp32 sums"""
    lin = [m for m in policy.logits_net if isinstance(m, torch.nn.Linear)]
    pre1 = x @ _bf16(lin[0].weight).T + lin[0].bias
    h1 = _bf16(torch.tanh(pre1))
    pre2 = h1 @ _bf16(lin[1].weight).T + lin[1].bias
    h2 = _bf16(torch.tanh(pre2))
    return pre1, pre2, h2 @ _bf16(lin[2].weight).T + lin[2].bias


@pytest.mark.parametrize("N,B,indirect", [(4, 128, False), (4, 128 * 5 + 37, False), (2, 4096, False), (3, 1000, True),
                                           (6, 300, False), (1, 64, False)])
def test_fused_policy_layers_match_fp32_evaluation(N, B, indirect):
    env, policy, fused = _setup(N, B, indirect)
    x = env.observations.float()
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        r1, r2, r3 = _reference_bf16(policy, x)
        full = policy.logits_net(x)
    pre1, pre2, logits = fused.debug()
    torch.cuda.synchronize()
    # layer 1: exact inputs, bf16 weights -- the kernel and the bf16 reference compute the same products in fp32
    assert float((pre1 - r1).abs().max()) < 2e-3, "layer 1 (A operand from tensor memory, W1 descriptor)"
    assert float((pre2 - r2).abs().max()) < 2e-2, "layer 2"
    assert float((logits - r3).abs().max()) < 2e-2, "layer 3"
    d = (logits - full).abs()
    assert float(d.max()) < 0.15 and float(d.mean()) < 0.03, (float(d.max()), float(d.mean()))


@pytest.mark.parametrize("N,B", [(4, 4096 + 5), (2, 1 << 16)])
def test_fused_policy_sampling_is_masked_and_consistent(N, B):
    env, policy, fused = _setup(N, B, seed=3)
    logits = torch.empty((B, 26), dtype=torch.float32, device=env.device)
    ent = torch.empty(B, dtype=torch.float32, device=env.device)
    actions, logp = fused.sample(seed=11, entropy=ent, logits=logits)
    torch.cuda.synchronize()
    a = actions.long()
    assert bool((env.action_mask.gather(1, a.unsqueeze(1)) == 1).all()), "an illegal action was drawn"
    masked = logits + torch.clamp(torch.log(env.action_mask.float()), min=torch.finfo(torch.float32).min)
    ls = torch.log_softmax(masked, dim=-1)
    assert float((ls.gather(1, a.unsqueeze(1)).squeeze(1) - logp).abs().max()) < 1e-4
    p = ls.exp()
    assert float((-(p * torch.where(p > 0, ls, torch.zeros_like(ls))).sum(-1) - ent).abs().max()) < 1e-3
    # the same logits through the stand-alone sample kernel draw the same actions (same Philox keying)
    a2, logp2 = env.sample_actions(logits, seed=11)
    assert torch.equal(a2, actions) and float((logp2 - logp).abs().max()) < 1e-5
    # the draws follow the distribution: mean log-probability of the drawn actions ~ minus the mean entropy
    assert abs(float(logp.mean()) + float(ent.mean())) < 0.02
    # a different seed draws different actions; the same seed the same
    a3, _ = fused.sample(seed=12)
    a4, _ = fused.sample(seed=11)
    assert not torch.equal(a3, actions) and torch.equal(a4, actions)
    # the actions drive the env: a legal step for every env
    env.step(actions)
    env.check()
    assert env.stats()["illegal"] == 0


def test_fused_value_head_and_rollout_loop():
    env, policy, fused = _setup(4, 1 << 14, seed=5)
    with torch.no_grad():
        ref = policy.value_net(env.observations.float()).squeeze(-1)
    v = fused.value()
    torch.cuda.synchronize()
    assert float((v - ref).abs().max()) < 0.15 and float((v - ref).abs().mean()) < 0.03
    ptr = (env.observations.data_ptr(), env.action_mask.data_ptr())
    for t in range(200):                 # the config-4 loop: fused policy kernel -> step kernel, zero copy
        a, _ = fused.sample(seed=t)
        env.step(a)
    env.check()
    st = env.stats()
    assert st["illegal"] == 0 and st["episodes"] > 0
    assert ptr == (env.observations.data_ptr(), env.action_mask.data_ptr())


def test_fused_policy_rejects_long_rows():
    from skyjo_rl_b200 import BatchedSkyjoEnv, _lib
    from skyjo_rl_b200.policy import ActionMaskPolicy, FusedPolicy
    env = BatchedSkyjoEnv(num_envs=64, num_players=8, seed=0)       # D = 115 > 96
    env.reset()
    with pytest.raises(_lib.SkyjoError):
        FusedPolicy(ActionMaskPolicy(env.obs_len).to(env.device), env)


def test_ppo_learns_with_fused_rollouts():
    """The learner of config 4 with its rollouts through the fused kernel (bf16 sampling, fp32 PPO epochs; the
    importance ratio absorbs the difference): self-play must still learn, as with ATen rollouts
    (tests/test_gpu_policy.py::test_ppo_self_play_learns_to_lower_scores)."""
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.ppo import PPOTrainer
    torch.manual_seed(0)
    env = BatchedSkyjoEnv(num_envs=8192, num_players=3, seed=1, observe_other_player_indirect=True,
                          reward_refunded=0.001)
    env.reset()
    tr = PPOTrainer(env, rollout_len=64, lr=3e-4, epochs=4, minibatches=8, ent_coef=0.01, fused=True)
    hist = [tr.train_iteration() for _ in range(14)]
    first = np.mean([h["mean_raw_score"] for h in hist[1:3]])
    last = np.mean([h["mean_raw_score"] for h in hist[-2:]])
    assert sum(h["illegal"] for h in hist) == 0
    assert all(abs(h["kl"]) < 0.5 for h in hist)            # bf16 behaviour policy vs fp32 learner: a small, finite gap
    assert first > 45 and last < 0.75 * first, (first, last)
    env.check()
