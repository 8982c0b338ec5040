"""GPU tests of the callers either side of the hot path: the single-env AEC view
driven like the reference's vanilla_env_example loop, and the torch action-mask policy reading the
env's device buffers in place (BASELINE config 4)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def test_simple_episode_runs_like_the_reference_loop():
    # reference tests/environment/test_skyjo_env_nojit.py sweeps configs through simple_episode
    from skyjo_rl_b200.aec import simple_episode
    rng = np.random.default_rng(0)
    for N in (1, 2, 3, 5):
        for indirect in (False, True):
            cfg = {"num_players": N, "score_penalty": 2.0, "observe_other_player_indirect": indirect,
                   "mean_reward": 1.0, "reward_refunded": 0.0}
            finished = simple_episode(cfg, rng=rng)
            assert [a for a, _ in finished] == [f"player_{i}" for i in range(N)]
            # skyjo_env.py:307-308: rewards sum to N * mean_reward
            assert abs(sum(r for _, r in finished) - N * 1.0) < 1e-9


def test_aec_seeded_runs_are_reproducible():
    # reference tests/environment/test_skyjo_env_jit.py::test_reproducability
    from skyjo_rl_b200.aec import env as make_env
    from skyjo_rl_b200.policy import policy_ra

    def run():
        e = make_env(num_players=3, score_penalty=2.0, observe_other_player_indirect=False,
                     mean_reward=1.0, reward_refunded=0.0)
        e.seed(42)
        rng = np.random.default_rng(42)
        obs_l, rew_l = [], []
        for agent in e.agent_iter(max_iter=900):
            obs, reward, done, info = e.last()
            obs_l.append(obs["observations"].tolist())
            rew_l.append(reward)
            e.step(None if done else policy_ra(obs["observations"], obs["action_mask"], rng))
        return obs_l, rew_l
    a, b = run(), run()
    assert a == b and len(a[0]) > 60


def test_action_mask_policy_rollout_is_zero_copy_and_legal():
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.policy import ActionMaskPolicy, rollout
    B, N, T = 4096, 4, 96
    env = BatchedSkyjoEnv(num_envs=B, num_players=N, seed=11)
    env.reset()
    ptrs = (env.observations.data_ptr(), env.action_mask.data_ptr(), env.rewards.data_ptr())
    torch.manual_seed(0)
    pol = ActionMaskPolicy(env.obs_len).to(env.device)
    buf = rollout(pol, env, T)
    assert ptrs == (env.observations.data_ptr(), env.action_mask.data_ptr(), env.rewards.data_ptr())
    st = env.stats()
    assert st["illegal"] == 0 and st["steps"] == B * T          # masked logits never pick an illegal action
    legal = buf["mask"].gather(2, buf["action"].long().unsqueeze(2)).squeeze(2)
    assert bool((legal == 1).all())
    assert buf["obs"].shape == (T, B, env.obs_len) and buf["reward"].shape == (T, B, N)
    ended = buf["done"] == 1
    assert int(ended.sum()) == st["episodes"]
    # rewards are published only on the step that ends an episode
    assert bool((buf["reward"][~ended] == 0).all())
    # value head and logits are finite; masked logits are hugely negative where illegal
    logits = pol({"observations": env.observations, "action_mask": env.action_mask})
    assert bool(torch.isfinite(pol.value_function()).all())
    assert bool((logits[env.action_mask == 0] < -1e30).all())
    env.check()


def test_gpu_aec_replays_notebook_trace():
    # the reference's recorded episode (notebooks/trainpettingzoo.ipynb cell 5) through the CUDA env
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.aec import SkyjoAECView
    from test_notebook_trace import load_trace, replay_through_aec
    env = BatchedSkyjoEnv(num_envs=1, num_players=3, score_penalty=2.0, mean_reward=1.0, reward_refunded=0.0,
                          auto_reset=False)
    replay_through_aec(SkyjoAECView(env, 0), load_trace())
    v = env.game_view(0)
    assert v.is_terminated and "u1" in v.render_table()          # the one card that stayed hidden (value 1)


def test_checkpoint_restore_resumes_bit_identically():
    # env-state checkpoint / resume
    from skyjo_rl_b200 import BatchedSkyjoEnv
    kw = dict(num_envs=3000, num_players=4, seed=99, reward_refunded=0.01)
    a = BatchedSkyjoEnv(**kw)
    a.reset()
    a.step_random(203)
    sd = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in a.state_dict().items()}
    a.step_random(150)
    b = BatchedSkyjoEnv(**kw)
    b.reset()
    b.load_state_dict({k: (v.to(b.device) if torch.is_tensor(v) else v) for k, v in sd.items()})
    b.step_random(150)
    assert torch.equal(a.observations, b.observations) and torch.equal(a.action_mask, b.action_mask)
    assert torch.equal(a.rewards, b.rewards) and torch.equal(a.done_code, b.done_code)
    assert torch.equal(a.agent_selection, b.agent_selection)
    a.check()
    b.check()
