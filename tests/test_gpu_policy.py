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


def test_fused_masked_sample_kernel_against_torch():
    """csrc/skyjo_sample.cuh: masked softmax + categorical sample.  Tolerance (float32): logp and
    entropy within 2e-5 of torch's log_softmax on logits + clamp(log(mask), FLOAT_MIN)
    (action_mask_model.py:63-71); actions always legal; deterministic in (seed, env, t)."""
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.policy import FLOAT_MIN
    B = 1 << 15
    env = BatchedSkyjoEnv(num_envs=B, num_players=4, seed=5)
    env.reset()
    env.step_random(37)                      # a mix of draw- and place-phase masks
    g = torch.Generator(device=env.device).manual_seed(1)
    logits = 3.0 * torch.randn((B, 26), device=env.device, generator=g)
    mask = env.action_mask
    ent = torch.empty(B, dtype=torch.float32, device=env.device)
    a, logp = env.sample_actions(logits, seed=9, entropy=ent)
    a2, logp2 = env.sample_actions(logits, seed=9)
    assert torch.equal(a, a2) and torch.equal(logp, logp2)
    a3, _ = env.sample_actions(logits, seed=10)
    assert not torch.equal(a, a3)
    assert bool((mask.gather(1, a.long().unsqueeze(1)) == 1).all())
    ref = torch.log_softmax(logits + torch.clamp(torch.log(mask.float()), min=FLOAT_MIN), dim=-1)
    assert torch.allclose(logp, ref.gather(1, a.long().unsqueeze(1)).squeeze(1), atol=2e-5, rtol=0)
    p = ref.exp() * (mask != 0)
    ref_ent = -(p * torch.where(mask != 0, ref, torch.zeros_like(ref))).sum(-1)
    assert torch.allclose(ent, ref_ent, atol=2e-5, rtol=1e-5)
    # distribution: identical logits / mask in every env -> empirical frequencies = masked softmax
    row = torch.tensor([0.3 * k for k in range(26)], device=env.device)
    m1 = torch.zeros(26, dtype=torch.int8, device=env.device)
    m1[[1, 4, 5, 17, 25]] = 1
    a, _ = env.sample_actions(row.repeat(B, 1).contiguous(), mask=m1.repeat(B, 1).contiguous(), seed=3)
    freq = torch.bincount(a.long(), minlength=26).double() / B
    want = torch.softmax(row.double().masked_fill(m1 == 0, -1e30), dim=0)
    assert float(freq[m1 == 0].sum()) == 0.0
    assert float((freq - want).abs().max()) < 4.0 * float(torch.sqrt(want.max() / B)) + 1e-3


def test_ppo_trainer_writes_rollouts_in_place_and_updates_the_policy():
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.ppo import PPOTrainer
    torch.manual_seed(0)
    env = BatchedSkyjoEnv(num_envs=2048, num_players=3, seed=21, observe_other_player_indirect=True,
                          reward_refunded=0.001)                   # the reference's DEFAULT_CONFIG (skyjo_env.py:10-16)
    env.reset()
    tr = PPOTrainer(env, rollout_len=48, lr=3e-4, epochs=2, minibatches=4, ent_coef=0.01)
    st = tr.storage
    ptrs = (st.obs.data_ptr(), st.mask.data_ptr(), st.reward.data_ptr())
    before = [p.detach().clone() for p in tr.policy.parameters()]
    m = [tr.train_iteration() for _ in range(3)]
    assert ptrs == (st.obs.data_ptr(), st.mask.data_ptr(), st.reward.data_ptr())
    assert env.observations.data_ptr() == st.obs[st.T].data_ptr()          # the kernel's output IS the storage slice
    for x in m:
        assert x["illegal"] == 0 and x["env_steps"] == 48 * 2048
        assert all(np.isfinite(x[k]) for k in ("policy_loss", "vf_loss", "entropy", "kl"))
        assert x["transitions_used"] > 0.9 * x["env_steps"]
    assert m[-1]["episodes"] > 0 and 10 < m[-1]["mean_episode_len"] < 400
    assert any(not torch.equal(a, b) for a, b in zip(before, tr.policy.parameters()))
    # storage consistency: the action taken at t was legal under mask[t]; rewards only where done
    legal = st.mask[:st.T].gather(2, st.action.long().unsqueeze(2)).squeeze(2)
    assert bool((legal == 1).all())
    assert bool((st.reward[st.done == 0] == 0).all())
    env.check()
    # checkpoint / resume of learner + env: identical next iteration
    sd = tr.state_dict()
    x1 = tr.train_iteration()
    tr.load_state_dict(sd)
    x2 = tr.train_iteration()
    assert x1["episodes"] == x2["episodes"] and abs(x1["policy_loss"] - x2["policy_loss"]) < 1e-4


def test_ppo_self_play_learns_to_lower_scores():
    """Functional check of the learner (no RLlib here): a few PPO iterations of
    self-play must play visibly better than the random-admissible start (mean unpenalised score
    per seat ~59 for random 3-player games, skyjo.py:477-498)."""
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.ppo import PPOTrainer
    torch.manual_seed(0)
    env = BatchedSkyjoEnv(num_envs=8192, num_players=3, seed=1, observe_other_player_indirect=True,
                          reward_refunded=0.001)
    env.reset()
    tr = PPOTrainer(env, rollout_len=64, lr=3e-4, epochs=4, minibatches=8, ent_coef=0.01)
    hist = [tr.train_iteration() for _ in range(14)]
    first = np.mean([h["mean_raw_score"] for h in hist[1:3]])
    last = np.mean([h["mean_raw_score"] for h in hist[-2:]])
    assert sum(h["illegal"] for h in hist) == 0
    assert first > 45 and last < 0.75 * first, (first, last)
    env.check()


def test_render_table_matches_reference_strings():
    """SkyjoGame.render_table (skyjo.py:507-564) of a replayed injected game, incl. the GAME DONE block."""
    import json
    import os
    from skyjo_rl_b200 import BatchedSkyjoEnv
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "render_n3.json")))
    env = BatchedSkyjoEnv(num_envs=1, num_players=gold["num_players"], auto_reset=False)
    env.reset_injected(np.array([gold["deck"]], np.int8), np.array([gold["flips"]], np.uint8))
    for t, a in enumerate(gold["actions"]):
        if str(t) in gold["renders"]:
            assert env.game_view(0).render_table() == gold["renders"][str(t)], t
        env.step(torch.tensor([a], dtype=torch.uint8))
    assert int(env.done_code[0]) == 1
    assert env.game_view(0).render_table() == gold["renders"]["final"]


def test_sample_run_driver_plays_the_requested_games():
    from skyjo_rl_b200.sample_game import sample_run          # reference sample_game.py:5-28
    st = sample_run(games=300, config={"num_players": 2}, num_envs=256)
    assert st["episodes"] >= 300 and st["illegal"] == 0
    assert 60 < st["episode_steps"] / st["episodes"] < 95     # ~76 act() calls per 2-player game


def test_gpu_aec_replays_reference_env_traces():
    # whole episodes recorded from the unmodified SimpleSkyjoEnv under the wrapper stack of skyjo_env.env()
    # (tests/golden/make_env_trace.py), half of them ended by an illegal action, through the CUDA env
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.aec import SkyjoAECView
    from test_env_trace import load_env_traces, replay_config
    iterations = 0
    for name, c in load_env_traces():
        iterations += replay_config(lambda kw: SkyjoAECView(BatchedSkyjoEnv(num_envs=1, auto_reset=False, **kw), 0), c)
    assert iterations > 900


def test_reference_config_sweep_288_against_the_oracle():
    # reference tests/environment/test_skyjo_env_nojit.py:11-48 (288 configurations, one AEC episode each), every
    # observation and the final rewards compared with the oracle
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from test_aec_cpu import config_sweep_288
    config_sweep_288(BatchedSkyjoEnv)


def test_integration_md_stub_plays_the_same_games_as_the_shipped_mirror():
    # the reference-side ctypes stub documented in INTEGRATION.md, executed as written, against BatchedSkyjoEnv
    import ctypes as C  # noqa: F401
    from integration_stub import load_stub
    from skyjo_rl_b200 import BatchedSkyjoEnv
    m = load_stub()
    cfg = {"num_players": 3, "score_penalty": 2.0, "observe_other_player_indirect": True, "mean_reward": 1.0,
           "reward_refunded": 0.001}                                            # skyjo_env.DEFAULT_CONFIG
    B = 4096
    h, state, b = m.make(B, seed=5, **cfg)
    env = BatchedSkyjoEnv(num_envs=B, seed=5, auto_reset=True, **cfg)
    s = torch.cuda.current_stream().cuda_stream
    assert m.L.skyjo_reset(h, s) == 0
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(0)
    ended = 0
    for t in range(300):
        assert torch.equal(b["obs"], env.observations) and torch.equal(b["mask"], env.action_mask)
        assert torch.equal(b["agent"], env.agent_selection)
        a = torch.multinomial(b["mask"].float(), 1, generator=g).squeeze(1)     # int64, uniform over legal actions
        assert m.L.skyjo_step(h, a.data_ptr(), 3, s) == 0                       # 3 = SKYJO_ACT_I64
        env.step(a)
        assert torch.equal(b["done"], env.done_code) and torch.equal(b["reward"], env.rewards)
        ended += int((b["done"] != 0).sum())
    assert ended > B


def test_in_library_nccl_stats_allreduce_single_rank():
    # skyjo_stats_allreduce (the library's only collective) over a one-rank NCCL communicator created with the NCCL
    # torch loaded: the sum over one rank is the rank's own vector; the multi-rank case is tools/gpu_nccl_stats.py
    from skyjo_rl_b200 import BatchedSkyjoEnv
    from skyjo_rl_b200.nccl import StatsComm
    env = BatchedSkyjoEnv(num_envs=5000, num_players=4, seed=3)
    env.reset()
    env.step_random(300)
    comm = StatsComm(env.device)
    a = env.stats(comm=comm)
    b = env.stats()
    assert a == b and a["episodes"] > 5000 and a["steps"] > 0
    comm.close()
    env.check()
