"""AEC bookkeeping of skyjo_rl_b200.aec.SkyjoAECView (the reference's consumer loop,
rlskyjo/environment/vanilla_env_example.py:14-35) on the host-compiled backend, checked against
the oracle playing the same injected game: every live agent sees the oracle's observation, the
game-over step hands each agent its final reward exactly once with done=True in `agents` order,
and the dead-step phase empties `agents` (PettingZoo 1.14 semantics)."""
import numpy as np
import pytest

from hostsim.sim import HostSimEnv
from oracle import oracle as O
from skyjo_rl_b200.aec import RLlibDictEnv, SkyjoAECView
from skyjo_rl_b200.policy import policy_ra


class _Backend(HostSimEnv):
    def observation_space(self, agent):
        return None

    def action_space(self, agent):
        return None


@pytest.mark.parametrize("N,indirect,mr,rr", [(2, False, 1.0, 0.0), (3, True, 1.0, 0.001), (5, False, 0.0, 0.01)])
def test_aec_episode_matches_oracle(N, indirect, mr, rr):
    seed, env_id = 77, 5
    be = _Backend(num_envs=1, num_players=N, observe_other_player_indirect=indirect, mean_reward=mr,
                  reward_refunded=rr, seed=seed, auto_reset=False, first_global_env_id=env_id)
    aec = SkyjoAECView(be, 0)
    aec.reset()
    g = O.OracleGame(N, 2.0, indirect)
    g.reset_rng(seed, env_id, 0)
    rng = np.random.default_rng(3)
    seen_done, steps, over = [], 0, False
    for agent in aec.agent_iter(max_iter=300 * N):
        obs, reward, done, info = aec.last()
        if not done:
            pid = g.expected_action[0]
            assert agent == f"player_{pid}" and not over
            o, m = g.collect_observation(pid)
            np.testing.assert_array_equal(obs["observations"], o)
            np.testing.assert_array_equal(obs["action_mask"], m)
            assert reward == 0
            a = policy_ra(obs["observations"], obs["action_mask"], rng)
            over = g.act(pid, a)
            aec.step(a)
            steps += 1
        else:
            seen_done.append((agent, reward))
            aec.step(None)
    assert over and not aec.agents
    exp = g.final_rewards(mr, rr)
    assert [a for a, _ in seen_done] == [f"player_{i}" for i in range(N)]      # first done agent first
    assert np.array([r for _, r in seen_done]).tobytes() == exp.tobytes()
    with pytest.raises(Exception):
        aec.last()


def test_aec_illegal_action_terminates():
    # TerminateIllegalWrapper(illegal_reward=-1), reference skyjo_env.py:23
    be = _Backend(num_envs=1, num_players=3, seed=1, auto_reset=False)
    aec = SkyjoAECView(be, 0)
    aec.reset()
    offender = aec.agent_selection
    aec.step(3)                                  # a place action in the draw phase
    assert all(aec.dones.values())
    got = {}
    for agent in aec.agent_iter():
        _, reward, done, _ = aec.last()
        assert done
        got[agent] = reward
        with pytest.raises(ValueError):
            aec.step(5)                          # only None is valid for a done agent
        aec.step(None)
    assert got == {a: (-1.0 if a == offender else 0.0) for a in aec.possible_agents}
    with pytest.raises(AssertionError):
        SkyjoAECView(be, 0).step(24)             # OrderEnforcingWrapper: reset first


def test_rllib_dict_view_protocol():
    # the multi-agent dict protocol of RLlib's PettingZooEnv wrapper (train_model_simple_rllib.py:30-33)
    N, seed, env_id, mr, rr = 3, 5, 2, 1.0, 0.001
    be = _Backend(num_envs=1, num_players=N, observe_other_player_indirect=True, mean_reward=mr,
                  reward_refunded=rr, seed=seed, auto_reset=False, first_global_env_id=env_id)
    menv = RLlibDictEnv(SkyjoAECView(be, 0))
    g = O.OracleGame(N, 2.0, True)
    g.reset_rng(seed, env_id, 0)
    rng = np.random.default_rng(0)
    obs = menv.reset()
    for _ in range(300 * N):
        assert len(obs) == 1
        (agent, o), = obs.items()
        pid = g.expected_action[0]
        assert agent == f"player_{pid}"
        eo, em = g.collect_observation(pid)
        np.testing.assert_array_equal(o["observations"], eo)
        np.testing.assert_array_equal(o["action_mask"], em)
        a = policy_ra(o["observations"], o["action_mask"], rng)
        over = g.act(pid, a)
        obs, rew, done, info = menv.step({agent: a})
        if over:
            break
        assert done == {next(iter(obs)): False, "__all__": False}
        assert list(rew.values()) == [0]
    assert done["__all__"] and all(done[f"player_{i}"] for i in range(N))
    exp = g.final_rewards(mr, rr)
    assert np.array([rew[f"player_{i}"] for i in range(N)]).tobytes() == exp.tobytes()
    assert set(obs) == {f"player_{i}" for i in range(N)}


def config_sweep_288(make_backend):
    # the reference's own env test (tests/environment/test_skyjo_env_nojit.py:11-48) plays one simple_episode for each
    # of num_players 1..12 x score_penalty {1,2} x indirect {T,F} x mean_reward {-1,0,1} x reward_refunded {0,0.01};
    # here every one of the 288 episodes is also compared with the oracle, observation by observation
    from itertools import product
    rng = np.random.default_rng(11)
    steps = 0
    for count, (N, pen, indirect, mr, rr) in enumerate(product(range(1, 13), [1.0, 2.0], [True, False],
                                                               [-1, 0.0, 1.0], [0.0, 0.01])):
        seed, env_id = 1000 + count, count
        be = make_backend(num_envs=1, num_players=N, score_penalty=pen, observe_other_player_indirect=indirect,
                          mean_reward=mr, reward_refunded=rr, seed=seed, auto_reset=False, first_global_env_id=env_id)
        aec = SkyjoAECView(be, 0)
        aec.reset()
        g = O.OracleGame(N, pen, indirect)
        g.reset_rng(seed, env_id, 0)
        finished = []
        for agent in aec.agent_iter(max_iter=400 * N):
            obs, reward, done, info = aec.last()
            if done:
                finished.append(reward)
                aec.step(None)
                continue
            pid = g.expected_action[0]
            o, m = g.collect_observation(pid)
            assert agent == f"player_{pid}"
            np.testing.assert_array_equal(obs["observations"], o)
            np.testing.assert_array_equal(obs["action_mask"], m)
            a = policy_ra(obs["observations"], obs["action_mask"], rng)
            g.act(pid, a)
            aec.step(a)
            steps += 1
        assert g.is_terminated and not aec.agents
        assert np.array(finished).tobytes() == g.final_rewards(mr, rr).tobytes(), (N, pen, indirect, mr, rr)
    assert count == 287 and steps > 288 * 20


def test_reference_config_sweep_288_against_the_oracle():
    config_sweep_288(_Backend)
