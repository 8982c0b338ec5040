"""The reference's one recorded known-answer episode (notebooks/trainpettingzoo.ipynb, cell 5:
109 actions of a 3-player game through skyjo_env.env() under PettingZoo 1.14, every last() tuple
printed) replayed from tests/golden/notebook_trace.npz (made by tests/golden/make_notebook_trace.py):
  * through the oracle -- pins the C restatement against the author's own run, and
  * through the AEC view on the host-compiled kernels -- pins agent order, reward visibility and the
    dead-step phase of skyjo_rl_b200.aec against what PettingZoo printed.
The GPU twin of the second test is tests/test_gpu_policy.py::test_gpu_aec_replays_notebook_trace."""
import os

import numpy as np

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def load_trace():
    z = np.load(os.path.join(HERE, "golden", "notebook_trace.npz"))
    return {k: z[k] for k in z.files}


def replay_through_aec(aec, z):
    """Drives `aec` exactly like the notebook's loop and compares every printed quantity."""
    T = int(z["steps"])
    aec.reset_injected(z["deck"], z["flips"])
    for t, agent in enumerate(aec.agent_iter(max_iter=600)):
        obs, reward, done, info = aec.last()
        assert agent == f"player_{z['agent'][t]}" or done
        np.testing.assert_array_equal(obs["observations"], z["obs"][t], err_msg=f"obs at {t}")
        np.testing.assert_array_equal(obs["action_mask"], z["mask"][t], err_msg=f"mask at {t}")
        assert reward == z["reward"][t] and done == bool(z["done"][t]) and info == {}
        if not done:
            aec.step(int(z["action"][t]))
        else:
            assert t == T and agent == "player_0"          # the first done agent in `agents` order
            aec.step(None)
            break
    # what the notebook printed after the loop: the remaining cumulative rewards
    assert aec._cumulative_rewards == {"player_1": z["rewards"][1], "player_2": z["rewards"][2]}
    assert aec.agents == ["player_1", "player_2"] and aec.agent_selection == "player_1"


def test_oracle_replays_the_notebook_episode():
    z = load_trace()
    g = O.OracleGame(3, 2.0, False)
    g.reset_injected(z["deck"], z["flips"])
    T = int(z["steps"])
    for t in range(T):
        pid, phase = g.expected_action
        assert pid == z["agent"][t]
        obs, mask = g.collect_observation(pid)
        np.testing.assert_array_equal(obs, z["obs"][t], err_msg=f"obs at {t}")
        np.testing.assert_array_equal(mask, z["mask"][t], err_msg=f"mask at {t}")
        assert g.hand_card == z["hand"][t]
        over = g.act(pid, int(z["action"][t]))
        assert over == (t == T - 1)
    assert g.game_metrics["final_score"] == z["final_score"].tolist()        # {0: 46.0, 1: 118.0, 2: 87.0}
    assert g.final_rewards(1.0, 0.0).tobytes() == z["rewards"].tobytes()     # 38.67, -33.33, -2.33 bit for bit
    obs, mask = g.collect_observation(g.expected_action[0])
    np.testing.assert_array_equal(obs, z["obs"][T])
    np.testing.assert_array_equal(mask, z["mask"][T])


def test_hostsim_aec_view_replays_the_notebook_episode():
    from hostsim.sim import HostSimEnv
    from skyjo_rl_b200.aec import SkyjoAECView

    class Backend(HostSimEnv):
        def observation_space(self, agent):
            return None

        def action_space(self, agent):
            return None
    be = Backend(num_envs=1, num_players=3, score_penalty=2.0, mean_reward=1.0, reward_refunded=0.0,
                 auto_reset=False)
    replay_through_aec(SkyjoAECView(be, 0), load_trace())
