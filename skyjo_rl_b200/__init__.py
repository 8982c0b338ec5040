"""skyjo_rl_b200 -- B200-native batched SkyJo environment (drop-in for the hot path of
michaelfeil/skyjo_rl: reset/deal -> step -> action mask -> observation encoding).

    from skyjo_rl_b200 import BatchedSkyjoEnv
    env = BatchedSkyjoEnv(num_envs=1 << 20, num_players=4)
    env.reset()
    obs = env.observe()            # {"observations": int8[B,D], "action_mask": int8[B,26]} on the GPU
    env.step(actions)              # one fused CUDA launch: step + mask + observe
"""
__version__ = "0.1.0"

from .env import DEFAULT_CONFIG, BatchedSkyjoEnv, GameView  # noqa: F401
