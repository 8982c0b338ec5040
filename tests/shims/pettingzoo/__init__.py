"""Stand-in for pettingzoo 1.14.0 (test infrastructure, see ../README.md): `AECEnv` with the helpers
`SimpleSkyjoEnv.step` calls (rlskyjo/environment/skyjo_env.py:235,250-252) and `agent_iter` / `last`
as used by rlskyjo/environment/vanilla_env_example.py:14-35."""
from .utils.env import AECEnv  # noqa: F401
