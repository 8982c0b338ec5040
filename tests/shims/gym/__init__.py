"""Stand-in for gym 0.21 (test infrastructure, see ../README.md): only `gym.spaces`."""
from . import spaces  # noqa: F401
