from . import wrappers  # noqa: F401
from .env import AECEnv  # noqa: F401
