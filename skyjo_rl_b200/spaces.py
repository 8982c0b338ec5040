"""Minimal stand-ins for the gym 0.21 spaces the reference env exposes
(rlskyjo/environment/skyjo_env.py:125-151): Box, Discrete, Dict with the same attributes.
If gym is importable its classes are used instead."""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym.spaces import Box, Dict, Discrete  # type: ignore
except Exception:  # noqa: BLE001

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete:
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def __contains__(self, k):
            return k in self.spaces

        def contains(self, x):
            return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def __repr__(self):
            return f"Dict({self.spaces})"
