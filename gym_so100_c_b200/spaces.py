"""Minimal Box / Dict spaces (gymnasium is not a dependency of this package; when it is importable
its classes are used so the VectorEnv plugs into gymnasium / SB3 tooling unchanged)."""
from __future__ import annotations

from typing import Dict as _Dict

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium.spaces import Box, Dict  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            self.shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:  # type: ignore[no-redef]
        def __init__(self, spaces: _Dict[str, object]):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x) -> bool:
            return isinstance(x, dict) and set(x) == set(self.spaces) and all(self.spaces[k].contains(x[k]) for k in x)

        def __repr__(self):
            return f"Dict({self.spaces})"


def batch_box(space: "Box", n: int) -> "Box":
    return Box(low=np.broadcast_to(space.low, (n,) + space.shape).copy(),
               high=np.broadcast_to(space.high, (n,) + space.shape).copy(), dtype=space.dtype)
