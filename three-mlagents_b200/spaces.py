"""Observation/action space descriptors.

The reference declares `gymnasium.spaces.Box/Discrete` in its env factories
(backend/mlagents/envs.py:38-44,169-199).  gymnasium is used when it is installed; otherwise these
two small classes provide the members the reference surface touches (`shape`, `dtype`, `n`,
`low`, `high`, `contains`, `sample`, `repr`).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium.spaces import Box, Discrete  # type: ignore
except Exception:  # noqa: BLE001

    class Box:  # type: ignore[no-redef]
        def __init__(self, low, high, shape, dtype=np.float32):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def __repr__(self) -> str:
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete:  # type: ignore[no-redef]
        def __init__(self, n: int):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x) -> bool:
            try:
                return 0 <= int(x) < self.n
            except (TypeError, ValueError):
                return False

        def sample(self) -> int:
            return int(np.random.randint(self.n))

        def __repr__(self) -> str:
            return f"Discrete({self.n})"


def spaces_for(task_id: str):
    """(observation_space, action_space) exactly as the reference declares them."""
    if task_id == "basic":        # envs.py:38-44
        return Box(0.0, 1.0, (21,), np.float32), Discrete(3)
    if task_id == "ball3d":       # envs.py:169-175
        return Box(-np.inf, np.inf, (6,), np.float32), Discrete(5)
    if task_id in ("gridworld", "push"):   # envs.py:181-199
        return Box(-1.0, 1.0, (4,), np.float32), Discrete(5)
    if task_id == "walljump":              # envs.py:202-213
        return Box(-1.0, 1.0, (4,), np.float32), Discrete(4)
    if task_id == "brickbreak":            # envs.py:216-227 (shape probed from BrickBreakEnv().reset(): 2 + 2 + 1 + 40)
        return Box(-np.inf, np.inf, (45,), np.float32), Discrete(3)
    if task_id == "bicycle":               # envs.py:230-241 (shape probed from BicycleEnv().reset(): 7 floats, bicycle.py:133-145)
        return Box(-np.inf, np.inf, (7,), np.float32), Discrete(3)
    if task_id == "glider":                # envs.py:244-255 (shape probed from GliderEnv().reset(): 9 + 3 + 3 + 1, glider.py:241-265)
        return Box(-np.inf, np.inf, (16,), np.float32), Discrete(5)
    raise KeyError(task_id)
