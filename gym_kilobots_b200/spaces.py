"""Minimal `Box` space used when neither gym nor gymnasium is importable (they are not in this image).

Mirrors the part of gym.spaces.Box the reference touches: low / high / shape / dtype / sample /
contains (gym_kilobots/lib/light.py:56-57, lib/kilobot.py:216-220, envs/yaml_kilobots_env.py:149-192).
"""
import numpy as np

try:  # prefer the real thing when present so downstream wrappers keep working
    from gym.spaces import Box  # noqa: F401
except Exception:  # pragma: no cover - exercised in this image
    try:
        from gymnasium.spaces import Box  # noqa: F401
    except Exception:

        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                low = np.asarray(low, dtype=dtype)
                high = np.asarray(high, dtype=dtype)
                if shape is not None:
                    low = np.broadcast_to(low, shape).copy()
                    high = np.broadcast_to(high, shape).copy()
                self.low, self.high = low, high
                self.shape = low.shape
                self.dtype = np.dtype(dtype)

            def sample(self):
                lo = np.where(np.isfinite(self.low), self.low, -1.0)
                hi = np.where(np.isfinite(self.high), self.high, 1.0)
                return np.random.uniform(lo, hi).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)
