"""Minimal stand-ins for the two `gym` names the reference wrappers use (`gym.Wrapper`, `gym.spaces.Box`).

`gym` is imported by `mqe/envs/wrappers/*.py` only for attribute forwarding and for the space objects OpenRL
inspects (`openrl_ws/utils.py:43-46`).  When the real package is importable it is used; otherwise these
classes provide the same surface so the VecEnv boundary stays drop-in without the dependency.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the host
    import gym as _gym
    Wrapper = _gym.Wrapper
    Box = _gym.spaces.Box
    HAVE_GYM = True
except Exception:  # noqa: BLE001
    HAVE_GYM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=float):
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Wrapper:
        """gym.Wrapper semantics: unknown attributes are looked up on the wrapped env."""

        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(f"attempted to get missing private attribute '{name}'")
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        def step(self, action):
            return self.env.step(action)

        def close(self):
            return self.env.close() if hasattr(self.env, "close") else None


class spaces:  # `from gym import spaces` lookalike
    Box = Box
