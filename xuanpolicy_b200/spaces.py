"""Minimal observation/action space descriptors.

The reference takes `gym.spaces.Box` / `Discrete` objects (xuance/common/memory_tools.py:3,
xuance/environment/gym/gym_vec_env.py:3); only `.shape`, `.n`, `.low`, `.high` and `.dtype` are read on the
PPO path (space2shape, xuance/common/common_tools.py:185-189).  gym is not a dependency of this package,
so the vec-env hands out these look-alikes; real gym/gymnasium spaces are accepted everywhere too
(duck-typed through `space_shape` / `is_discrete`).
"""
import numpy as np


class Space:
    def __init__(self, shape, dtype):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def seed(self, seed=None):
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        super().__init__(shape, dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape).copy()

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def __repr__(self):
        return "Discrete(%d)" % self.n


def is_discrete(space):
    return hasattr(space, "n") and not hasattr(space, "low")


def space_shape(space):
    """space2shape for the two space kinds on this path: () for Discrete, .shape for Box."""
    return tuple(space.shape)
