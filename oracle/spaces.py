"""CPU oracle helper: the two ``gym.spaces`` classes the hot path touches.

TEST INFRASTRUCTURE ONLY (see oracle/mj_point.py header).  ``gym`` is absent
from this image; the reference only needs ``Box`` / ``Dict`` with ``shape``,
``contains`` (asserted at ``main/envs/zone_envs/ZoneEnvBase.py:234``) and
``sample`` (used once by ``ZoneWrapper.split_zone_obs_space``,
``main/envs/wrappers.py:144-153``).
"""
from collections import OrderedDict

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), self.shape).copy()

    def contains(self, x):
        x = np.asarray(x)  # gym of the Safety Gym era checks shape and bounds only
        return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def __repr__(self):
        return f'Box{self.shape}'


class Dict:
    def __init__(self, spaces):
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(spaces.items()))
        self.spaces = OrderedDict(spaces)

    def contains(self, x):
        if not isinstance(x, dict) or len(x) != len(self.spaces):
            return False
        for k, sp in self.spaces.items():
            if k not in x or not sp.contains(x[k]):
                return False
        return True

    def sample(self):
        return OrderedDict((k, sp.sample()) for k, sp in self.spaces.items())

    def __getitem__(self, k):
        return self.spaces[k]
