"""Minimal observation/action space descriptors answering the attribute probes the
reference's callers make (``envs[0].observation_space`` / ``action_space.shape``,
train_ppo.py:127,133; ``policy_network.py:25-26`` asserts low == -1, high == 1)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def __repr__(self):
        return f'Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape}, {self.dtype})'


class Dict:
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, key):
        return self.spaces[key]

    def __repr__(self):
        return f'Dict({self.spaces})'
