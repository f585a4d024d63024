"""Stand-in for ``gym`` (absent from this image) so that the reference's own env
modules under /root/reference/main/envs import unmodified when golden vectors
are generated (tests/golden/gen_golden.py).  Only what those modules touch."""
import importlib

from oracle.sg_engine import Env  # noqa: F401
from . import spaces  # noqa: F401
from .envs import registration  # noqa: F401


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def step(self, action):
        return self.env.step(action)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def seed(self, seed=None):
        return self.env.seed(seed)


def make(env_id, **kwargs):
    spec = registration.registry[env_id]
    mod_name, cls_name = spec['entry_point'].split(':')
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kw = dict(spec['kwargs'])
    kw.update(kwargs)
    return cls(**kw)
