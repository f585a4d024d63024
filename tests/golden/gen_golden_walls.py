"""Generate tests/golden/walls_*.npz by running the REAL task classes with ``walled=True``.

    python tests/golden/gen_golden_walls.py          (build container only: needs /root/reference)

No registration of the reference sets ``walled=True`` (main/envs/__init__.py:7-50 all say False), so there is no
``gym.make`` id for it: the env is built the way ``gym.make`` would -- ``TSPEnv(config=...)`` /
``ColourMatchEnv(config=...)`` from the UNMODIFIED main/envs classes with the registered config plus
``walled=True`` -- and wrapped like ``make_fixed_env`` (make_env.py:37-51).  Stored under the ids
'walled/PointTSP-v0', 'walled/ColourMatch-v0'.

What the fixtures pin (the reference's own code): ZoneEnvBase.__init__'s wall list (ZoneEnvBase.py:55-62: 244 boxes of
size 0.1 on the square of half-width 3), that the walls enter the layout before the zones (and move the numpy stream the
zones are drawn from), that observations / rewards are otherwise unchanged.  What they carry underneath, unpinned: the
oracle's Engine (oracle/sg_engine.py) and its SIMPLIFIED sphere-wall contact (oracle/mj_point.py::wall_force).
The scripted drivers ram the walls (full throttle, little steering), so most steps of each episode are in contact.
"""
import copy
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/main'
sys.path[:0] = [ROOT, os.path.join(HERE, 'stubs'), REF, os.path.join(REF, 'envs')]

import envs  # noqa: E402  (reference registrations and configs)
from envs.wrappers import FixedSeedsWrapper, ZoneWrapper  # noqa: E402
from envs.TSP_env import TSPEnv  # noqa: E402
from envs.colour_match_env import ColourMatchEnv  # noqa: E402
from tests.golden.gen_golden import snapshot_layout  # noqa: E402

CASES = [('PointTSP-v0', TSPEnv, envs.config_point, 1000000, 'ram', 500),
         ('PointTSP-v0', TSPEnv, envs.config_point, 1000001, 'corner', 700),
         ('ColourMatch-v0', ColourMatchEnv, envs.config_point_colour, 1000002, 'ram', 400)]


def wall_policy(mode, t, rs):
    """ram: full throttle straight ahead, then along the wall with slow turns; corner: throttle with a steady turn rate
    that walks the robot into a corner of the arena."""
    if mode == 'ram':
        turn = 0.0 if t < 180 else (0.25 if (t // 60) % 2 == 0 else -0.1)
        return np.array([1.0, turn], dtype=np.float32)
    return np.array([1.0 if t % 90 < 70 else -0.3, 0.12 * np.sin(t / 25.0) + 0.02 * rs.uniform(-1, 1)], dtype=np.float32)


def record(env_id, cls, config, env_seed, mode, max_len):
    cfg = copy.deepcopy(config)
    cfg['walled'] = True
    env = cls(config=cfg)
    env.seed(7)
    env = ZoneWrapper(FixedSeedsWrapper(env, min_seed=env_seed, max_seed=env_seed, rng_seed=7))
    rs = np.random.RandomState(env_seed % 1000 + 29)
    obs = env.reset()
    u = env.unwrapped
    lay = snapshot_layout(env)
    rec = {k: [] for k in ('actions', 'obs', 'zone_obs', 'reward', 'done', 'goal_met', 'qpos', 'qvel')}
    rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
    rec['qpos'].append(u.data.qpos.copy()); rec['qvel'].append(u.data.qvel.copy())
    for t in range(max_len):
        a = wall_policy(mode, t, rs)
        obs, reward, done, info = env.step(a)
        rec['actions'].append(a)
        rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
        rec['reward'].append(float(reward)); rec['done'].append(bool(done)); rec['goal_met'].append(bool(info.get('goal_met', False)))
        rec['qpos'].append(u.data.qpos.copy()); rec['qvel'].append(u.data.qvel.copy())
        if done:
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out['actions'] = out['actions'].astype(np.float32)
    out.update({'layout_' + k: v for k, v in lay.items()})
    out['env_id'] = np.array('walled/' + env_id)
    out['env_seed'] = np.int64(env_seed)
    # the reference's own wall list, as ZoneEnvBase.__init__ computed it
    out['walls_locations'] = np.array(u.walls_locations, dtype=np.float64)
    out['walls_size'] = np.float64(u.walls_size)
    out['wall_layout'] = np.array([u.layout[f'wall{i}'] for i in range(u.walls_num)], dtype=np.float64)
    return out


if __name__ == '__main__':
    for env_id, cls, config, seed, mode, max_len in CASES:
        out = record(env_id, cls, config, seed, mode, max_len)
        name = f'walls_{env_id.split("-")[0]}_{seed}_{mode}.npz'
        np.savez_compressed(os.path.join(HERE, name), **out)
        pos = out['obs'][:, 1:3] * 3.0
        print(name, 'steps', len(out['reward']), 'walls', len(out['walls_locations']), 'max |pos|', np.abs(pos).max(),
              'steps within 0.1 of a wall face', int((np.abs(pos).max(axis=1) > 2.79).sum()), 'return', out['reward'].sum())
