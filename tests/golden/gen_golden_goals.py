"""Generate tests/golden/goals_*.npz by running the REAL goal-conditioned task code.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden_goals.py

Imported unmodified from /root/reference/zone-goals: ``envs/__init__.py`` (registrations
PointTSP-v3 / PointTTSP-v3 / ColourMatch-v3), ``envs/TSP_next_city_env.py``,
``envs/TTSP_next_city_env.py``, ``envs/colour_match_next_city_env.py``,
``envs/zone_envs/ZoneEnvBase.py``, ``envs/wrappers.py`` (ZoneWrapper, FixedSeedsWrapper,
WaitWrapper) and ``envs/make_env.py``; the worker protocol of
``src/torch_ac/torch_utils/penv.py:4-28`` (step + reset on done, set_goal, get_goal,
needs_goal, available_goals) is re-enacted serially.  Stubbed exactly as in gen_golden.py
(gym, glfw, mujoco_py, safety_gym = the oracle's Engine restatement) plus the module
``src.utils.TSP_Solver`` (OR-tools route solver), which TSP_next_city_env.py imports a name
from and never calls (``src.utils`` would otherwise pull in the whole torch_ac tree).

A fixture is one rollout of ``hier`` high-level decisions: whenever the env needs a goal
(``goal_zone is None``) the driver reads ``get_available_goals()``, picks one (nearest
available zone, or a seeded random one), calls ``set_goal`` and records ``get_goal()``; the
low level steers towards the goal.  Recorded per step: action, goal in force, obs, reward,
done, goal_met, info['shaped_reward'], info['need_next_goal'], qpos/qvel.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/zone-goals'
sys.path[:0] = [ROOT, os.path.join(HERE, 'stubs'), REF, os.path.join(REF, 'envs')]

import types  # noqa: E402

for _name in ('src', 'src.utils', 'src.utils.TSP_Solver'):
    sys.modules[_name] = types.ModuleType(_name)
sys.modules['src.utils.TSP_Solver'].get_optim_route = None     # imported by name, never called

import envs  # noqa: E402,F401  (reference registrations)
from envs.make_env import make_fixed_env, make_train_env  # noqa: E402
from envs.wrappers import WaitWrapper  # noqa: E402

from tests._driver import steer  # noqa: E402
from tests.golden.gen_golden_goals_pick import pick_goal  # noqa: E402

COLOUR_INDEX = {'Blue': 0, 'Green': 1, 'Red': 2}


def snapshot_layout(env):
    u = env.unwrapped
    N = u.zones_num
    lay = {
        'xy0': np.array(u.layout['robot'], dtype=np.float64),
        'rot0': np.float64(u.world_config_dict['robot_rot']),
        'zone_xy': np.array([u.layout[f'zone{i}'] for i in range(N)], dtype=np.float64),
    }
    if hasattr(u, 'zone_max_steps'):
        lay['zone_max_steps'] = np.array(u.zone_max_steps, dtype=np.int64)
    if hasattr(u, 'zone_cooldowns'):
        lay['colours'] = np.array([COLOUR_INDEX[z.name] for z in u.zones], dtype=np.int64)
    return lay


def record_episode(env_id, env_seed, mode, max_len=2000):
    env = make_fixed_env(env_id, seed=7, env_seed=env_seed)
    u = env.unwrapped
    rs = np.random.RandomState(env_seed % 1000 + 23)
    obs = env.reset()
    lay = snapshot_layout(env)
    keys = ('actions', 'goal', 'goal_xy', 'available', 'obs', 'zone_obs', 'reward', 'done', 'goal_met', 'shaped_reward',
            'need_next_goal', 'qpos', 'qvel')
    rec = {k: [] for k in keys}
    rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
    rec['qpos'].append(u.data.qpos.copy()); rec['qvel'].append(u.data.qvel.copy())
    N = u.num_cities
    for t in range(max_len):
        avail = np.zeros(N, dtype=bool)
        if u.goal_zone is None:                     # penv.py "needs_goal"
            avail = np.array(u.get_available_goals(), dtype=bool)
            u.set_goal(pick_goal(obs, avail, rs, mode))
        g = int(u.goal_zone)
        gxy = np.array(u.get_goal(), dtype=np.float64)
        a = steer(obs, gxy * 3.0, rs, 0.1) if mode != 'noise' else rs.uniform(-1.5, 1.5, 2).astype(np.float32)
        obs, reward, done, info = env.step(a)
        rec['actions'].append(a); rec['goal'].append(g); rec['goal_xy'].append(gxy); rec['available'].append(avail)
        rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
        rec['reward'].append(float(reward)); rec['done'].append(bool(done))
        rec['goal_met'].append(bool(info.get('goal_met', False)))
        rec['shaped_reward'].append(float(info['shaped_reward']))
        rec['need_next_goal'].append(bool(info['need_next_goal']))
        rec['qpos'].append(u.data.qpos.copy()); rec['qvel'].append(u.data.qvel.copy())
        if done:
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out['actions'] = out['actions'].astype(np.float32)
    out.update({'layout_' + k: v for k, v in lay.items()})
    out['env_id'] = np.array(env_id)
    out['env_seed'] = np.int64(env_seed)
    return out


def record_wait(env_id, n_steps):
    """WaitWrapper(ZoneWrapper(FixedSeedsWrapper(env))) driven past the end of its episode with
    step (no reset): the no-op tail (zeros, reward 0, done, empty info), then reset."""
    env = make_train_env(env_id, hier=True, num_training_tasks=3, rng_seed=11)
    assert isinstance(env, WaitWrapper)
    u = env.unwrapped
    rs = np.random.RandomState(3)
    rec = {k: [] for k in ('actions', 'obs', 'zone_obs', 'reward', 'done', 'info_empty', 'reset_at')}
    obs = env.reset()
    lays = [snapshot_layout(env)]
    rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
    tail = 0
    for t in range(n_steps):
        if u.goal_zone is None and not env.inner_done:
            u.set_goal(pick_goal(obs, np.array(u.get_available_goals(), dtype=bool), rs, 'near'))
        a = np.zeros(2, dtype=np.float32) if t > 40 else rs.uniform(-1, 1, 2).astype(np.float32)
        if not env.inner_done:
            # shorten the episode: jump the step counter close to the end once
            if t == 60:
                u.steps = u.num_steps - 5
        obs, reward, done, info = env.step(a)
        rec['actions'].append(a)
        rec['obs'].append(obs['obs']); rec['zone_obs'].append(obs['zone_obs'])
        rec['reward'].append(float(reward)); rec['done'].append(bool(done)); rec['info_empty'].append(len(info) == 0)
        if env.inner_done:
            tail += 1
        if tail == 6:                               # reset after a few no-op steps
            obs = env.reset()
            lays.append(snapshot_layout(env))
            rec['reset_at'].append(t)
            rec['obs'][-1] = obs['obs']; rec['zone_obs'][-1] = obs['zone_obs']
            tail = 0
    out = {k: np.array(v) for k, v in rec.items()}
    out['actions'] = out['actions'].astype(np.float32)
    out['env_id'] = np.array(env_id)
    out['n_layouts'] = np.int64(len(lays))
    for j, lay in enumerate(lays):
        for k, v in lay.items():
            out[f'layout_{j}_{k}'] = v
    return out


EPISODES = [
    ('PointTSP-v3', 1000000, 'near', 2000),
    ('PointTSP-v3', 1000001, 'far', 2000),
    ('PointTSP-v3', 1000002, 'random', 1200),
    ('PointTTSP-v3', 1000000, 'near', 2000),
    ('PointTTSP-v3', 1000001, 'far', 2000),
    ('PointTTSP-v3', 1000003, 'noise', 2000),
    ('ColourMatch-v3', 1000000, 'near', 1500),
    ('ColourMatch-v3', 1000001, 'far', 1500),
    ('ColourMatch-v3', 1000002, 'random', 1500),
]


def main():
    for env_id, seed, mode, max_len in EPISODES:
        out = record_episode(env_id, seed, mode, max_len)
        name = f"goals_{env_id.split('-')[0]}_{seed}_{mode}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, 'steps', len(out['reward']), 'return', out['reward'].sum(), 'shaped', out['shaped_reward'].sum(),
              'goals set', int(out['available'].any(axis=1).sum()), 'reached', int(out['need_next_goal'].sum()),
              'goal_met', bool(out['goal_met'].any()))
    out = record_wait('PointTSP-v3', 160)
    np.savez_compressed(os.path.join(HERE, 'goals_wait_PointTSP.npz'), **out)
    print('goals_wait_PointTSP.npz', 'noop steps', int(out['info_empty'].sum()), 'resets', len(out['reset_at']))


if __name__ == '__main__':
    main()
