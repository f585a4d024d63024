"""Generate tests/golden/*.npz by running the REAL reference task code.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py

What is real and what is stubbed
--------------------------------
Imported unmodified from /root/reference/main: ``envs/__init__.py`` (the gym
registrations and per-task configs), ``envs/TSP_env.py``, ``TTSP_env.py``,
``colour_match_env.py``, ``zone_envs/ZoneEnvBase.py``, ``wrappers.py``
(ZoneWrapper, FixedSeedsWrapper), ``make_env.py`` and
``src/torch_ac/torch_utils/penv.py``'s worker protocol (re-enacted serially:
step, and reset on done).  Stubbed, because they are third-party packages that
are neither vendored under /root/reference nor installable here: ``gym``,
``glfw``, ``mujoco_py`` (tests/golden/stubs/) and ``safety_gym``, whose Engine is
the oracle's restatement (oracle/sg_engine.py over oracle/mj_point.py).

So the fixtures pin everything the reference itself owns -- visit events,
rewards, bonuses, timeouts, colour cycling, done / goal_met, observation layout
and seeding order -- and carry the oracle's physics and layout draws underneath
(parity unpinned for those; see oracle/mj_point.py).

Each fixture holds, for one episode driven by a recorded float32 action
sequence: the layout the reference env built, every observation, reward, done
and goal_met it returned, and the MuJoCo-frame state (qpos, qvel) after every
env step.
"""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/main'
sys.path[:0] = [ROOT, os.path.join(HERE, 'stubs'), REF, os.path.join(REF, 'envs')]

import envs  # noqa: E402,F401  (reference registrations)
from envs.make_env import make_fixed_env, make_train_env  # noqa: E402

COLOUR_INDEX = {'Blue': 0, 'Green': 1, 'Red': 2}


from tests._driver import policy  # noqa: E402


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
    rs = np.random.RandomState(env_seed % 1000 + 17)
    obs = env.reset()
    lay = snapshot_layout(env)
    rec = {k: [] for k in ('actions', 'obs', 'zone_obs', 'reward', 'done', 'goal_met', 'qpos', 'qvel')}
    rec['obs'].append(obs['obs'])
    rec['zone_obs'].append(obs['zone_obs'])
    u = env.unwrapped
    rec['qpos'].append(u.data.qpos.copy())
    rec['qvel'].append(u.data.qvel.copy())
    for t in range(max_len):
        a = policy(env_id, obs, rs, mode, t)
        obs, reward, done, info = env.step(a)
        rec['actions'].append(a)
        rec['obs'].append(obs['obs'])
        rec['zone_obs'].append(obs['zone_obs'])
        rec['reward'].append(float(reward))
        rec['done'].append(bool(done))
        rec['goal_met'].append(bool(info.get('goal_met', False)))
        assert info['cost'] == 0
        rec['qpos'].append(u.data.qpos.copy())
        rec['qvel'].append(u.data.qvel.copy())
        if done:
            break
    out = {k: np.array(v) for k, v in rec.items()}
    out['actions'] = out['actions'].astype(np.float32)
    out.update({'layout_' + k: v for k, v in lay.items()})
    out['env_id'] = np.array(env_id)
    out['env_seed'] = np.int64(env_seed)
    return out


def record_vector(env_id, n_envs, n_steps):
    """penv.py worker protocol (step; reset on done) over make_train_env envs,
    as train_ppo.py:110-112 builds them, but with 2 training maps and short
    scripted episodes so that several auto-resets happen."""
    es = [make_train_env(env_id, num_training_tasks=2, rng_seed=1 + 10000 * i) for i in range(n_envs)]
    rs = np.random.RandomState(5)
    obs = [e.reset() for e in es]
    layouts = [[snapshot_layout(e)] for e in es]
    rec = {k: [] for k in ('actions', 'obs', 'zone_obs', 'reward', 'done', 'goal_met', 'qpos', 'qvel')}
    rec['obs'].append(np.array([o['obs'] for o in obs]))
    rec['zone_obs'].append(np.array([o['zone_obs'] for o in obs]))
    for t in range(n_steps):
        acts, row = [], []
        # physics state each env steps FROM (after a reset: the new episode's), for teacher forcing
        rec['qpos'].append(np.array([e.unwrapped.data.qpos.copy() for e in es]))
        rec['qvel'].append(np.array([e.unwrapped.data.qvel.copy() for e in es]))
        for i, e in enumerate(es):
            a = policy(env_id, obs[i], rs, 'greedy', t)
            o, r, d, info = e.step(a)
            if d:
                o = e.reset()
                layouts[i].append(snapshot_layout(e))
            obs[i] = o
            acts.append(a)
            row.append((r, d, bool(info.get('goal_met', False))))
        rec['actions'].append(np.array(acts))
        rec['obs'].append(np.array([o['obs'] for o in obs]))
        rec['zone_obs'].append(np.array([o['zone_obs'] for o in obs]))
        rec['reward'].append([float(x[0]) for x in row])
        rec['done'].append([x[1] for x in row])
        rec['goal_met'].append([x[2] for x in row])
    out = {k: np.array(v) for k, v in rec.items()}
    out['actions'] = out['actions'].astype(np.float32)
    out['env_id'] = np.array(env_id)
    out['n_layouts'] = np.array([len(l) for l in layouts])
    for i, ls in enumerate(layouts):
        for j, lay in enumerate(ls):
            for k, v in lay.items():
                out[f'layout_{i}_{j}_{k}'] = v
    return out


EPISODES = [
    ('PointTSP-v0', 1000000, 'greedy', 2000),
    ('PointTSP-v0', 1000001, 'greedy', 2000),
    ('PointTSP-v0', 1000002, 'random', 2000),
    ('PointTSP-v0', 1000003, 'idle', 2000),
    ('PointTTSP-v0', 1000000, 'greedy', 2000),
    ('PointTTSP-v0', 1000001, 'deadline', 2000),
    ('PointTTSP-v0', 1000002, 'idle', 2000),
    ('PointTTSP-v0', 1000003, 'random', 2000),
    ('ColourMatch-v0', 1000000, 'greedy', 2000),
    ('ColourMatch-v0', 1000001, 'greedy', 2000),
    ('ColourMatch-v0', 1000005, 'greedy', 2000),
    ('ColourMatch-v0', 1000002, 'random', 600),
]


def main():
    for env_id, seed, mode, max_len in EPISODES:
        out = record_episode(env_id, seed, mode, max_len)
        name = f"{env_id.split('-')[0]}_{seed}_{mode}.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, 'steps', len(out['reward']), 'return', out['reward'].sum(),
              'goal_met', bool(out['goal_met'].any()), 'done', bool(out['done'][-1]))
    for env_id in ('PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0'):
        out = record_vector(env_id, n_envs=3, n_steps=1200 if env_id != 'PointTSP-v0' else 2500)
        name = f"{env_id.split('-')[0]}_vector.npz"
        np.savez_compressed(os.path.join(HERE, name), **out)
        print(name, 'resets', int(out['done'].sum()), 'return', out['reward'].sum())


if __name__ == '__main__':
    main()
