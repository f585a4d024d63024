"""The oracle's own restatement of the task logic (oracle/zone_env.py) against
fixtures recorded from the REAL reference task code (tests/golden/gen_golden.py).
Everything must be identical: both sides run fp64 on the same physics."""
import glob
import os

import numpy as np
import pytest

from oracle import zone_env as ze

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
EPISODES = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.npz'))
                  if not f.endswith('_vector.npz') and not os.path.basename(f).startswith(('goals_', 'gae_', 'model_', 'hardgoals_', 'hardvec_')))
VECTORS = sorted(glob.glob(os.path.join(GOLDEN, '*_vector.npz')))


def test_fixtures_present():
    assert len(EPISODES) >= 16 and len(VECTORS) == 3
    assert sum(os.path.basename(f).startswith('hard_') for f in EPISODES) == 4     # PointTSP-v4 / v5


@pytest.mark.parametrize('path', EPISODES, ids=os.path.basename)
def test_seeded_episode_matches_reference(path):
    g = dict(np.load(path))
    env = ze.make_fixed_env(str(g['env_id']), seed=7, env_seed=int(g['env_seed']))
    obs = env.reset()
    e = env.env
    # the layout the reference built from the same seed (numpy legacy stream)
    assert np.array_equal(e.xy0, g['layout_xy0'])
    assert e.rot0 == float(g['layout_rot0'])
    assert np.array_equal(e.zone_xy, g['layout_zone_xy'])
    if 'layout_zone_max_steps' in g:
        assert np.array_equal(e.zone_max_steps, g['layout_zone_max_steps'])
    if 'layout_colours' in g:
        assert np.array_equal(e.colours, g['layout_colours'])
    assert np.array_equal(obs['obs'], g['obs'][0])
    assert np.array_equal(obs['zone_obs'], g['zone_obs'][0])
    for t, a in enumerate(g['actions']):
        obs, reward, done, info = env.step(a)
        assert np.array_equal(obs['obs'], g['obs'][t + 1]), t
        assert np.array_equal(obs['zone_obs'], g['zone_obs'][t + 1]), t
        assert reward == g['reward'][t], t
        assert done == g['done'][t], t
        assert bool(info.get('goal_met', False)) == g['goal_met'][t], t
        assert np.array_equal(e.sim.data.qpos, g['qpos'][t + 1]), t
        assert np.array_equal(e.sim.data.qvel, g['qvel'][t + 1]), t


HOST_LAYOUT_EPISODES = [f for f in EPISODES if os.path.basename(f) in (
    'PointTSP_1000001_greedy.npz', 'PointTTSP_1000002_idle.npz', 'ColourMatch_1000001_greedy.npz')]


@pytest.mark.parametrize('path', HOST_LAYOUT_EPISODES, ids=os.path.basename)
def test_host_supplied_layout_matches_reference(path):
    """Same episode, but the layout is handed in instead of sampled."""
    g = dict(np.load(path))
    env = ze.ZoneTaskEnv(ze.TASK_OF_ENV_ID[str(g['env_id'])])
    lay = {k[len('layout_'):]: g[k] for k in g if k.startswith('layout_')}
    obs = env.reset(layout=lay)
    assert np.array_equal(obs['zone_obs'], g['zone_obs'][0])
    stride = 1 if len(g['actions']) < 800 else 3   # keep the CPU suite short: check every 3rd obs
    for t, a in enumerate(g['actions']):
        obs, reward, done, info = env.step(a)
        if t % stride == 0 or done:
            assert np.array_equal(obs['obs'], g['obs'][t + 1]), t
            assert np.array_equal(obs['zone_obs'], g['zone_obs'][t + 1]), t
        assert reward == g['reward'][t] and done == g['done'][t], t


def test_hard_instance_vector_trace_matches_reference():
    """Three make_test_env('PointTSP-v5') envs (seeded once, Engine.reset increments the seed) under the vector-env
    protocol, recorded from the REAL TSPHardEnv: two auto-resets each, the distractors re-sampled around the fixed
    cities every time."""
    g = dict(np.load(os.path.join(GOLDEN, 'hardvec_PointTSP-v5.npz')))
    envs = [ze.make_task_env('PointTSP-v5') for _ in range(3)]
    for i, e in enumerate(envs):
        e.seed(1000 + 50 * i)
    vec = ze.SerialVecEnv(envs)
    obs = vec.reset()
    assert np.array_equal(np.array([o['zone_obs'] for o in obs]), g['zone_obs'][0])
    for t, acts in enumerate(g['actions']):
        obs, reward, done, info = vec.step(acts)
        assert np.array_equal(np.array([o['obs'] for o in obs]), g['obs'][t + 1]), t
        assert np.array_equal(np.array([o['zone_obs'] for o in obs]), g['zone_obs'][t + 1]), t
        assert np.array_equal(np.array(reward, dtype=np.float64), g['reward'][t]) and np.array_equal(np.array(done), g['done'][t]), t
    assert int(g['done'].sum()) == 6 and all(e._seed == 1000 + 50 * i + 3 for i, e in enumerate(envs))


@pytest.mark.parametrize('path', VECTORS, ids=os.path.basename)
def test_vector_env_autoreset_matches_reference(path):
    g = dict(np.load(path))
    env_id = str(g['env_id'])
    n = g['actions'].shape[1]
    vec = ze.SerialVecEnv([ze.make_train_env(env_id, num_training_tasks=2, rng_seed=1 + 10000 * i)
                           for i in range(n)])
    obs = vec.reset()
    assert np.array_equal(np.array([o['obs'] for o in obs]), g['obs'][0])
    assert int(g['done'].sum()) >= 1
    for t, acts in enumerate(g['actions']):
        obs, reward, done, info = vec.step(acts)
        assert np.array_equal(np.array([o['obs'] for o in obs]), g['obs'][t + 1]), t
        assert np.array_equal(np.array([o['zone_obs'] for o in obs]), g['zone_obs'][t + 1]), t
        assert np.array_equal(np.array(reward, dtype=np.float64), g['reward'][t]), t
        assert np.array_equal(np.array(done), g['done'][t]), t
        assert [bool(i.get('goal_met', False)) for i in info] == list(g['goal_met'][t]), t
