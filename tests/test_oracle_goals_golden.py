"""The oracle's goal-conditioned variants (PointTSP-v3 / PointTTSP-v3 / ColourMatch-v3) and
WaitWrapper against fixtures recorded from the REAL zone-goals task code
(tests/golden/gen_golden_goals.py): same seeds, same actions, same goal choices must give the
same shaped rewards, need_next_goal flags, available-goal masks, goal coordinates and
everything the base tasks already pin."""
import glob
import os

import numpy as np
import pytest

from oracle import zone_env as ze

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
EPISODES = sorted(glob.glob(os.path.join(GOLDEN, 'goals_*_*_*.npz')) +
                  glob.glob(os.path.join(GOLDEN, 'hardgoals_*.npz')))     # + zone-goals PointTSP-v4 / v5


def test_fixtures_present():
    assert len(EPISODES) == 11 and os.path.exists(os.path.join(GOLDEN, 'goals_wait_PointTSP.npz'))


@pytest.mark.parametrize('path', EPISODES, ids=os.path.basename)
def test_goal_episode_matches_reference(path):
    g = dict(np.load(path))
    env_id = str(g['env_id'])
    env = ze.make_fixed_env(env_id, seed=7, env_seed=int(g['env_seed']))
    obs = env.reset()
    u = env.env
    assert np.array_equal(u.zone_xy, g['layout_zone_xy']) and np.array_equal(u.xy0, g['layout_xy0'])
    assert np.array_equal(obs['zone_obs'], g['zone_obs'][0]) and np.array_equal(obs['obs'], g['obs'][0])
    n_set = 0
    for t, a in enumerate(g['actions']):
        if u.goal_zone is None:                    # needs_goal (penv.py:22-23)
            assert g['available'][t].any(), t
            assert np.array_equal(u.get_available_goals(), g['available'][t]), t
            u.set_goal(int(g['goal'][t]))
            n_set += 1
        else:
            assert not g['available'][t].any(), t
        assert u.goal_zone == g['goal'][t]
        assert np.array_equal(u.get_goal(), g['goal_xy'][t]), t
        obs, reward, done, info = env.step(a)
        assert reward == g['reward'][t] and done == g['done'][t], t
        assert bool(info.get('goal_met', False)) == g['goal_met'][t], t
        assert info['shaped_reward'] == g['shaped_reward'][t], t
        assert info['need_next_goal'] == g['need_next_goal'][t], t
        assert np.array_equal(obs['zone_obs'], g['zone_obs'][t + 1]), t
        assert np.array_equal(obs['obs'], g['obs'][t + 1]), t
        assert np.array_equal(u.sim.data.qpos, g['qpos'][t + 1]) and np.array_equal(u.sim.data.qvel, g['qvel'][t + 1]), t
    assert bool(done) == bool(g['done'][-1]) and n_set == int(g['available'].any(axis=1).sum())


def test_wait_wrapper_matches_reference():
    g = dict(np.load(os.path.join(GOLDEN, 'goals_wait_PointTSP.npz')))
    env = ze.make_train_env(str(g['env_id']), hier=True, num_training_tasks=3, rng_seed=11)
    rs = np.random.RandomState(3)
    obs = env.reset()
    u = env.env.env
    resets = list(g['reset_at'])
    assert np.array_equal(u.zone_xy, g['layout_0_zone_xy'])
    from tests.golden.gen_golden_goals_pick import pick_goal
    tail, layout_no = 0, 0
    for t, a in enumerate(g['actions']):
        if u.goal_zone is None and not env.inner_done:
            u.set_goal(pick_goal(obs, np.array(u.get_available_goals(), dtype=bool), rs, 'near'))
        if t <= 40:
            rs.uniform(-1, 1, 2)                    # the recorder drew the action from the same stream
        if not env.inner_done and t == 60:
            u.steps = u.num_steps - 5
        obs, reward, done, info = env.step(a)
        assert reward == g['reward'][t] and done == g['done'][t] and (len(info) == 0) == g['info_empty'][t], t
        if t in resets:
            obs = env.reset()
            layout_no += 1
            assert np.array_equal(u.zone_xy, g[f'layout_{layout_no}_zone_xy'])
        assert np.array_equal(obs['zone_obs'], g['zone_obs'][t + 1]) and np.array_equal(obs['obs'], g['obs'][t + 1]), t
    assert g['info_empty'].sum() == 5 and layout_no == 1
    assert np.all(g['zone_obs'][1:][g['info_empty']][:-1] == 0)
