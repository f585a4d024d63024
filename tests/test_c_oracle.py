"""The oracle's C twin (oracle/crl_oracle.c) against numpy's legacy stream, the Python
oracle and the fixtures recorded from the real reference task code."""
import glob
import os

import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import mj_point as mj
from oracle import zone_env as ze

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
EPISODES = sorted(f for f in glob.glob(os.path.join(GOLDEN, '*.npz'))
                  if not f.endswith('_vector.npz') and not os.path.basename(f).startswith(('goals_', 'gae_', 'model_', 'hard', 'walls_')))   # the C twin has no walls


def test_appendix_c_known_answers():
    """numpy-legacy draws at env.seed(1000000)."""
    e = co.CEnv('PointTTSP-v0'); e.seed(1000000); e.reset()
    assert list(e.layout()['zone_max_steps']) == [1690, 1579, 1406, 1714, 1374, 767, 1702, 694, 1639, 1482,
                                                  1882, 1751, 717, 1825, 1678]
    e = co.CEnv('ColourMatch-v0'); e.seed(1000000); e.reset()
    assert list(e.layout()['colours']) == [2, 2, 1, 0, 2, 0]
    assert e.task_state()['goal_dist'] == 5
    rs, twin = np.random.RandomState(1000001), np.random.RandomState(1000001)
    twin.random_sample()                              # binomial(10, 1.0) consumes exactly one double
    assert rs.binomial(10, 1.0) == 10 and rs.random_sample() == twin.random_sample()


@pytest.mark.parametrize('env_id', ['PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0'])
def test_legacy_stream_resets_match_numpy(env_id):
    for seed in [0, 1, 7, 99, 1000000, 1000057, 2 ** 31 + 5]:
        c = co.CEnv(env_id); c.seed(seed); oc = c.reset()
        p = ze.ZoneTaskEnv(ze.TASK_OF_ENV_ID[env_id]); p.seed(seed); op = p.reset()
        lay = c.layout()
        assert np.array_equal(lay['xy0'], p.xy0) and lay['rot0'] == p.rot0
        assert np.array_equal(lay['zone_xy'], p.zone_xy)
        if 'zone_max_steps' in lay:
            assert np.array_equal(lay['zone_max_steps'], p.zone_max_steps)
        if 'colours' in lay:
            assert np.array_equal(lay['colours'], p.colours)
        assert np.allclose(oc['obs'], op['obs'], atol=1e-15) and np.array_equal(oc['zone_obs'], op['zone_obs'])


def test_substep_matches_python_oracle():
    rs = np.random.RandomState(1)
    for _ in range(500):
        q = np.array([rs.uniform(-2, 2), rs.uniform(-2, 2), rs.uniform(-40, 40)])
        v = np.array([rs.uniform(-1.5, 1.5), rs.uniform(-1.5, 1.5), rs.uniform(-5, 5)])
        a = rs.uniform(-1.5, 1.5, 2) * (0.05 if rs.rand() < 0.3 else 1.0)
        q1, v1 = mj.substep(q, v, a)
        q2, v2 = co.substep(q, v, a)
        assert np.allclose(q1, q2, rtol=1e-12, atol=1e-13) and np.allclose(v1, v2, rtol=1e-12, atol=1e-13)


def test_closed_forms():
    """SURVEY.md A.2 sanity: terminal speed 0.3*0.05/0.01 = 1.5 m/s, straight line."""
    q, v = np.zeros(3), np.zeros(3)
    for _ in range(20000):
        q, v = co.substep(q, v, [1.0, 0.0])
    assert abs(v[0] - 1.5) < 1e-6 and abs(v[1]) < 1e-12 and abs(v[2]) < 1e-12
    assert abs(mj.MASS - 0.005188790204786391) < 1e-18 and abs(mj.COM_X - 0.01927231513576232) < 1e-16
    assert abs(mj.I_HINGE - 2.8421827485812237e-05) < 1e-19


@pytest.mark.parametrize('path', EPISODES, ids=os.path.basename)
def test_c_oracle_matches_reference_fixture(path):
    g = dict(np.load(path))
    env = co.CEnv(str(g['env_id']))
    env.seed(int(g['env_seed']))
    obs = env.reset()
    assert np.array_equal(obs['zone_obs'], g['zone_obs'][0])
    for t, a in enumerate(g['actions']):
        # every step starts from the recorded state: the C twin solves the 3x3 system with its
        # own elimination order, and the servo chatter (x3.9 per unsaturated substep) would
        # otherwise turn last-bit differences into visible ones over an episode
        env.set_state(g['qpos'][t], g['qvel'][t])
        obs, reward, done, info = env.step(a)
        assert reward == g['reward'][t] or abs(reward - g['reward'][t]) < 1e-12, t
        assert done == g['done'][t] and bool(info.get('goal_met', False)) == g['goal_met'][t], t
        assert np.array_equal(obs['zone_obs'], g['zone_obs'][t + 1]), t
        assert np.allclose(obs['obs'], g['obs'][t + 1], rtol=0, atol=2e-7), t   # robot_dir is float32 in the reference
        qp, qv = env.get_state()
        assert np.allclose(qp, g['qpos'][t + 1], rtol=0, atol=1e-9), t
        assert np.allclose(qv, g['qvel'][t + 1], rtol=0, atol=1e-8), t


def test_timed_rollout_runs():
    rate, steps, wall = co.timed_rollout('PointTSP-v0', threads=2, seconds=0.3)
    assert steps > 1000 and rate > 1e4


def test_device_gamma_and_beta_streams_have_the_reference_distribution():
    """The device draws TimedTSP's beta(3, 1.5) (TTSP_env.py:20) as Ga / (Ga + Gb) with
    gamma(k/2) built exactly from exponentials and a squared normal (design twin here, the
    kernel is compared with the twin bit for bit on the GPU).  Kolmogorov-Smirnov against
    scipy's gamma / beta CDFs, and against the Marsaglia-Tsang fallback for other shapes."""
    from scipy import stats
    n = 20000
    for shape in (3.0, 1.5, 0.5, 1.0, 2.2):
        g = np.array([co.philox_gamma(1000 + s, shape, s % 15, 0) for s in range(n)])
        assert stats.kstest(g, stats.gamma(shape).cdf).pvalue > 1e-3, shape
    tm = np.concatenate([co.philox_reset('PointTTSP-v0', 5000 + s)['zone_max_steps'] for s in range(1500)])
    ref = (np.random.RandomState(3).beta(3, 1.5, size=tm.size) * 2000).astype(np.int64)
    assert stats.ks_2samp(tm, ref).pvalue > 1e-3
    assert tm.min() >= 0 and tm.max() < 2000
