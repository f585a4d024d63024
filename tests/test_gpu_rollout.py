"""On-device rollout storage and crl_gae against the oracle's restatement
of base.py:195-205 and the fixture recorded from the REAL BaseAlgo.collect_experiences:
advantages and returns bit-exact (every operation is a single float32 IEEE operation in the
reference's order, so there is no tolerance to state)."""
import ctypes
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import gae as og  # noqa: E402

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'gae_base_algo.npz'))


@pytest.fixture(scope='module')
def crl():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import combinatorial_rl_tasks_b200 as m
    return m


def records(rewards, masks, last_mask):
    """The [T+1][B] result slots a rollout with these rewards / masks would hold."""
    T, B = rewards.shape
    rec = np.zeros((T + 1, B, 8), dtype=np.uint8)
    rec[1:, :, :4] = rewards.astype(np.float32).view(np.uint8).reshape(T, B, 4)
    rec[:T, :, 4] = (masks == 0)
    rec[T, :, 4] = (last_mask == 0)
    return rec


def run_gae(crl, rec, values, next_value, discount, lam, override=None):
    from combinatorial_rl_tasks_b200 import _lib
    lib = _lib.load()
    T, B = values.shape
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    r, v, nv = d(rec), d(values.astype(np.float32)), d(next_value.astype(np.float32))
    o = None if override is None else d(override.astype(np.float32))
    adv, ret = torch.zeros(T, B, device='cuda'), torch.zeros(T, B, device='cuda')
    s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.crl_gae(r.data_ptr(), None if o is None else o.data_ptr(), v.data_ptr(), nv.data_ptr(),
                           ctypes.c_double(discount), ctypes.c_double(lam), T, B, adv.data_ptr(), ret.data_ptr(), s))
    return adv.cpu().numpy(), ret.cpu().numpy()


def test_gae_kernel_reproduces_the_real_collector(crl):
    for k in range(2):
        rec = records(G[f'rewards_{k}'], G[f'masks_{k}'], G[f'last_mask_{k}'])
        adv, ret = run_gae(crl, rec, G[f'values_{k}'], G[f'next_value_{k}'], float(G['discount']), float(G['gae_lambda']))
        assert np.array_equal(adv.view(np.uint32), G[f'advantages_{k}'].view(np.uint32)), k
        assert np.array_equal(og.flatten_pt(ret), G[f'exps_returnn_{k}']), k


def test_gae_kernel_vs_oracle_random_and_override(crl):
    rs = np.random.RandomState(0)
    T, B = 37, 1003                                   # ragged: not multiples of the block or unroll sizes
    rewards = (rs.rand(T, B) < 0.1) * rs.uniform(0, 21, (T, B)).astype(np.float32)
    masks = (rs.rand(T, B) > 0.07).astype(np.float32)
    last = (rs.rand(B) > 0.07).astype(np.float32)
    values, nv = rs.normal(0, 3, (T, B)).astype(np.float32), rs.normal(0, 3, B).astype(np.float32)
    rec = records(rewards, masks, last)
    for disc, lam in [(0.99, 0.95), (0.998, 1.0), (1.0, 0.0)]:
        adv, ret = run_gae(crl, rec, values, nv, disc, lam)
        want = og.gae(rewards, values, masks, last, nv, disc, lam)
        assert np.array_equal(adv.view(np.uint32), want.view(np.uint32)), (disc, lam)
        assert np.array_equal(ret, values + want)
    shaped = rs.normal(0, 0.03, (T + 1, B)).astype(np.float32)
    adv, _ = run_gae(crl, rec, values, nv, 0.99, 0.95, override=shaped)
    assert np.array_equal(adv, og.gae(shaped[1:], values, masks, last, nv, 0.99, 0.95))


@pytest.mark.parametrize('env_id', ['PointTSP-v0', 'PointTTSP-v3'])
def test_rollout_stores_in_place_and_matches_stepwise(crl, env_id):
    """Rollout.step writes each frame straight into its slot; the slots equal what a twin env
    stepped the ordinary way returns, and finish() equals the oracle on the recorded rollout --
    over two rollouts, the second starting from the carried-over frame and mask."""
    from combinatorial_rl_tasks_b200.rollout import Rollout
    B, T = 64, 12
    env, twin = crl.ZoneVecEnv(env_id, B), crl.ZoneVecEnv(env_id, B)
    for e in (env, twin):
        e.seed(77)
        e.cfg.num_steps = 5                           # episodes end inside the rollout
    ro = Rollout(env, T, discount=0.99, gae_lambda=0.95)
    twin.reset()
    rs = np.random.RandomState(3)
    goals = env.spec.goals
    for k in range(2):
        obs = ro.begin()
        if k == 0:
            assert torch.equal(obs['obs'], twin.obs) and torch.equal(obs['zone_obs'], twin.zone_obs)
        vals = []
        for t in range(T):
            a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
            v = torch.from_numpy(rs.normal(0, 2, B).astype(np.float32)).cuda()
            vals.append(v)
            if goals:
                for e in (env, twin):
                    need = e.needs_goal()
                    e.set_goal(torch.where(need, torch.full((B,), t % 15, dtype=torch.int32, device='cuda'),
                                           torch.full((B,), -1, dtype=torch.int32, device='cuda')))
            o, r, d, info = ro.step(t, a, v, torch.zeros(B, 2, device='cuda'))
            o2, r2, d2, i2 = twin.step(a)
            assert torch.equal(o['obs'], o2['obs']) and torch.equal(o['zone_obs'], o2['zone_obs']), (k, t)
            assert torch.equal(r, r2) and torch.equal(d, d2), (k, t)
            assert o['obs'].data_ptr() == ro.obs[t + 1].data_ptr()      # written in place, not copied
            if goals:
                assert torch.equal(info['shaped_reward'], i2['shaped_reward'])
        nv = torch.from_numpy(rs.normal(0, 2, B).astype(np.float32)).cuda()
        exps = ro.finish(nv)
        rewards = (ro.shaped[1:] if goals else ro.rewards).cpu().numpy()
        masks, last = ro.masks().cpu().numpy(), 1.0 - ro.result[T, :, 4].float().cpu().numpy()
        values = torch.stack(vals).cpu().numpy()
        assert np.array_equal(ro.values.cpu().numpy(), values)
        want = og.gae(rewards, values, masks, last, nv.cpu().numpy(), 0.99, 0.95)
        assert np.array_equal(exps['advantage'].cpu().numpy().view(np.uint32), want.view(np.uint32)), k
        assert np.array_equal(exps['returnn'].cpu().numpy(), values + want)
        assert (masks == 0).any() and (k == 0 or (masks[0] == last_prev).all())
        last_prev = last
        flat = ro.finish(nv, flatten=True)
        assert np.array_equal(flat['advantage'].cpu().numpy(), og.flatten_pt(want))
        assert flat['obs']['zone_obs'].shape[0] == B * T
    assert env.counters()['episodes'] == twin.counters()['episodes'] > 0


@pytest.mark.parametrize('extra', [[], ['--fused-encoder']], ids=['torch', 'fused-encoder'])
def test_example_collect_loop_runs(crl, extra):
    """examples/collect_ppo.py end to end at a small size: rollout in place, GAE, one update; also with the
    collection forward on the tcgen05 zone encoder (weights re-packed after each update)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'examples', 'collect_ppo.py'), '--envs', '2048', '--frames', '16',
                          '--updates', '2', '--env', 'ColourMatch-v0'] + extra, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count('update ') >= 2 and 'loss' in out.stdout
