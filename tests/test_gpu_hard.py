"""The hard instances PointTSP-v4 / PointTSP-v5 (main/envs/TSP_hard_env.py over the configs of
main/envs/__init__.py:52-81) on the device: fixed robot / city placements in every sampler
(CrlState.fixed_layout), distractor zones that start visited (CrlConfig.initial_visited), the
short step limits.  Fixture replays of the real TSPHardEnv (hard_*.npz, hardgoals_*.npz) and the
closed-loop runs live with the other tasks in test_gpu_parity.py / test_gpu_goals.py."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle import zone_env as ze  # noqa: E402


@pytest.fixture(scope='module')
def crl():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import combinatorial_rl_tasks_b200 as m
    return m


def keepouts_hold(origin, zxy):
    """Engine.sample_layout's acceptance test on a finished layout (fp32 positions)."""
    pts = np.concatenate([origin[None, :2], zxy]).astype(np.float64)
    keep = np.array([0.4] + [0.55] * len(zxy))
    d = np.linalg.norm(pts[:, None] - pts[None], axis=2)
    need = keep[:, None] + keep[None]
    np.fill_diagonal(d, 1e9)
    return bool(np.all(d >= need - 1e-5))


@pytest.mark.parametrize('env_id', ['PointTSP-v4', 'PointTSP-v5'])
def test_device_reset_with_fixed_placements_equals_design_twin(crl, env_id):
    """crl_reset (one-lane-per-layout sampler + copy): fixed objects sit exactly at their
    locations, the others are the design twin's draws bit for bit, the distractors start
    visited and show Yellow, v4's heading is the fixed -1 rad."""
    B = 1024
    spec = crl.ENV_SPECS[env_id]
    table = spec.fixed_layout()
    h = ze.HARD[env_id]
    n_fixed = len(h['zones_locations'])
    env = crl.ZoneVecEnv(env_id, B)
    env.seed(1000000)
    obs = env.reset()
    torch.cuda.synchronize()
    zxy, origin = env.zone_xy.cpu().numpy(), env.origin.cpu().numpy()
    bits = env.aux[:, 3].view(torch.int32).cpu().numpy()
    assert np.all((bits >> 16) & 0x7fff == spec.initial_visited) and np.all(bits & 0xffff == 0)   # bit 31 = episode parity
    assert np.all(origin[:, :2] == np.float32(h['robot_locations'][0]))
    for k in range(n_fixed):
        assert np.all(zxy[k] == np.array(h['zones_locations'][k], dtype=np.float32)), k
    if h['robot_rot'] is not None:
        assert np.all(origin[:, 2] == np.float32(h['robot_rot']))
    else:
        assert origin[:, 2].min() >= 0 and origin[:, 2].max() <= 6.2832 and origin[:, 2].std() > 1.0
    for i in list(range(0, B, 37)) + [B - 1]:
        tw = co.philox_reset(env_id, 1000000 + i, fixed=table)
        assert np.array_equal(zxy[:, i, :], tw['zone_xy']), i
        assert np.array_equal(origin[i, :2], tw['xy0']) and origin[i, 2] == np.float32(tw['rot0']), i
        assert keepouts_hold(origin[i], zxy[:, i, :]), i
    # sampled distractors really are spread over the arena, and differ between envs
    assert zxy[n_fixed:].std() > 1.0 and np.abs(zxy[n_fixed:]).max() <= 2.45 + 1e-5
    zo = obs['zone_obs'].cpu().numpy()
    cyan, yellow = np.array([0, 1, 1, 0.25], np.float32), np.array([1, 1, 0, 0.25], np.float32)
    assert np.all(zo[:, :n_fixed, 2:6] == cyan) and np.all(zo[:, n_fixed:, 2:6] == yellow)
    o = obs['obs'].cpu().numpy()
    assert np.all(o[:, 0] == 1.0) and np.allclose(o[:, 1:3] * 3, origin[:, :2], atol=1e-6)
    assert not any(env.check_state())


def test_auto_reset_of_hard_instance_prefetched_and_inline(crl):
    """PointTSP-v5 ends after 250 steps: every env auto-resets inside crl_step.  With the
    background sampler the new maps come from parked layouts, without it from the warp-cooperative
    inline sampler; both must equal the design twin, keep the fixed placements and restart with the
    distractors visited."""
    spec = crl.ENV_SPECS['PointTSP-v5']
    table = spec.fixed_layout()
    B = 512
    results = []
    for prefetch_every in (8, 0):
        env = crl.ZoneVecEnv('PointTSP-v5', B, prefetch_every=prefetch_every)
        env.seed(777000)
        env.reset()
        if not prefetch_every:
            env.next_ready.zero_()                         # drop the layouts the full reset parked ahead
        n_done = 0
        for t in range(250):
            obs, reward, done, info = env.step_random(action_seed=5)
            n_done += int(done.sum())
        torch.cuda.synchronize()
        assert n_done == B and bool(done.all())            # the 250-step limit (num_steps of the config)
        assert torch.all(env.episode == 2) and torch.all(env.steps == 0)
        bits = env.aux[:, 3].view(torch.int32).cpu().numpy()
        assert np.all((bits >> 16) & 0x7fff == spec.initial_visited)
        zxy, origin = env.zone_xy.cpu().numpy(), env.origin.cpu().numpy()
        for i in range(0, B, 41):
            tw = co.philox_reset('PointTSP-v5', 777000 + i + 1, fixed=table)     # second episode: seed + 1
            assert np.array_equal(zxy[:, i, :], tw['zone_xy']) and origin[i, 2] == np.float32(tw['rot0']), i
        c = env.counters()
        if prefetch_every:
            assert c['resets_prefetched'] == 2 * B and c['resets_inline'] == 0
        else:
            assert c['resets_prefetched'] == B and c['resets_inline'] == B
        assert not any(env.check_state())
        results.append((zxy, origin))
    assert np.array_equal(results[0][0], results[1][0]) and np.array_equal(results[0][1], results[1][1])


def test_hard_instance_success_needs_only_the_cities(crl):
    """goal_met = no zone left unvisited (TSP_env.py:71-72): with the distractors visited from the
    start, visiting the 5 cities of PointTSP-v4 ends the episode with the time bonus
    (num_steps - steps) * 0.01 of the 1000-step config."""
    env = crl.ZoneVecEnv('PointTSP-v4', 32)
    env.seed(5)
    env.reset()
    # teleport: four cities already visited, the robot parked on the fifth
    bits = env.aux[:, 3].view(torch.int32)
    bits.copy_(bits | (0b01111 << 16) | 123)
    env.pose[:, 0:2] = env.zone_xy[4]
    env.pose[:, 3] = 0
    obs, reward, done, info = env.step_no_reset(torch.zeros(32, 2, device='cuda'))
    assert bool(done.all()) and bool(info['goal_met'].all()) and torch.all(info['event'] == 1)
    assert torch.allclose(reward, torch.full_like(reward, 1 + (1000 - 123) * 0.01), atol=1e-6)
    assert torch.all(obs['obs'][:, 0] == np.float32(1 - 124 / 1000))
