"""`walled=True` on the GPU (CrlConfig.walled, ABI 7).  The teacher-forced replay of the walls_*
fixtures (recorded from the REAL ZoneEnvBase(walled=True)) runs with the other fixtures in test_gpu_parity.py; here:
closed loops against the oracle from identical positions while the robot rams the walls, and batch invariants.
The contact model itself is simplified and unpinned (oracle/mj_point.py::wall_force); both sides implement the same one."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import zone_env as ze  # noqa: E402
from tests.test_gpu_parity import ang_diff, batched, check_obs, rel_err, world_state  # noqa: E402

WALL_SUBSTEP_RTOL = 2e-5   # ONE substep in contact from identical states: the bar (measured 8e-6 on the host build)
# After the ten fused substeps of an env step the comparison has two regimes.  A contact is a discontinuity: the force
# switches on when the signed distance crosses zero, so when the robot grazes the face (|dist| ~ 1e-7, the fp32
# resolution of a coordinate near 2.8) the two sides can disagree about ONE substep's contact, which is worth
# h d (a_ref - a_0) ~ 1e-2 m/s -- the same kind of sensitivity as the servo chatter, and MuJoCo's own.  So:
# almost every step within WALL_STEP_RTOL, a few grazing steps within the size of one contact impulse; the states are
# re-synchronised every step, nothing accumulates.
WALL_STEP_RTOL = 2e-4
GRAZE_VEL_ATOL = 5e-2      # one or two substeps of contact force
GRAZE_POS_ATOL = 1e-3
GRAZE_FRACTION = 0.03
YAW_ATOL = 1e-2 / 3        # what check_obs allows the yaw-rate entry


def tight(a, ref):
    """(x, y, heading, vx, vy, yaw rate) within the ordinary ten-substep bars."""
    return (rel_err(a[[0, 1, 3, 4]], ref[[0, 1, 3, 4]]) <= WALL_STEP_RTOL and abs(ang_diff(a[2], ref[2])) <= WALL_STEP_RTOL
            and abs(a[5] - ref[5]) <= YAW_ATOL)


def loose(a, ref):
    """within one or two substeps' worth of contact force.  The same impulse also turns the body -- the centre of mass is
    off the hinge axis and the yaw inertia is 3e-5 kg m^2, so 1e-2 m/s of disagreement is ~1 rad/s of yaw rate, which the
    velocity servo takes back within a few substeps (measured: 0.78 rad/s on the worst grazing step of the fixtures)."""
    return (np.max(np.abs(a[:2] - ref[:2])) <= GRAZE_POS_ATOL and np.max(np.abs(a[3:5] - ref[3:5])) <= GRAZE_VEL_ATOL
            and abs(ang_diff(a[2], ref[2])) <= 5e-3 and abs(a[5] - ref[5]) <= 2.0)


@pytest.fixture(scope='module')
def crl():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import combinatorial_rl_tasks_b200 as m
    return m


def ram(mode, t, rs):
    if mode == 'ram':
        turn = 0.0 if t < 150 else (0.3 if (t // 50) % 2 == 0 else -0.15)
        return np.array([1.0, turn], dtype=np.float32)
    return np.array([1.0 if t % 80 < 65 else -0.4, 0.15 * np.sin(t / 20.0) + 0.03 * rs.uniform(-1, 1)], dtype=np.float32)


@pytest.mark.parametrize('env_id,mode,seed', [('walled/PointTSP-v0', 'ram', 3), ('walled/PointTSP-v0', 'weave', 4),
                                              ('walled/ColourMatch-v0', 'ram', 5), ('walled/PointTTSP-v0', 'weave', 6)])
def test_closed_loop_against_the_walls(crl, env_id, mode, seed):
    rs = np.random.RandomState(seed)
    task = ze.TASK_OF_ENV_ID[env_id]
    ref_env = ze.make_task_env(env_id)
    ref_env.seed(seed)
    ref_env.reset()
    lay = {'xy0': ref_env.xy0, 'rot0': ref_env.rot0, 'zone_xy': ref_env.zone_xy}
    if task == ze.TTSP:
        lay['zone_max_steps'] = ref_env.zone_max_steps
    if task == ze.CM:
        lay['colours'] = ref_env.colours
    env = crl.ZoneVecEnv(env_id, 1)
    assert env.cfg.walled == 1
    obs = env.reset(layout=batched(lay))
    zone32 = env.zone_xy[:, 0, :].cpu().numpy().astype(np.float64)
    ref_env.reset(layout=dict(lay, xy0=np.zeros(2), rot0=0.0, zone_xy=zone32))     # the kernel's frame
    in_contact, worst, grazing = 0, 0.0, 0
    for t in range(700):
        w = world_state(env)
        ref_env.set_state(w[:3], w[3:])
        a = ram(mode, t, rs)
        obs, r, d, info = env.step_no_reset(torch.from_numpy(a[None]).cuda())
        o_ref, r_ref, d_ref, i_ref = ref_env.step(a)
        res = env.result[0].cpu().numpy()
        assert bool(res[4]) == d_ref and int(res[6:7].view(np.int8)[0]) == ref_env.event, t
        w1, w_ref = world_state(env), ref_env.world_state()
        in_contact += max(abs(w_ref[0]), abs(w_ref[1])) > 2.8
        e = rel_err(w1[[0, 1, 3, 4]], w_ref[[0, 1, 3, 4]])
        worst = max(worst, e)
        if not tight(w1, w_ref):
            grazing += 1
            assert loose(w1, w_ref), (t, w1, w_ref)
        assert max(abs(w1[0]), abs(w1[1])) < 2.82, (t, w1)
        if d_ref:
            break
    print(f'{env_id} {mode}: {in_contact} steps in contact, {grazing} grazing steps, worst relative error {worst:.2e}')
    assert in_contact > 100 and grazing <= GRAZE_FRACTION * (t + 1), (grazing, t)


def test_physics_per_substep_in_contact(crl):
    """From identical states inside the contact band of a wall or a corner: ONE substep, 4096 envs (origin 0, rot0 0, so
    qpos is the world position)."""
    B, N = 4096, 15
    env = crl.ZoneVecEnv('walled/PointTSP-v0', B)
    rs = np.random.RandomState(8)
    env.reset(layout={'xy0': np.zeros((B, 2)), 'rot0': np.zeros(B), 'zone_xy': rs.uniform(-2.4, 2.4, (B, N, 2))})
    walls = dict(p0=np.zeros(2), rot0=0.0, boxes=np.array(ze.wall_locations(3), dtype=np.float64), half=0.1)
    pen = rs.uniform(-0.002, 0.012, B)
    along = np.where(np.arange(B) % 4 > 0, rs.uniform(-2.81, 2.81, B), rs.choice([-1, 1], B) * (2.8 + rs.uniform(-0.002, 0.01, B)))
    side = rs.randint(4, size=B)
    X = np.where(side == 0, 2.8 + pen, np.where(side == 1, -2.8 - pen, along))
    Y = np.where(side == 2, 2.8 + pen, np.where(side == 3, -2.8 - pen, np.where(side < 2, along, 0)))
    qpos = np.stack([X, Y, rs.uniform(-np.pi, np.pi, B)], 1)
    qvel = np.stack([rs.uniform(-1.5, 1.5, B), rs.uniform(-1.5, 1.5, B), rs.uniform(-4, 4, B)], 1)
    act = rs.uniform(-1.2, 1.2, (B, 2)).astype(np.float32)
    env.set_qpos_qvel(qpos, qvel)
    qp0, qv0 = (t.cpu().numpy() for t in env.get_qpos_qvel())
    env.physics_substeps(torch.from_numpy(act).cuda(), 1)
    qp1, qv1 = (t.cpu().numpy() for t in env.get_qpos_qvel())
    from oracle import mj_point as mj
    worst, touching = 0.0, 0
    for i in range(0, B, 2):
        touching += bool(np.any(mj.wall_force(qp0[i], qv0[i], np.zeros(3), walls) != 0))
        q_ref, v_ref = mj.substep(qp0[i], qv0[i], act[i].astype(np.float64), walls=walls)
        e = max(rel_err(qp1[i][:2], q_ref[:2]), abs(ang_diff(qp1[i][2], q_ref[2])), rel_err(qv1[i], v_ref))
        worst = max(worst, e)
        assert e <= WALL_SUBSTEP_RTOL, (i, qp0[i], qp1[i], q_ref, qv1[i], v_ref)
    print(f'worst per-substep error in contact {worst:.2e} ({touching} of {B // 2} states touching)')
    assert touching > 800


@pytest.mark.parametrize('name', ['walls_PointTSP_1000000_ram.npz', 'walls_PointTSP_1000001_corner.npz',
                                  'walls_ColourMatch_1000002_ram.npz'])
def test_wall_fixture_teacher_forced(crl, name):
    """The episodes recorded from the REAL ZoneEnvBase(walled=True) (tests/golden/gen_golden_walls.py), physics state
    forced to the recorded one before every step: task logic and observations as in test_gpu_parity, the post-physics
    state with the grazing allowance above."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', name)))
    env_id = str(g['env_id'])
    env = crl.ZoneVecEnv(env_id, 1)
    lay = {k[len('layout_'):]: np.asarray(g[k])[None] for k in g if k.startswith('layout_')}
    env.reset(layout=lay)
    T = len(g['actions'])
    acts = torch.from_numpy(g['actions']).cuda()
    qpos, qvel = torch.from_numpy(g['qpos']).cuda(), torch.from_numpy(g['qvel']).cuda()
    grazing = 0
    for t in range(T):
        env.set_qpos_qvel(qpos[t:t + 1], qvel[t:t + 1])
        o, r, d, info = env.step_no_reset(acts[t:t + 1])
        res = env.result[0].cpu().numpy()
        assert bool(res[4]) == bool(g['done'][t]) and bool(res[5]) == bool(g['goal_met'][t]), t
        assert abs(float(res[:4].view(np.float32)[0]) - g['reward'][t]) <= 1e-6, t
        qp, qv = (x.cpu().numpy()[0] for x in env.get_qpos_qvel())
        mine = np.concatenate([qp, qv])
        ref = np.concatenate([g['qpos'][t + 1], g['qvel'][t + 1]])
        if not tight(mine, ref):
            grazing += 1
            assert loose(mine, ref), (t, mine, ref)
        else:
            check_obs(ze.TASK_OF_ENV_ID[env_id], o['obs'][0].cpu().numpy(), o['zone_obs'][0].cpu().numpy(),
                      g['obs'][t + 1], g['zone_obs'][t + 1], t)
    print(f'{name}: {grazing} grazing steps of {T}')
    assert grazing <= GRAZE_FRACTION * T


def test_batch_stays_inside_the_walls_and_the_open_arena_does_not(crl):
    """4,096 envs, iid actions with a strong forward bias, auto-reset: a walled batch never has a robot beyond the inner
    faces by more than the contact's penetration; the same batch without walls leaves the square."""
    B = 4096
    far = {}
    for env_id in ('walled/PointTSP-v0', 'PointTSP-v0'):
        env = crl.ZoneVecEnv(env_id, B)
        env.seed(100)
        env.reset()
        g = torch.Generator(device='cuda').manual_seed(1)
        hold = torch.zeros(B, 2, device='cuda')
        worst = 0.0
        for t in range(1200):
            if t % 40 == 0:
                hold = torch.rand(B, 2, device='cuda', generator=g) * torch.tensor([1.0, 0.6], device='cuda') \
                    + torch.tensor([0.0, -0.3], device='cuda')
            env.step(hold)
            if t % 10 == 9:
                worst = max(worst, float(env.pose[:, :2].abs().max()))
        far[env_id] = worst
        env.check_state()
    assert far['walled/PointTSP-v0'] < 2.82, far
    assert far['PointTSP-v0'] > 3.2, far


def test_walled_flag_is_validated(crl):
    from combinatorial_rl_tasks_b200 import _lib
    env = crl.ZoneVecEnv('PointTSP-v0', 64)
    env.seed(1)
    env.reset()
    env.cfg.walled = 2
    with pytest.raises(Exception):
        env.step(torch.zeros(64, 2, device='cuda'))


@pytest.mark.parametrize('env_id', ['walled/PointTSP-v0', 'walled/PointTTSP-v0'])
def test_walled_host_step_equals_the_device_step(crl, env_id):
    """A walled env steps through the EXT kernels on every path: the host-facing call (one kernel writing into the caller's
    pinned buffers) returns byte for byte what the device step returns, while half the batch is pressed against the walls."""
    B = 512
    a_dev, a_host = crl.ZoneVecEnv(env_id, B), crl.ZoneVecEnv(env_id, B)
    for e in (a_dev, a_host):
        e.seed(77)
        e.reset()
    rs = np.random.RandomState(2)
    hold = np.zeros((B, 2), dtype=np.float32)
    touched = 0
    for t in range(400):
        if t % 50 == 0:
            hold = np.stack([rs.uniform(0.3, 1.0, B), rs.uniform(-0.2, 0.2, B)], 1).astype(np.float32)
        o_d, r_d, d_d, _ = a_dev.step(torch.from_numpy(hold).cuda())
        o_h, r_h, d_h, _ = a_host.step_host(hold)
        assert np.array_equal(o_d['obs'].cpu().numpy(), o_h['obs']), t
        assert np.array_equal(o_d['zone_obs'].cpu().numpy(), np.asarray(o_h['zone_obs'])), t
        assert np.array_equal(r_d.cpu().numpy(), r_h) and np.array_equal(d_d.cpu().numpy().astype(bool), np.asarray(d_h).astype(bool)), t
        touched = max(touched, int((a_dev.pose[:, :2].abs().max(dim=1).values > 2.8).sum()))
    assert touched > B // 8
