"""Goal-conditioned variants (PointTSP-v3 / PointTTSP-v3 / ColourMatch-v3) and WaitWrapper
semantics of the CUDA path, through the C ABI, against (a) the fixtures recorded from the REAL
zone-goals task code (tests/golden/gen_golden_goals.py) and (b) the oracle stepped from
identical positions.  Bars: need_next_goal, available-goal masks, done, goal_met, integer
reward components bit-exact; rewards and shaped rewards 1e-6 from identical positions (2e-6
against the fixtures, whose fp64 positions the kernel holds rounded to fp32: 2.4e-7 per
coordinate at |x| < 4 enters both distances of the difference)."""
import glob
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from oracle import zone_env as ze  # noqa: E402
from tests._driver import steer  # noqa: E402
from tests.golden.gen_golden_goals_pick import pick_goal  # noqa: E402
from tests.test_gpu_parity import REWARD_ATOL, batched, check_obs, load_layout, world_state  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
EPISODES = sorted(glob.glob(os.path.join(GOLDEN, 'goals_*_*_*.npz')) +
                  glob.glob(os.path.join(GOLDEN, 'hardgoals_*.npz')))     # + zone-goals PointTSP-v4 / v5
SHAPED_FIXTURE_ATOL = 2e-6


@pytest.fixture(scope='module')
def crl():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    import combinatorial_rl_tasks_b200 as m
    return m


@pytest.mark.parametrize('path', EPISODES, ids=os.path.basename)
def test_goal_fixture_teacher_forced(crl, path):
    """What the real TSPNextCityEnv / TimedTSPNextCityEnv / ColourMatchNextCityEnv returned,
    replayed on the GPU with the physics state forced to the recorded qpos/qvel before every
    step and the recorded goal choices made through the batched RPCs."""
    g = dict(np.load(path))
    env_id = str(g['env_id'])
    task = ze.TASK_OF_ENV_ID[env_id]
    env = crl.ZoneVecEnv(env_id, 1)
    obs = env.reset(layout=batched(load_layout(g)))
    check_obs(task, obs['obs'][0].cpu().numpy(), obs['zone_obs'][0].cpu().numpy(), g['obs'][0], g['zone_obs'][0], 'reset')
    assert env.envs[0].goal_zone is None and env.envs[0].goal_dim == 2
    T = len(g['actions'])
    acts = torch.from_numpy(g['actions']).cuda()
    qpos, qvel = torch.from_numpy(g['qpos']).cuda(), torch.from_numpy(g['qvel']).cuda()
    n_set = 0
    rec = {k: [] for k in ('res', 'shaped', 'obs', 'zobs')}
    for t in range(T):
        env.set_qpos_qvel(qpos[t:t + 1], qvel[t:t + 1])
        need = bool(env.needs_goal()[0].item())
        assert need == bool(g['available'][t].any()), (t, 'needs_goal')
        if need:
            assert np.array_equal(env.available_goals()[0].cpu().numpy(), g['available'][t]), (t, 'available')
            env.set_goal(torch.tensor([int(g['goal'][t])]))
            n_set += 1
        if need or t % 50 == 0:
            assert int(env.goal[0].item()) == int(g['goal'][t]), (t, 'goal_zone')
            assert np.max(np.abs(env.get_goal()[0].cpu().numpy() - g['goal_xy'][t])) <= 2e-6, (t, 'get_goal')
        o, r, d, info = env.step_no_reset(acts[t:t + 1])
        rec['res'].append(env.result.clone()); rec['shaped'].append(info['shaped_reward'].clone())
        rec['obs'].append(o['obs'].clone()); rec['zobs'].append(o['zone_obs'].clone())
    res = torch.cat(rec['res']).cpu().numpy()
    shaped = torch.cat(rec['shaped']).cpu().numpy().astype(np.float64)
    assert np.array_equal(res[:, 4].astype(bool), g['done']), 'done'
    assert np.array_equal(res[:, 5].astype(bool), g['goal_met']), 'goal_met'
    assert np.array_equal(res[:, 7].astype(bool), g['need_next_goal']), 'need_next_goal'
    assert np.max(np.abs(res.view(np.float32)[:, 0].astype(np.float64) - g['reward'])) <= REWARD_ATOL
    assert np.max(np.abs(shaped - g['shaped_reward'])) <= SHAPED_FIXTURE_ATOL, np.max(np.abs(shaped - g['shaped_reward']))
    assert n_set == int(g['available'].any(axis=1).sum()) and n_set >= 1
    obs_g, zobs_g = torch.cat(rec['obs']).cpu().numpy(), torch.cat(rec['zobs']).cpu().numpy()
    for t in range(0, T, 7):
        check_obs(task, obs_g[t], zobs_g[t], g['obs'][t + 1], g['zone_obs'][t + 1], t)
    assert (int(env.goal[0].item()) == -1) == bool(g['need_next_goal'][-1])   # episode over: goal_zone is None
    assert env.counters()['goals_rejected'] == 0


@pytest.mark.parametrize('env_id,mode,seed', [('PointTSP-v3', 'near', 21), ('PointTSP-v3', 'random', 22),
                                              ('PointTTSP-v3', 'near', 23), ('ColourMatch-v3', 'far', 24),
                                              ('ColourMatch-v3', 'random', 25)])
def test_goal_closed_loop_identical_positions(crl, env_id, mode, seed):
    """Oracle (goal variant) and CUDA path side by side from IDENTICAL positions, goals chosen by
    a scripted high level through the batched RPCs, the low level steering at the goal."""
    task = ze.TASK_OF_ENV_ID[env_id]
    rs = np.random.RandomState(seed)
    ref = ze.ZoneTaskEnv(task, goals=True)
    ref.seed(seed)
    ref.reset()
    lay = {'xy0': ref.xy0, 'rot0': ref.rot0, 'zone_xy': ref.zone_xy}
    if task == ze.TTSP:
        lay['zone_max_steps'] = ref.zone_max_steps
    if task == ze.CM:
        lay['colours'] = ref.colours
    env = crl.ZoneVecEnv(env_id, 1)
    obs = env.reset(layout=batched(lay))
    zone32 = env.zone_xy[:, 0, :].cpu().numpy().astype(np.float64)
    ref.reset(layout=dict(lay, xy0=np.zeros(2), rot0=0.0, zone_xy=zone32))
    n_goals = n_wrong = 0
    for t in range(2000):
        w = world_state(env)
        ref.set_state(w[:3], w[3:])
        o_np = {'obs': obs['obs'][0].cpu().numpy(), 'zone_obs': obs['zone_obs'][0].cpu().numpy()}
        assert bool(env.needs_goal()[0].item()) == (ref.goal_zone is None), t
        if ref.goal_zone is None:
            avail = env.available_goals()[0].cpu().numpy()
            assert np.array_equal(avail, ref.get_available_goals()), (t, 'available')
            goal = pick_goal(o_np, avail, rs, mode)
            env.set_goal(torch.tensor([goal]))
            ref.set_goal(goal)
            assert np.max(np.abs(env.get_goal()[0].cpu().numpy() - ref.get_goal())) <= 2e-6
            n_goals += 1
        else:
            ref.last_dist_to_goal = ref._dist_to_goal()       # identical positions: the kernel's own state
        a = steer(o_np, ref.zone_xy[ref.goal_zone], rs, 0.3)
        obs, r, d, info = env.step_no_reset(torch.from_numpy(a[None]).cuda())
        o_ref, r_ref, d_ref, i_ref = ref.step(a)
        res = env.result[0].cpu().numpy()
        assert bool(res[4]) == d_ref and bool(res[5]) == bool(i_ref.get('goal_met', False)), t
        assert int(res[6:7].view(np.int8)[0]) == ref.event, (t, 'event')
        assert bool(res[7]) == i_ref['need_next_goal'], (t, 'need_next_goal')
        assert abs(float(res[:4].view(np.float32)[0]) - r_ref) <= REWARD_ATOL
        sh = float(info['shaped_reward'][0].item())
        assert abs(sh - i_ref['shaped_reward']) <= REWARD_ATOL, (t, sh, i_ref['shaped_reward'])
        n_wrong += i_ref['shaped_reward'] < -0.5
        check_obs(task, obs['obs'][0].cpu().numpy(), obs['zone_obs'][0].cpu().numpy(), o_ref['obs'], o_ref['zone_obs'], t)
        if d_ref:
            break
    assert n_goals >= 2
    assert (int(env.goal[0].item()) == -1) == bool(res[7])
    print(env_id, mode, 'steps', t + 1, 'goals set', n_goals, 'wrong-zone penalties', n_wrong)


def test_goal_rpcs_batched_and_rejections(crl):
    """set_goal on a visited zone or out of range is the reference's AssertionError: the goal
    stays unset and the request is counted; negative entries leave envs alone; auto-reset and
    reset clear the goal (done => goal_zone = None, TSP_next_city_env.py:69-72)."""
    B = 96
    env = crl.ZoneVecEnv('PointTSP-v3', B)
    env.seed(5)
    env.reset()
    assert bool(env.needs_goal().all().item()) and bool(env.available_goals().all().item())
    g = torch.full((B,), -1, dtype=torch.int32)
    g[:32] = torch.arange(32) % 15
    g[40] = 15                                               # out of range
    env.set_goal(g)
    goals = env.goal.cpu().numpy()
    assert np.array_equal(goals[:32], np.arange(32) % 15) and np.all(goals[32:] == -1)
    assert env.counters()['goals_rejected'] == 1
    assert not env.needs_goal()[:32].any() and env.needs_goal()[32:].all()
    xy = env.get_goal().cpu().numpy()
    want = env.zone_xy.cpu().numpy()[np.arange(32) % 15, np.arange(32)] / 3.0
    assert np.max(np.abs(xy[:32] - want)) <= 2e-6 and np.all(xy[32:] == 0.0)
    # penv.py:75-80: one env at a time
    env.set_goal_at(50, 3)
    assert int(env.goal[50].item()) == 3 and env.envs[0].goal_zone == 0
    # visit zone 2 in env 60, then ask for it
    env.pose[60, :2] = env.zone_xy[2, 60, :]
    env.set_goal_at(60, 7)
    env.step_no_reset(torch.zeros(B, 2, device='cuda'))
    assert not bool(env.available_goals()[60, 2].item()) and int(env.available_goals()[60].sum().item()) == 14
    assert int(env.goal[60].item()) == 7 and not bool(env.need_next_goal[60].item())
    assert bool(env.need_next_goal[70].item())              # no goal set: keeps asking
    env.pose[60, :2] = env.zone_xy[7, 60, :]                 # now reach the goal zone itself
    env.step_no_reset(torch.zeros(B, 2, device='cuda'))
    assert bool(env.need_next_goal[60].item()) and float(env.shaped_reward[60].item()) == 0.0
    assert int(env.goal[60].item()) == -1
    env.set_goal_at(60, 2)                                      # visited: rejected
    assert int(env.goal[60].item()) == -1 and env.counters()['goals_rejected'] == 2
    # the step limit ends the episode: need_next_goal, goal cleared, auto-reset clears too
    twin = crl.ZoneVecEnv('PointTSP-v3', B)
    twin.seed(5)
    twin.reset()
    rs = np.random.RandomState(2)
    for e in (env, twin):
        e.reset(mask=torch.ones(B, dtype=torch.uint8))       # both at episode 2 of the same seeds
    assert torch.equal(env.zone_xy, twin.zone_xy)
    for t in range(30):                                       # get the robots moving
        a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
        for e in (env, twin):
            if t == 0:
                e.set_goal(torch.full((B,), 1, dtype=torch.int32))
            e.step_no_reset(a)
    for e in (env, twin):
        e.aux[:, 3] = (e.aux[:, 3].view(torch.int32) & ~0xffff | 1999).view(torch.float32)
    a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
    o, r, d, info = env.step(a)                               # auto-reset inside the call
    o2, r2, d2, info2 = twin.step_no_reset(a)
    assert bool(d.all().item()) and bool(info['need_next_goal'].all().item())
    assert bool((env.goal == -1).all().item()) and bool((env.steps == 0).all().item())
    # the final step's shaped reward is measured at the OLD episode's post-physics position
    assert torch.equal(info['shaped_reward'], info2['shaped_reward']) and torch.equal(r, r2)
    assert float(info['shaped_reward'].abs().max().item()) > 1e-4
    # ColourMatch: a zone other than the goal changes colour -> shaped reward - 1
    # (colour_match_next_city_env.py:125-127); the goal zone itself -> exactly 0
    cm = crl.ZoneVecEnv('ColourMatch-v3', 64)
    cm.seed(9)
    cm.reset()
    assert bool(cm.available_goals().all().item())          # every zone is always eligible
    cm.set_goal(torch.full((64,), 1, dtype=torch.int32))
    cm.pose[:32, :2] = cm.zone_xy[0, :32, :]                 # wrong zone
    cm.pose[32:, :2] = cm.zone_xy[1, 32:, :]                 # the goal zone
    o, r, d, info = cm.step_no_reset(torch.zeros(64, 2, device='cuda'))
    sh = info['shaped_reward'].cpu().numpy()
    assert np.all(np.abs(sh[:32] + 1.0) < 1e-6) and np.all(sh[32:] == 0.0)
    nn = info['need_next_goal'].cpu().numpy()
    assert not nn[:32].any() and (nn[32:] | d.cpu().numpy()[32:]).all()


def test_wait_wrapper_semantics(crl):
    """wrappers.py:29-54: under step_no_reset an env that finished is a no-op returning zeros,
    reward 0, done True, until reset; compared with the oracle's Wait from identical positions."""
    B = 8
    refs = []
    lays = {'xy0': [], 'rot0': [], 'zone_xy': []}
    for i in range(B):
        e = ze.ZoneTaskEnv(ze.TSP)
        e.seed(100 + i)
        e.reset()
        for k in lays:
            lays[k].append(getattr(e, k))
        refs.append(ze.Wait(e))
    env = crl.ZoneVecEnv('PointTSP-v0', B, wait=True)
    obs = env.reset(layout={k: np.array(v) for k, v in lays.items()})
    zone32 = env.zone_xy.cpu().numpy().astype(np.float64)
    for i, r in enumerate(refs):
        r.env.reset(layout={'xy0': np.zeros(2), 'rot0': 0.0, 'zone_xy': zone32[:, i, :]})
        r.inner_done = False
        r.env.steps = 2000 - 3 - i                            # finish at staggered times
    steps = torch.tensor([2000 - 3 - i for i in range(B)], dtype=torch.int32, device='cuda')
    env.aux[:, 3] = (env.aux[:, 3].view(torch.int32) & ~0xffff | steps).view(torch.float32)
    rs = np.random.RandomState(1)
    for t in range(14):
        a = rs.uniform(-1, 1, (B, 2)).astype(np.float32)
        pose, aux = env.pose.cpu().numpy().astype(np.float64), env.aux.cpu().numpy().astype(np.float64)
        for i, r in enumerate(refs):
            if not r.inner_done:
                r.env.set_state([pose[i, 0], pose[i, 1], pose[i, 2]], [pose[i, 3], aux[i, 0], aux[i, 1]])
        o, rew, d, info = env.step_no_reset(torch.from_numpy(a).cuda())
        og, zg = o['obs'].cpu().numpy(), o['zone_obs'].cpu().numpy()
        for i, r in enumerate(refs):
            o_ref, r_ref, d_ref, i_ref = r.step(a[i])
            assert bool(d[i].item()) == d_ref and float(rew[i].item()) == r_ref, (t, i)
            if len(i_ref) == 0:                               # parked: WaitWrapper's no-op
                assert not og[i].any() and not zg[i].any(), (t, i)
                assert not bool(env.goal_met[i].item()) and int(env.event[i].item()) == 0
            else:
                check_obs(ze.TSP, og[i], zg[i], o_ref['obs'], o_ref['zone_obs'], (t, i))
    assert all(r.inner_done for r in refs)
    c = env.counters()
    assert c['episodes'] == B                                 # each episode counted once, not once per parked step
    assert bool((env.steps == 0xffff).all().item())
    # a masked reset wakes the envs it rebuilds; the others stay parked
    mask = torch.zeros(B, dtype=torch.uint8); mask[:4] = 1
    env.reset(mask=mask)
    o, rew, d, info = env.step_no_reset(torch.zeros(B, 2, device='cuda'))
    assert not d[:4].any() and bool(d[4:].all().item())
    assert bool((o['obs'][:4, 0] == np.float32(1999 / 2000)).all().item()) and not o['obs'][4:].any()
    # step() (auto-reset) never parks
    env2 = crl.ZoneVecEnv('PointTSP-v0', B, wait=True)
    env2.reset()
    env2.aux[:, 3] = (env2.aux[:, 3].view(torch.int32) & ~0xffff | 1999).view(torch.float32)
    o, rew, d, info = env2.step(torch.zeros(B, 2, device='cuda'))
    assert bool(d.all().item()) and bool((env2.steps == 0).all().item())


def test_parked_env_under_auto_reset_step_is_noop_then_reset(crl):
    """hier_base.py:180-183: step_no_reset for skill_len - 1 steps, then step().  An env parked
    in between goes through WaitWrapper's no-op (wrappers.py:36-44: reward 0, done, info {}) and the
    worker's `if done: obs = env.reset()` (penv.py:7-10): the call returns the NEW episode's first
    observation, reward 0, done True, and counts no second episode."""
    B = 8
    kw = dict(seed_mode='fixed_range', min_seed=1, max_seed=50, wait=True)
    for env_id in ('PointTSP-v0', 'PointTSP-v3', 'ColourMatch-v0'):
        env, twin = crl.ZoneVecEnv(env_id, B, **kw), crl.ZoneVecEnv(env_id, B, **kw)
        for v in (env, twin):
            v.reset()
            # envs 0-2 run into the step limit after 1 / 2 / 3 steps
            st = torch.tensor([1999, 1998, 1997], dtype=torch.int32, device='cuda')
            v.aux[:3, 3] = (v.aux[:3, 3].view(torch.int32) & ~0xffff | st).view(torch.float32)
            if v.spec.task != 2:
                # env 3 succeeds at its first step: every zone but the last visited, robot on the last one
                bits = v.aux[3, 3].view(torch.int32)
                v.aux[3, 3] = (bits & 0xffff | (((1 << 14) - 1) << 16)).view(torch.float32)
                v.pose[3, :2] = v.zone_xy[14, 3, :]
        n_done = 3 if env.spec.task == 2 else 4
        rs = np.random.RandomState(3)
        for t in range(4):
            a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
            o, r, d, info = env.step_no_reset(a)
            twin.step_no_reset(a)
        assert bool(d[:n_done].all().item()) and not bool(d[n_done:].any().item())
        assert env.counters()['episodes'] == n_done
        before = env.counters()
        # what a reset of exactly the parked envs builds (same seeds, same episode numbers)
        mask = torch.zeros(B, dtype=torch.uint8); mask[:n_done] = 1
        first = twin.reset(mask=mask)
        first_obs, first_zone = first['obs'].clone(), first['zone_obs'].clone()
        a = torch.from_numpy(rs.uniform(-1, 1, (B, 2)).astype(np.float32)).cuda()
        o, r, d, info = env.step(a)                           # the auto-reset step
        o2, r2, d2, info2 = twin.step(a)
        assert bool(d[:n_done].all().item()) and not bool(d[n_done:].any().item())
        assert not r[:n_done].any() and not env.goal_met[:n_done].any() and not env.event[:n_done].any()
        assert torch.equal(o['obs'][:n_done], first_obs[:n_done]) and torch.equal(o['zone_obs'][:n_done], first_zone[:n_done])
        assert bool((o['obs'][:n_done, 0] == 1.0).all().item()) and bool(o['zone_obs'][:n_done].any().item())
        assert bool((env.steps[:n_done] == 0).all().item())
        # the envs that were never parked stepped as usual, identically in both replicas
        assert torch.equal(o['obs'][n_done:], o2['obs'][n_done:]) and torch.equal(r[n_done:], r2[n_done:])
        after = env.counters()
        for k in ('episodes', 'successes', 'return_sum', 'length_sum'):
            assert after[k] == before[k], k                   # nothing counted twice
        if env.spec.goals:
            assert not info['shaped_reward'][:n_done].any() and not info['need_next_goal'][:n_done].any()
        # and the revived envs now run: the next step is an ordinary one
        o, r, d, info = env.step(a)
        assert not bool(d.any().item()) and bool((env.steps[:n_done] == 1).all().item())
    # the reference's list / tuple protocol: info {} for the parked envs
    pe = crl.ParallelEnv('PointTSP-v0', 4, num_training_tasks=5, hier=True)
    pe.reset()
    pe.vec.aux[:2, 3] = (pe.vec.aux[:2, 3].view(torch.int32) & ~0xffff | 1999).view(torch.float32)
    acts = [np.zeros(2)] * 4
    pe.step_no_reset(acts)
    o, r, d, info = pe.step_no_reset(acts)
    assert info[0] == {} and info[1] == {} and d[0] and not o[0]['obs'].any()
    o, r, d, info = pe.step(acts)
    assert info[0] == {} and info[1] == {} and d[0] and d[1] and r[0] == 0.0 and not d[2]
    assert o[0]['obs'][0] == 1.0 and o[0]['zone_obs'].any() and info[2] == {'cost': 0}
    o, r, d, info = pe.step(acts)
    assert not any(d) and info[0] == {'cost': 0} and o[0]['obs'][0] == np.float32(1999 / 2000)


def test_parallel_env_compat_has_the_reference_protocol(crl):
    """compat.ParallelEnv: the call and return types of penv.py:46-66 and zone-goals
    penv.py:75-99 (lists / tuples of per-env Python objects), values equal to the tensor API's."""
    B = 6
    pe = crl.ParallelEnv('PointTSP-v3', B, num_training_tasks=5, hier=True)
    tw = crl.ZoneVecEnv('PointTSP-v3', B, seed_mode='fixed_range', min_seed=1, max_seed=5, wait=True)
    obs = pe.reset()
    tw.reset()
    assert isinstance(obs, list) and len(obs) == B and obs[0]['zone_obs'].shape == (15, 6) and obs[0]['obs'].shape == (8,)
    assert pe.envs[0].observation_space is pe.observation_space and pe.action_space.shape == (2,)
    assert pe.envs[0].unwrapped.num_cities == 15 and pe.envs[0].unwrapped.goal_dim == 2 and pe.envs[0].goal_zone is None
    assert pe.needs_goal() == [True] * B
    av = pe.available_goals(2)
    assert av.dtype == bool and av.shape == (15,) and av.all()
    for i in range(B):
        pe.set_goal(i, np.array(i))                           # the collector passes a 0-d array
    tw.set_goal(torch.arange(B, dtype=torch.int32))
    g = pe.get_goal(3)
    assert g.shape == (2,) and np.allclose(g, obs[3]['zone_obs'][3, :2], atol=1e-6)
    with pytest.raises(AssertionError):
        pe.available_goals(0)                                 # a goal is set (TSP_next_city_env.py:88)
    rs = np.random.RandomState(4)
    tw.aux[:3, 3] = (tw.aux[:3, 3].view(torch.int32) & ~0xffff | 1998).view(torch.float32)
    pe.vec.aux[:3, 3] = (pe.vec.aux[:3, 3].view(torch.int32) & ~0xffff | 1998).view(torch.float32)
    for t in range(4):
        a = [rs.uniform(-1, 1, 2) for _ in range(B)]          # list of per-env arrays, as the algos pass
        o, r, d, info = pe.step_no_reset(a)
        o2, r2, d2, i2 = tw.step_no_reset(torch.tensor(np.array(a), dtype=torch.float32).cuda())
        assert all(isinstance(x, tuple) and len(x) == B for x in (o, r, d, info))
        assert isinstance(r[0], float) and isinstance(d[0], bool) and isinstance(info[0], dict)
        assert np.array_equal(np.stack([x['obs'] for x in o]), o2['obs'].cpu().numpy())
        assert np.array_equal(np.stack([x['zone_obs'] for x in o]), o2['zone_obs'].cpu().numpy())
        assert list(r) == [float(x) for x in r2.cpu().numpy()] and list(d) == [bool(x) for x in d2.cpu().numpy()]
        for i in range(B):
            if t >= 2 and i < 3:                              # parked by WaitWrapper after the step limit at t = 1
                assert info[i] == {} and d[i] and r[i] == 0.0 and not o[i]['obs'].any() and not o[i]['zone_obs'].any()
            else:
                assert info[i]['cost'] == 0 and 'goal_met' not in info[i]
                assert info[i]['need_next_goal'] == (t == 1 and i < 3)
                assert abs(info[i]['shaped_reward'] - float(i2['shaped_reward'][i].item())) == 0.0
    assert pe.needs_goal() == [True] * 3 + [False] * 3
    first = pe.reset()
    o, r, d, info = pe.step([np.zeros(2)] * B)                 # auto-reset variant: nobody is parked after reset
    assert not any(d) and all(i_['need_next_goal'] for i_ in info)   # reset cleared the goals
    assert first[0]['obs'][0] == 1.0 and o[0]['obs'][0] == np.float32(1999 / 2000)


def test_reference_factories_and_single_env_api(crl):
    """make_train_env / make_test_env / make_fixed_env (make_env.py:3-51) and the single-env gym
    surface evaluate.py:48-60 drives: same call shapes, and the seeding rules behind them."""
    envs = [crl.make_train_env('ColourMatch-v0', hier=False, num_training_tasks=4, rng_seed=1 + 10000 * i) for i in range(5)]
    assert envs[0].observation_space.spaces['zone_obs'].shape == (6, 7) and envs[0].action_space.shape == (2,)
    pe = crl.ParallelEnv(envs)                                # exactly what base.py:50 does with the list
    obs = pe.reset()
    assert len(obs) == 5 and set((pe.vec.seeds.cpu().numpy() - 1).tolist()) <= {1, 2, 3, 4}
    o, r, d, info = pe.step([np.zeros(2)] * 5)
    assert len(o) == 5 and info[0]['cost'] == 0
    # one fixed map, as evaluate.py:48-56 builds it; step() does not auto-reset
    single = crl.make_fixed_env('PointTSP-v0', hier=False, seed=3, env_seed=1000007)
    first = single.reset()
    z0 = first['zone_obs'].copy()
    assert first['obs'].shape == (8,) and first['obs'][0] == 1.0
    for t in range(5):
        o, r, d, info = single.step(np.array([0.5, -0.2]))
        assert isinstance(r, float) and isinstance(d, bool) and info == {'cost': 0} and o['zone_obs'].shape == (15, 6)
    assert o['obs'][0] == np.float32(1995 / 2000)
    assert np.array_equal(single.reset()['zone_obs'], z0)     # FixedSeedsWrapper(min = max = env_seed): same map again
    # make_test_env: seeded once, the next reset moves on to the next seed; its first map is make_fixed_env's of that seed
    test = crl.make_test_env('PointTSP-v0', seed=1000007)
    a = test.reset()['zone_obs'].copy()
    b = test.reset()['zone_obs'].copy()
    assert np.array_equal(a, z0) and not np.array_equal(a, b)
    nxt = crl.make_fixed_env('PointTSP-v0', env_seed=1000008)
    assert np.array_equal(nxt.reset()['zone_obs'], b)
    # goal variant through the single-env surface (visualize_hier.py:60-72)
    g = crl.make_fixed_env('PointTSP-v3', env_seed=5)
    ob = g.reset()
    assert g.goal_zone is None and g.get_available_goals().all()
    g.set_goal(np.array(4))
    assert g.goal_zone == 4 and np.allclose(g.get_goal(), ob['zone_obs'][4, :2], atol=1e-6)
    o, r, d, info = g.step(np.zeros(2))
    assert set(info) == {'cost', 'shaped_reward', 'need_next_goal'} and info['need_next_goal'] is False
