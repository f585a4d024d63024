"""CPU oracle, part 3: the task logic, observation build and vector-env
protocol of PointTSP-v0 / PointTTSP-v0 / ColourMatch-v0, restated on flat
arrays (one Python object per env, fp64 like the reference).

TEST INFRASTRUCTURE ONLY (see oracle/mj_point.py header).

PINNED by tests/golden/*.npz: tests/golden/gen_golden.py imports the REAL
reference modules (``main/envs/{TSP,TTSP,colour_match}_env.py``,
``zone_envs/ZoneEnvBase.py``, ``wrappers.py``, ``make_env.py``) over the stubbed
upstream packages and records their outputs; tests/test_oracle_golden.py replays
the same seeds and actions through this file and requires identical results.
(The physics and layout sampler underneath both are oracle/mj_point.py and
oracle/sg_engine.py -- parity unpinned, see their headers.)

What each piece follows:
  per-step order ............ Engine.step [upstream] + SURVEY.md Appendix B
  visit event ............... TSP_env.py:54-69 (first unvisited zone, index
                              order, fp64 sqrt(dx^2+dy^2) <= 0.2, PRE-physics pos)
  TSP reward / bonus / goal . TSP_env.py:37-42, 71-72
  TimedTSP timeouts ......... TTSP_env.py:19-27 (RandomState(_seed) BEFORE the
                              Engine increments it), failure :62-71
  ColourMatch ............... colour_match_env.py:26-36 (cycle B->G->R->B, cooldown
                              150), :38-55 (Hamming), :57-68 (reset), :86-101
  observation ............... ZoneEnvBase.py:143-235, obs_zones in the three task
                              files, ZoneWrapper.split_zone_obs wrappers.py:136-142
  per-episode seeding ....... FixedSeedsWrapper wrappers.py:10-23
  auto-reset ................ penv.py:4-21, 52-66
  goal-conditioned variants . zone-goals/envs/TSP_next_city_env.py (PointTSP-v3),
                              TTSP_next_city_env.py (PointTTSP-v3),
                              colour_match_next_city_env.py (ColourMatch-v3): set_goal /
                              get_goal / get_available_goals, info['shaped_reward'],
                              info['need_next_goal']; worker RPCs
                              zone-goals/src/torch_ac/torch_utils/penv.py:18-25
  WaitWrapper ............... wrappers.py:29-54 (no-op zeros after the inner env is done)
  hard instances ............ PointTSP-v4 / v5: TSP_hard_env.py:11-31 over the configs of
                              main/envs/__init__.py:52-81 (fixed robot / city locations,
                              `zones_colours`: Cyan cities + Yellow distractors that start
                              visited); the goal-conditioned flavour zone-goals/envs/TSP_hard_env.py
"""
import math

import numpy as np

from . import mj_point

TSP, TTSP, CM = 0, 1, 2
TASK_OF_ENV_ID = {'PointTSP-v0': TSP, 'PointTTSP-v0': TTSP, 'ColourMatch-v0': CM,
                  'PointTSP-v3': TSP, 'PointTTSP-v3': TTSP, 'ColourMatch-v3': CM}
GOAL_ENV_IDS = ('PointTSP-v3', 'PointTTSP-v3', 'ColourMatch-v3',   # zone-goals/envs/__init__.py
                'zone-goals/PointTSP-v4', 'zone-goals/PointTSP-v5')

# Hard instances (main/envs/__init__.py:52-81; zone enum ZoneEnvBase.py:13-21: 6 = Cyan = a city,
# 5 = Yellow = visited from the start).  The zone-goals registrations differ in v5's num_steps
# only (zone-goals/envs/__init__.py:69-81) and step through TSPNextCityEnv.
HARD = {
    'PointTSP-v4': dict(num_zones=15, num_steps=1000,
                        zones_locations=[(-2.6, -1.6), (-0., -0.5), (1., 0.5), (1.8, 1.5), (2.6, 2.6)],
                        zones_colours=[6] * 5 + [5] * 10, robot_locations=[(-0.9, -0.9)], robot_rot=-1),
    'PointTSP-v5': dict(num_zones=15, num_steps=250,
                        zones_locations=[(-2.6, -2.6), (-2, -1.6), (2, 1)],
                        zones_colours=[6] * 3 + [5] * 12, robot_locations=[(0.8, 0.8)], robot_rot=None),
}
HARD['zone-goals/PointTSP-v4'] = dict(HARD['PointTSP-v4'])
HARD['zone-goals/PointTSP-v5'] = dict(HARD['PointTSP-v5'], num_steps=300)
TASK_OF_ENV_ID.update({k: TSP for k in HARD})
# `walled=True` builds of the registered configs (make_task_env; tests/golden/gen_golden_walls.py)
TASK_OF_ENV_ID.update({'walled/' + k: v for k, v in list(TASK_OF_ENV_ID.items()) if k in ('PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0')})

NUM_STEPS = 2000            # envs/__init__.py:13, :49
NUM_ZONES = {TSP: 15, TTSP: 15, CM: 6}   # envs/__init__.py:9, :45
ZONE_DIM = {TSP: 6, TTSP: 7, CM: 7}
ZONE_SIZE = 0.2             # ZoneEnvBase.py:51
ZONE_KEEPOUT = 0.55         # ZoneEnvBase.py:50
ROBOT_KEEPOUT = 0.4         # Engine.DEFAULT [upstream]
EXTENT = 3.0                # ZoneEnvBase.py:41,52
FRAMESKIP = 10              # Engine.DEFAULT frameskip_binom_n, p = 1.0 [upstream]
MAX_COOLDOWN = 150          # colour_match_env.py:16
TIME_SAVED_REWARD = 0.01    # TSP_env.py:14
BETA_A, BETA_B = 3, 1.5     # TTSP_env.py:13

# colour codes: 0 Blue, 1 Green, 2 Red (the order of ``colours``,
# colour_match_env.py:9); TSP zones: Cyan = unvisited, Yellow = visited.
RGB_CM = np.array([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])
RGB_UNVISITED = np.array([0., 1., 1.])
RGB_VISITED = np.array([1., 1., 0.])
ALPHA = 0.25


def hamming_to_goal(colours):
    """colour_match_env.py:38-55."""
    nb = int(np.sum(colours == 0))
    ng = int(np.sum(colours == 1))
    nr = int(np.sum(colours == 2))
    return min(2 * ng + nr, 2 * nr + nb, 2 * nb + ng)


def wall_locations(extent=3):
    """The box centres of a `walled=True` env, in the reference's order (ZoneEnvBase.py:55-57): the two rows y = -+extent
    point by point in x (steps of 0.1), then the two columns x = -+extent point by point in y."""
    w = [(i / 10, j) for i in range(int(-extent * 10), int(extent * 10 + 1), 1) for j in [-extent, extent]]
    w += [(i, j / 10) for i in [-extent, extent] for j in range(int(-extent * 10), int(extent * 10 + 1), 1)]
    return w


WALLS_SIZE = 0.1          # ZoneEnvBase.py:61
WALLS_KEEPOUT = 0.0       # Engine default [upstream]


def sample_layout(rs, num_zones, robot_locations=(), zones_locations=(), robot_rot=None, walls=(), walls_out=None):
    """Engine.build_layout / sample_layout / build_world_config [upstream],
    specialised to: robot (keepout .4) then ``num_zones`` zones (keepout .55),
    extents +-3.  Consumes ``rs`` exactly as upstream does
    (x then y per candidate; robot_rot after the layout; one cosmetic
    ``random_rot`` per zone, ZoneEnvBase.py:132).  An object with a fixed location
    (``robot_locations[0]``, ``zones_locations[i]``; placements_dict_from_object) is drawn
    -- still consuming two uniforms -- from the box of half-width keepout + 1e-9 around it
    shrunk by keepout, i.e. within 1e-9 of the location, and must pass the same keepout
    test.  A fixed ``robot_rot`` draws nothing.  Returns xy0, rot0, zone_xy."""
    # build_placements_dict order [upstream]: robot, walls (fixed locations, keepout 0: they never reject anything but
    # each consumes two uniforms per layout attempt), then the zones (ZoneEnvBase.py:118-122)
    keepouts = [ROBOT_KEEPOUT] + [WALLS_KEEPOUT] * len(walls) + [ZONE_KEEPOUT] * num_zones
    locations = [robot_locations[0] if len(robot_locations) else None] + list(walls)
    locations += [zones_locations[i] if i < len(zones_locations) else None for i in range(num_zones)]
    for _ in range(10000):
        placed = []
        ok_layout = True
        for k, loc in zip(keepouts, locations):
            if loc is None:
                xlo = ylo = -EXTENT + k
                xhi = yhi = EXTENT - k
            else:
                kk = k + 1e-9
                xlo, ylo, xhi, yhi = loc[0] - kk + k, loc[1] - kk + k, loc[0] + kk - k, loc[1] + kk - k
            found = False
            for _try in range(100):
                xy = np.array([rs.uniform(xlo, xhi), rs.uniform(ylo, yhi)])
                if all(np.sqrt(np.sum(np.square(xy - oxy))) >= ok + k for oxy, ok in placed):
                    found = True
                    break
            if not found:
                ok_layout = False
                break
            placed.append((xy, k))
        if ok_layout:
            break
    else:
        raise RuntimeError('layout sampling failed')
    rot0 = rs.uniform(0, 2 * np.pi) if robot_rot is None else float(robot_rot)
    for _ in range(num_zones):
        rs.uniform(0, 2 * np.pi)
    if walls_out is not None:        # where the boxes really are: each drawn within 1e-9 of its location
        walls_out[:] = [p for p, _ in placed[1:1 + len(walls)]]
    return placed[0][0], rot0, np.array([p for p, _ in placed[1 + len(walls):]])


class ZoneTaskEnv:
    """One env, equivalent to ``ZoneWrapper(gym.make(env_id))``.

    ``seed(s)`` / ``reset()`` / ``step(a)`` follow the reference; ``reset`` also
    accepts an explicit ``layout`` dict (host-supplied-layout mode) with keys
    xy0 (2,), rot0, zone_xy (N,2) and, per task, zone_max_steps (N,) / colours (N,).
    """

    def __init__(self, task, num_zones=None, num_steps=NUM_STEPS, goals=False, hard=None, walled=False):
        self.task = task
        self.walls = wall_locations(EXTENT) if walled else []   # `walled=True` (ZoneEnvBase.py:39,55-62)
        self.wall_xy = list(self.walls)                         # as placed by the last sampled layout
        # hard instance: dict with zones_locations, zones_colours, robot_locations, robot_rot
        self.hard = hard
        self.goals = goals          # the *_next_city_env.py variants
        self.goal_zone = None
        self.last_dist_to_goal = None
        self.N = NUM_ZONES[task] if num_zones is None else num_zones
        self.Z = ZONE_DIM[task]
        self.num_steps = num_steps
        self._seed = None
        self.done = True
        self.substep_trace = None   # set to a list to record (qpos, qvel) after each substep

    # -- seeding and reset ---------------------------------------------------
    def seed(self, seed):
        self._seed = seed

    def reset(self, layout=None):
        N = self.N
        if layout is None:
            # task-level draws use the seed BEFORE Engine.reset increments it
            if self.task == TTSP:
                rs = np.random.RandomState(self._seed)
                self.zone_max_steps = np.array([int(rs.beta(BETA_A, BETA_B) * self.num_steps) for _ in range(N)])
            elif self.task == CM:
                rs = np.random.RandomState(self._seed)
                self.colours = np.array([int(rs.choice(3)) for _ in range(N)])
            self._seed += 1
            self.rs = np.random.RandomState(self._seed)
            h = self.hard or {}
            xy0, rot0, zone_xy = sample_layout(self.rs, N, h.get('robot_locations', ()),
                                               h.get('zones_locations', ()), h.get('robot_rot'), walls=self.walls,
                                               walls_out=self.wall_xy)
        else:
            xy0, rot0, zone_xy = layout['xy0'], layout['rot0'], layout['zone_xy']
            if self.task == TTSP:
                self.zone_max_steps = np.asarray(layout['zone_max_steps']).astype(np.int64)
            elif self.task == CM:
                self.colours = np.asarray(layout['colours']).astype(np.int64).copy()
            self.rs = None
        self.xy0 = np.asarray(xy0, dtype=np.float64)
        self.rot0 = float(rot0)
        self.zone_xy = np.asarray(zone_xy, dtype=np.float64).reshape(N, 2)
        self.visited = np.zeros(N, dtype=bool)
        if self.hard is not None:       # TSP_hard_env.py:27-30: zones restart from zones_colours
            self.visited = np.array([c == 5 for c in self.hard['zones_colours']], dtype=bool)
        if self.task == CM:
            self.cooldown = np.zeros(N, dtype=np.int64)
            self.goal_dist = hamming_to_goal(self.colours)
        self.sim = mj_point.PointSim(self.xy0, self.rot0, wall_boxes=self.wall_xy or None, wall_half=WALLS_SIZE)
        self.steps = 0
        self.done = False
        self.event = 0
        return self._obs()

    def set_state(self, qpos, qvel):
        """Teacher forcing for parity tests: overwrite the physics state."""
        self.sim.data.qpos = np.asarray(qpos, dtype=np.float64).copy()
        self.sim.data.qvel = np.asarray(qvel, dtype=np.float64).copy()
        self.sim.forward()

    # -- one env step -----------------------------------
    def step(self, action):
        assert not self.done, 'Environment must be reset before stepping'
        N = self.N
        if self.task == CM:
            self.cooldown[self.cooldown > 0] -= 1
        ctrl = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        self.sim.data.ctrl[:] = ctrl
        if self.rs is not None:
            self.rs.binomial(FRAMESKIP, 1.0)    # upstream draws the frameskip

        # zone event on the pre-physics position
        p = self.sim.data.get_body_xpos('robot')[:2]
        fired = -1
        for i in range(N):
            eligible = (self.cooldown[i] == 0) if self.task == CM else (not self.visited[i])
            if eligible and np.sqrt(np.sum(np.square(self.zone_xy[i] - p))) <= ZONE_SIZE:
                fired = i
                break
        reward = 0
        if fired >= 0:
            if self.task == CM:
                self.colours[fired] = (self.colours[fired] + 1) % 3
                self.cooldown[fired] = MAX_COOLDOWN
            else:
                self.visited[fired] = True

        for _ in range(FRAMESKIP):
            self.sim.step()
            if self.substep_trace is not None:
                self.substep_trace.append((self.sim.data.qpos.copy(), self.sim.data.qvel.copy()))
        self.sim.forward()

        if self.task == CM:
            if fired >= 0:
                new_dist = hamming_to_goal(self.colours)
                reward = self.goal_dist - new_dist
                self.goal_dist = new_dist
            goal = self.goal_dist == 0
        else:
            reward = 1 if fired >= 0 else 0
            goal = bool(self.visited.all())
        self.event = reward            # integer reward component
        info = {'cost': 0}
        if goal:
            info['goal_met'] = True
            reward = reward + (self.num_steps - self.steps) * TIME_SAVED_REWARD
            self.done = True
        self.steps += 1
        if self.steps >= self.num_steps:
            self.done = True
        if self.goals:
            # TSP_next_city_env.py:57-75 / colour_match_next_city_env.py:110-133, evaluated right
            # after Engine.step (post-physics position), before TimedTSP's timeout test
            assert self.goal_zone is not None
            reached = fired >= 0 and fired == self.goal_zone
            if reached:
                info['shaped_reward'] = 0.
            else:
                dist = self._dist_to_goal()
                info['shaped_reward'] = self.last_dist_to_goal - dist
                self.last_dist_to_goal = dist
                if self.task == CM and fired >= 0:
                    info['shaped_reward'] -= 1.
            if reached or self.done:
                info['need_next_goal'] = True
                self.goal_zone = None
            else:
                info['need_next_goal'] = False
        if self.task == TTSP and not self.done:
            if (self._zone_times() <= 0).any():
                self.done = True
                if self.goals:          # TTSP_next_city_env.py:49-53
                    info['need_next_goal'] = True
                    self.goal_zone = None
        return self._obs(), reward, self.done, info

    # -- goal RPCs (TSP_next_city_env.py:77-95, colour_match_next_city_env.py:135-148) ---
    def _dist_to_goal(self):
        p = self.sim.data.get_body_xpos('robot')[:2]
        return np.sqrt(np.sum(np.square(self.zone_xy[self.goal_zone] - p)))

    def set_goal(self, next_goal):
        next_goal = int(next_goal)
        if self.task == CM:
            assert 0 <= next_goal < self.N
        else:
            assert not self.visited[next_goal]
        self.goal_zone = next_goal
        self.last_dist_to_goal = self._dist_to_goal()

    def get_goal(self):
        assert self.goal_zone is not None
        return self.zone_xy[self.goal_zone] / 3.

    def get_available_goals(self):
        assert self.goal_zone is None
        if self.task == CM:
            return np.ones(self.N, dtype=bool)
        return ~self.visited

    # -- observation -----------------------------------------------------------
    def _zone_times(self):
        t = (self.zone_max_steps - self.steps) / self.num_steps
        t[self.visited] = 1.0
        return t

    def _obs(self):
        self.sim.forward()
        d = self.sim.data
        N, Z = self.N, self.Z
        zone_obs = np.zeros((N, Z))
        zone_obs[:, 0:2] = self.zone_xy / 3.0
        if self.task == CM:
            zone_obs[:, 2:5] = RGB_CM[self.colours]
            zone_obs[:, 6] = (self.cooldown.astype(np.float32) / np.float32(MAX_COOLDOWN)).astype(np.float64)
        else:
            zone_obs[:, 2:5] = np.where(self.visited[:, None], RGB_VISITED, RGB_UNVISITED)
            if self.task == TTSP:
                zone_obs[:, 6] = self._zone_times()
        zone_obs[:, 5] = ALPHA
        quat = d.get_body_xquat('robot').astype(np.float32)
        direction = np.array([quat[0] ** 2 - quat[3] ** 2, 2 * quat[0] * quat[3]])
        obs = np.concatenate([
            [1.0 - self.steps / self.num_steps],
            d.get_body_xpos('robot')[:2] / 3.0,
            direction,
            d.get_body_xvelp('robot')[:2] / 1.5,
            [d.get_body_xvelr('robot')[2] / 3.0],
        ])
        return {'zone_obs': zone_obs, 'obs': obs}

    # -- state export (world frame), used by the CUDA parity tests -------------
    def world_state(self):
        d = self.sim.data
        p = d.get_body_xpos('robot')
        v = d.get_body_xvelp('robot')
        return np.array([p[0], p[1], self.rot0 + d.qpos[2], v[0], v[1], d.qvel[2]])


class FixedSeeds:
    """FixedSeedsWrapper (wrappers.py:10-23) around a ZoneTaskEnv."""

    def __init__(self, env, min_seed, max_seed, rng_seed=0):
        self.env = env
        self.min_seed, self.max_seed = min_seed, max_seed
        self.rng = np.random.default_rng(seed=rng_seed)

    def reset(self):
        new_seed = self.rng.integers(low=self.min_seed, high=self.max_seed + 1, size=1)[0]
        self.env.seed(new_seed)
        return self.env.reset()

    def step(self, action):
        return self.env.step(action)

    def __getattr__(self, name):            # gym.Wrapper forwards unknown attributes (set_goal, goal_zone, ...)
        return getattr(self.env, name)


class SerialVecEnv:
    """ParallelEnv (penv.py:23-69) without the processes: same results."""

    def __init__(self, envs):
        self.envs = envs

    def reset(self):
        return [e.reset() for e in self.envs]

    def step(self, actions):
        out = []
        for e, a in zip(self.envs, actions):
            obs, reward, done, info = e.step(a)
            if done:
                obs = e.reset()
            out.append((obs, reward, done, info))
        return tuple(zip(*out))

    def step_no_reset(self, actions):
        return tuple(zip(*[e.step(a) for e, a in zip(self.envs, actions)]))


def make_task_env(env_id):
    """gym.make(env_id) for the registrations this oracle covers."""
    if env_id in HARD:
        h = HARD[env_id]
        return ZoneTaskEnv(TSP, num_zones=h['num_zones'], num_steps=h['num_steps'], goals=env_id in GOAL_ENV_IDS,
                           hard=h)
    if env_id.startswith('walled/'):
        # no registration of the reference says walled=True (main/envs/__init__.py:7-50); 'walled/<id>' names the env
        # the real class builds from that id's config with walled=True (tests/golden/gen_golden_walls.py)
        return ZoneTaskEnv(TASK_OF_ENV_ID[env_id[len('walled/'):]], walled=True)
    return ZoneTaskEnv(TASK_OF_ENV_ID[env_id], goals=env_id in GOAL_ENV_IDS)


def make_fixed_env(env_id, seed=1000, env_seed=0):
    """make_env.make_fixed_env (make_env.py:37-51) for the three zone tasks."""
    return FixedSeeds(make_task_env(env_id), env_seed, env_seed, rng_seed=seed)


def make_train_env(env_id, num_training_tasks=100, rng_seed=0, hier=False):
    """make_env.make_train_env (make_env.py:3-18)."""
    env = FixedSeeds(make_task_env(env_id), 1, num_training_tasks, rng_seed=rng_seed)
    return Wait(env) if hier else env


class Wait:
    """WaitWrapper (wrappers.py:29-54): stepping an env whose inner env is done is a no-op that
    returns an all-zero observation, zero reward, done=True and an empty info."""

    def __init__(self, env):
        self.env = env
        self.inner_done = False

    def step(self, action):
        if not self.inner_done:
            obs, rew, done, info = self.env.step(action)
            if done:
                self.inner_done = True
        else:
            obs, rew, done, info = self.noop_obs(), 0, True, {}
        return obs, rew, done, info

    def noop_obs(self):
        e = self.env.env if hasattr(self.env, 'env') else self.env
        return {'zone_obs': np.zeros((e.N, e.Z)), 'obs': np.zeros(8)}

    def reset(self):
        self.inner_done = False
        return self.env.reset()

    def __getattr__(self, name):
        return getattr(self.env, name)
