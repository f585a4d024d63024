"""CPU oracle, part 2: the slice of Safety Gym's ``Engine`` / ``World`` the
PointTSP / TimedTSP / ColourMatch hot path runs through.

TEST INFRASTRUCTURE ONLY (see oracle/mj_point.py header).
PARITY UNPINNED for this file: ``safety_gym/envs/engine.py`` and ``world.py``
are not under /root/reference (un-pinned sibling checkout, reference
``README.md:36-37``); this restates their published control flow as recorded in
SURVEY.md Appendix A.4.  The class keeps upstream's attribute surface so that
the reference's own ``ZoneEnvBase`` (``main/envs/zone_envs/ZoneEnvBase.py:33``,
``from safety_gym.envs.engine import *`` at ``:5``) can subclass it unmodified:
that is how tests/golden/gen_golden.py runs the REAL reference task code here.

Control flow restated (upstream names):
  seed(s)      _seed = s
  reset()      _seed += 1; rs = RandomState(_seed); done = False; steps = 0;
               build(); return obs()
  build()      build_layout() -> build_world_config() -> World (fresh sim,
               qpos = qvel = 0)
  sample_layout()  objects in placements order (robot, then zones); <= 100
               tries each of xy = (U(xmin+k, xmax-k), U(ymin+k, ymax-k)); valid
               iff dist >= keepout_a + keepout_b to every placed object; 100
               misses abandon the whole layout; build_layout retries <= 10000
  step(a)      ctrl = clip(a, ctrlrange); for _ in range(rs.binomial(10, 1.0)):
               set_mocaps(); sim.step()
               sim.forward(); reward(); cost(); goal_met() -> reward_goal, done;
               steps += 1; steps >= num_steps -> done; return obs(), ...
"""
from copy import deepcopy

import numpy as np

from . import mj_point
from .spaces import Box, Dict

# Names ``from safety_gym.envs.engine import *`` is expected to provide and that
# ZoneEnvBase.py mentions (only in branches the three tasks never take).
GROUP_GOAL = 0
GROUP_BOX = 1
GROUP_BUTTON = 1
GROUP_WALL = 2
GROUP_PILLAR = 2
GROUP_HAZARD = 3
GROUP_VASE = 4
GROUP_GREMLIN = 5
GROUP_CIRCLE = 6


class ResamplingError(AssertionError):
    pass


class _Robot:
    """``safety_gym.envs.world.Robot`` for xmls/point.xml: two actuators, body
    origin height 0.1."""

    def __init__(self, path):
        assert path.endswith('point.xml'), 'the oracle restates the Point robot only'
        self.nu = 2
        self.z_height = mj_point.BODY_Z
        self.hinge_vel_names = []
        self.hinge_pos_names = []
        self.ballangvel_names = []
        self.ballquat_names = []
        self.sensor_dim = {}


class _World:
    def __init__(self, config):
        self.config = config
        geoms = config.get('geoms', {})
        static = {name: np.asarray(g['pos'], dtype=np.float64) for name, g in geoms.items()}
        rgba = {name: np.asarray(g['rgba'], dtype=np.float64) for name, g in geoms.items()}
        # box geoms with collisions on = the walls (Engine.build_world_config [upstream]: size = walls_size in all three
        # dimensions, centre height = walls_size -- the height of the robot's sphere centre for walls_size 0.1)
        boxes = [g for g in geoms.values() if g.get('type') == 'box' and g.get('group') == GROUP_WALL]
        self.sim = mj_point.PointSim(config['robot_xy'], config['robot_rot'], static, rgba,
                                     wall_boxes=[g['pos'][:2] for g in boxes] if boxes else None,
                                     wall_half=float(boxes[0]['size'][0]) if boxes else 0.1)
        self.model = self.sim.model
        self.data = self.sim.data

    def robot_pos(self):
        return self.data.get_body_xpos('robot').copy()

    def robot_vel(self):
        return self.data.get_body_xvelp('robot').copy()


class Env:
    """gym.Env surface (``unwrapped``) without gym."""
    metadata = {}

    @property
    def unwrapped(self):
        return self


class Engine(Env):
    # Only the keys the three tasks read or override; upstream has ~150 more,
    # all inert here (lidar, vases, gremlins, vision, buttons, ...).
    DEFAULT = {
        'num_steps': 1000,
        'action_noise': 0.0,
        'placements_extents': [-2, -2, 2, 2],
        'placements_margin': 0.0,
        'floor_display_mode': False,
        'robot_placements': None,
        'robot_locations': [],
        'robot_keepout': 0.4,
        'robot_base': 'xmls/car.xml',
        'robot_rot': None,
        'randomize_layout': True,
        'build_resample': True,
        'continue_goal': True,
        'terminate_resample_failure': True,
        'observation_flatten': True,
        'observe_sensors': True,
        'observe_goal_dist': False,
        'observe_goal_comp': False,
        'observe_goal_lidar': False,
        'observe_box_comp': False,
        'observe_box_lidar': False,
        'observe_circle': False,
        'observe_remaining': False,
        'observe_walls': False,
        'observe_hazards': False,
        'observe_vases': False,
        'observe_pillars': False,
        'observe_buttons': False,
        'observe_gremlins': False,
        'observe_vision': False,
        'observe_qpos': False,
        'observe_qvel': False,
        'observe_ctrl': False,
        'observe_freejoint': False,
        'observe_com': False,
        'render_labels': False,
        'render_lidar_markers': True,
        'render_lidar_radius': 0.15,
        'render_lidar_size': 0.025,
        'render_lidar_offset_init': 0.5,
        'render_lidar_offset_delta': 0.06,
        'vision_size': (60, 40),
        'vision_render': True,
        'vision_render_size': (300, 200),
        'lidar_num_bins': 10,
        'lidar_max_dist': None,
        'lidar_exp_gain': 1.0,
        'lidar_type': 'pseudo',
        'lidar_alias': True,
        'task': 'goal',
        'reward_goal': 1.0,
        'reward_exception': -10.0,
        'walls_num': 0,
        'walls_placements': None,
        'walls_locations': [],
        'walls_keepout': 0.0,
        'walls_size': 0.5,
        'gremlins_num': 0,
        'pillars_num': 0,
        'buttons_num': 0,
        'hazards_num': 0,
        'vases_num': 0,
        'constrain_hazards': False,
        'constrain_vases': False,
        'constrain_pillars': False,
        'constrain_buttons': False,
        'constrain_gremlins': False,
        'constrain_indicator': True,
        'frameskip_binom_n': 10,
        'frameskip_binom_p': 1.0,
        '_seed': None,
    }

    def __init__(self, config={}):
        self.parse(config)
        self.robot = _Robot(self.robot_base)
        self.action_space = Box(-1, 1, (self.robot.nu,), dtype=np.float32)
        self.build_observation_space()
        self.build_placements_dict()
        self.viewer = None
        self.world = None
        self.clear()
        self.seed(self._seed)
        self.done = True

    def parse(self, config):
        self.config = deepcopy(self.DEFAULT)
        self.config.update(deepcopy(config))
        for key, value in self.config.items():
            assert key in self.DEFAULT, f'Bad key {key}'
            setattr(self, key, value)

    @property
    def sim(self):
        return self.world.sim

    @property
    def model(self):
        return self.sim.model

    @property
    def data(self):
        return self.sim.data

    def clear(self):
        self.layout = None

    def seed(self, seed=None):
        self._seed = np.random.randint(2 ** 32) if seed is None else seed

    # ---- spaces and placements (built once in __init__) -------------------
    def build_observation_space(self):
        obs_space_dict = {}
        if self.observe_remaining:
            obs_space_dict['remaining'] = Box(0.0, 1.0, (1,), dtype=np.float32)
        assert not self.observe_sensors, 'sensor observations are outside the oracle'
        self.obs_space_dict = obs_space_dict
        if self.observation_flatten:
            self.obs_flat_size = sum(int(np.prod(b.shape)) for b in obs_space_dict.values())
            self.observation_space = Box(-np.inf, np.inf, (self.obs_flat_size,), dtype=np.float32)
        else:
            self.observation_space = Dict(obs_space_dict)

    def placements_dict_from_object(self, object_name):
        if object_name == 'robot':
            num, fmt, prefix = 1, 'robot', 'robot_'
        else:
            num, fmt, prefix = getattr(self, object_name + 's_num'), object_name + '{i}', object_name + 's_'
        locations = getattr(self, prefix + 'locations', [])
        placements = getattr(self, prefix + 'placements', None)
        keepout = getattr(self, prefix + 'keepout')
        out = {}
        for i in range(num):
            if i < len(locations):
                x, y = locations[i]
                k = keepout + 1e-9
                pl = [(x - k, y - k, x + k, y + k)]
            else:
                pl = placements
            out[fmt.format(i=i)] = (pl, keepout)
        return out

    def build_placements_dict(self):
        placements = {}
        placements.update(self.placements_dict_from_object('robot'))
        placements.update(self.placements_dict_from_object('wall'))
        self.placements = placements

    # ---- layout ------------------------------------------------------------
    def random_rot(self):
        return self.rs.uniform(0, 2 * np.pi)

    def constrain_placement(self, placement, keepout):
        xmin, ymin, xmax, ymax = placement
        return (xmin + keepout, ymin + keepout, xmax - keepout, ymax - keepout)

    def draw_placement(self, placements, keepout):
        if placements is None:
            choice = self.constrain_placement(self.placements_extents, keepout)
        else:
            constrained = []
            for placement in placements:
                xmin, ymin, xmax, ymax = self.constrain_placement(placement, keepout)
                if xmin > xmax or ymin > ymax:
                    continue
                constrained.append((xmin, ymin, xmax, ymax))
            assert len(constrained), 'Failed to find any placements with satisfy keepout'
            if len(constrained) == 1:
                choice = constrained[0]
            else:
                areas = [(x2 - x1) * (y2 - y1) for x1, y1, x2, y2 in constrained]
                probs = np.array(areas) / np.sum(areas)
                choice = constrained[self.rs.choice(len(constrained), p=probs)]
        xmin, ymin, xmax, ymax = choice
        return np.array([self.rs.uniform(xmin, xmax), self.rs.uniform(ymin, ymax)])

    def sample_layout(self):
        layout = {}
        for name, (placements, keepout) in self.placements.items():
            conflicted = True
            for _ in range(100):
                xy = self.draw_placement(placements, keepout)
                ok = True
                for other_name, other_xy in layout.items():
                    other_keepout = self.placements[other_name][1]
                    dist = np.sqrt(np.sum(np.square(xy - other_xy)))
                    if dist < other_keepout + self.placements_margin + keepout:
                        ok = False
                        break
                if ok:
                    conflicted = False
                    break
            if conflicted:
                return False
            layout[name] = xy
        self.layout = layout
        return True

    def build_layout(self):
        if not self.randomize_layout:
            self.rs = np.random.RandomState(0)
        for _ in range(10000):
            if self.sample_layout():
                break
        else:
            raise ResamplingError('Failed to sample layout of objects')

    def build_world_config(self):
        world_config = {}
        world_config['robot_base'] = self.robot_base
        world_config['robot_xy'] = self.layout['robot']
        if self.robot_rot is None:
            world_config['robot_rot'] = self.random_rot()
        else:
            world_config['robot_rot'] = float(self.robot_rot)
        world_config['objects'] = {}
        world_config['geoms'] = {}
        for i in range(self.walls_num):                            # Engine.build_world_config [upstream]
            name = f'wall{i}'
            assert self.walls_size == mj_point.BODY_Z, 'the oracle\'s planar sphere-box contact needs walls_size == 0.1'
            world_config['geoms'][name] = {'name': name, 'size': np.ones(3) * self.walls_size,
                                           'pos': np.r_[self.layout[name], self.walls_size], 'rot': 0, 'type': 'box',
                                           'group': GROUP_WALL, 'rgba': np.array([.5, .5, .5, 1.0])}
        return world_config

    def build(self):
        self.build_layout()
        self.world_config_dict = self.build_world_config()
        self.world = _World(self.world_config_dict)

    # ---- episode -----------------------------------------------------------
    def reset(self):
        self._seed += 1
        self.rs = np.random.RandomState(self._seed)
        self.done = False
        self.steps = 0
        self.clear()
        self.build()
        cost = self.cost()
        assert cost['cost'] == 0, f'World has starting cost! {cost}'
        return self.obs()

    def dist_xy(self, pos):
        pos = np.asarray(pos)
        if pos.shape == (3,):
            pos = pos[:2]
        robot_pos = self.world.robot_pos()
        return np.sqrt(np.sum(np.square(pos - robot_pos[:2])))

    def set_mocaps(self):
        pass

    def reward(self):
        return 0.0

    def goal_met(self):
        return False

    def cost(self):
        self.sim.forward()
        cost = {}
        cost['cost'] = sum(v for k, v in cost.items() if k.startswith('cost_'))
        self._cost = cost
        return cost

    def obs(self):
        raise NotImplementedError('ZoneEnvBase overrides obs()')

    def step(self, action):
        action = np.asarray(action)
        assert not self.done, 'Environment must be reset before stepping'
        info = {}
        rng = self.model.actuator_ctrlrange
        self.data.ctrl[:] = np.clip(action, rng[:, 0], rng[:, 1])
        if self.action_noise:
            self.data.ctrl[:] += self.action_noise * self.rs.randn(self.model.nu)
        for _ in range(self.rs.binomial(self.frameskip_binom_n, self.frameskip_binom_p)):
            self.set_mocaps()
            self.sim.step()
        self.sim.forward()
        reward = self.reward()
        info.update(self.cost())
        if self.goal_met():
            info['goal_met'] = True
            reward += self.reward_goal
            assert not self.continue_goal, 'continue_goal is outside the oracle'
            self.done = True
        self.steps += 1
        if self.steps >= self.num_steps:
            self.done = True
        return self.obs(), reward, self.done, info
