"""Per-task constants, from the reference's gym registrations and env classes."""
from dataclasses import dataclass

from . import _lib


@dataclass(frozen=True)
class TaskSpec:
    task: int
    num_zones: int
    num_steps: int
    zone_dim: int
    frameskip: int = 10            # Engine frameskip_binom_n with p = 1.0
    max_cooldown: int = 150        # colour_match_env.py:16
    zone_size: float = 0.2         # ZoneEnvBase.py:51
    time_saved_reward: float = 0.01  # TSP_env.py:14, colour_match_env.py:14
    beta_a: float = 3.0            # TTSP_env.py:13
    beta_b: float = 1.5
    robot_keepout: float = 0.4     # Engine default
    zone_keepout: float = 0.55     # ZoneEnvBase.py:50
    extent: float = 3.0            # ZoneEnvBase.py:41
    goals: bool = False            # goal-conditioned "next city" variant (zone-goals/envs/*_next_city_env.py)
    walled: bool = False           # `walled=True` (ZoneEnvBase.py:39,55-62): wall boxes around the arena; no registration sets it
    # hard instances (TSP_hard_env.py over main/envs/__init__.py:52-81): Engine's robot_locations[0],
    # robot_rot, zones_locations (the first len() zones) and zones_colours (5 = Yellow = starts visited,
    # 6 = Cyan = a city; zone enum ZoneEnvBase.py:13-21)
    robot_location: tuple = None
    robot_rot: float = None
    zones_locations: tuple = ()
    zones_colours: tuple = None

    @property
    def initial_visited(self):
        if self.zones_colours is None:
            return 0
        return sum(1 << i for i, c in enumerate(self.zones_colours) if c == 5)

    def fixed_layout(self):
        """The float32 (1 + N, 4) table of CrlState.fixed_layout, or None when nothing is fixed.
        Raises if two fixed objects violate each other's keepout (the reference would then retry its
        layout 10,000 times and fail, Engine.build_layout)."""
        if self.robot_location is None and self.robot_rot is None and not self.zones_locations:
            return None
        import numpy as np
        t = np.zeros((1 + self.num_zones, 4), dtype=np.float32)
        if self.robot_location is not None:
            t[0, :2], t[0, 3] = self.robot_location, 1
        if self.robot_rot is not None:
            t[0, 2] = self.robot_rot
            t[0, 3] += 2
        for i, xy in enumerate(self.zones_locations):
            t[1 + i, :2], t[1 + i, 3] = xy, 1
        keep = [self.robot_keepout] + [self.zone_keepout] * self.num_zones
        pinned = [k for k in range(1 + self.num_zones) if t[k, 3] in (1, 3)]
        for a in pinned:
            for b in pinned:
                if a < b and np.hypot(*(t[a, :2].astype(np.float64) - t[b, :2])) < keep[a] + keep[b]:
                    raise ValueError(f'fixed objects {a} and {b} are closer than their keepouts allow')
        return t


# main/envs/__init__.py:7-14 (config_point), :16-23 (config_point_easy), :43-50 (config_point_colour)
ENV_SPECS = {
    'PointTSP-v0': TaskSpec(_lib.TASK_TSP, 15, 2000, 6),       # :88-90
    'PointTSP-v1': TaskSpec(_lib.TASK_TSP, 5, 1000, 6),        # :94-96
    'PointTTSP-v0': TaskSpec(_lib.TASK_TTSP, 15, 2000, 7),     # :130-132
    'PointTTSP-v1': TaskSpec(_lib.TASK_TTSP, 5, 1000, 7),      # :134-136
    'ColourMatch-v0': TaskSpec(_lib.TASK_CM, 6, 2000, 7),      # :139-141
    # zone-goals/envs/__init__.py:105-156: same maps and dynamics, plus set_goal / shaped_reward /
    # need_next_goal (TSP_next_city_env.py, TTSP_next_city_env.py, colour_match_next_city_env.py)
    'PointTSP-v3': TaskSpec(_lib.TASK_TSP, 15, 2000, 6, goals=True),
    'PointTTSP-v3': TaskSpec(_lib.TASK_TTSP, 15, 2000, 7, goals=True),
    'ColourMatch-v3': TaskSpec(_lib.TASK_CM, 6, 2000, 7, goals=True),
}

# Hard instances, main/envs/__init__.py:52-81 and :110-118 (TSPHardEnv, TSP_hard_env.py): a few cities
# at fixed places, a fixed start, and distractor zones that are visited from the start.
_ZONES_1 = ((-2.6, -1.6), (-0., -0.5), (1., 0.5), (1.8, 1.5), (2.6, 2.6))
_ZONES_2 = ((-2.6, -2.6), (-2, -1.6), (2, 1))
ENV_SPECS.update({
    'PointTSP-v4': TaskSpec(_lib.TASK_TSP, 15, 1000, 6, robot_location=(-0.9, -0.9), robot_rot=-1.0,
                            zones_locations=_ZONES_1, zones_colours=(6,) * 5 + (5,) * 10),
    'PointTSP-v5': TaskSpec(_lib.TASK_TSP, 15, 250, 6, robot_location=(0.8, 0.8),
                            zones_locations=_ZONES_2, zones_colours=(6,) * 3 + (5,) * 12),
})
# the zone-goals registrations of the same ids (zone-goals/envs/__init__.py:52-81, :111-119; TSPHardEnv over
# TSPNextCityEnv, zone-goals/envs/TSP_hard_env.py): goal-conditioned, v5 with 300 steps
ENV_SPECS.update({
    'zone-goals/PointTSP-v4': TaskSpec(_lib.TASK_TSP, 15, 1000, 6, goals=True, robot_location=(-0.9, -0.9),
                                       robot_rot=-1.0, zones_locations=_ZONES_1,
                                       zones_colours=(6,) * 5 + (5,) * 10),
    'zone-goals/PointTSP-v5': TaskSpec(_lib.TASK_TSP, 15, 300, 6, goals=True, robot_location=(0.8, 0.8),
                                       zones_locations=_ZONES_2, zones_colours=(6,) * 3 + (5,) * 12),
})
# `walled=True` (main/envs/zone_envs/ZoneEnvBase.py:39,55-62): no registration of the reference sets it
# (main/envs/__init__.py:7-50), so there is no gym id; 'walled/<id>' names the env the reference's class builds
# from <id>'s config with walled=True (tests/golden/gen_golden_walls.py builds exactly that).
import dataclasses as _dc
ENV_SPECS.update({'walled/' + k: _dc.replace(ENV_SPECS[k], walled=True)
                  for k in ('PointTSP-v0', 'PointTSP-v1', 'PointTTSP-v0', 'PointTTSP-v1', 'ColourMatch-v0')})
