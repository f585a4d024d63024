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
