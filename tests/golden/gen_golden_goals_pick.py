"""The high-level goal chooser used when the goal fixtures were recorded (shared by
gen_golden_goals.py and the tests; imports nothing from the reference)."""
import numpy as np


def pick_goal(obs, available, rs, mode):
    idx = np.flatnonzero(available)
    if mode == 'random' or len(idx) == 0:
        return int(rs.choice(idx)) if len(idx) else 0
    pos = obs['obs'][1:3] * 3.0
    d = np.linalg.norm(obs['zone_obs'][:, 0:2] * 3.0 - pos, axis=1)
    if mode == 'far':                       # walks past other zones: wrong-zone visits happen
        return int(idx[np.argmax(d[idx])])
    return int(idx[np.argmin(d[idx])])
