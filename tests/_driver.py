"""Scripted drivers that make the task logic fire (visits, timeouts, colour cycles).
Not part of the path: used to record the golden fixtures and to drive closed-loop
parity runs.  Works on any obs dict with 'obs' (8,) and 'zone_obs' (N,Z) arrays."""
import math

import numpy as np


def steer(obs, target_xy, rs, noise):
    pos = obs['obs'][1:3] * 3.0
    heading = math.atan2(obs['obs'][4], obs['obs'][3])
    want = math.atan2(target_xy[1] - pos[1], target_xy[0] - pos[0])
    err = (want - heading + math.pi) % (2 * math.pi) - math.pi
    dist = math.hypot(target_xy[1] - pos[1], target_xy[0] - pos[0])
    speed = math.hypot(obs['obs'][5], obs['obs'][6]) * 1.5
    # the motor saturates at |a0| >= 0.05 (forcerange), so thrust is modulated below that
    if abs(err) < 0.35:
        a0 = 0.05 if dist > 0.6 or speed < 0.5 else 0.0
    else:
        a0 = -0.02 if speed > 0.3 else 0.0
    a = np.array([a0, np.clip(1.5 * err, -1, 1)])
    a = a + noise * rs.uniform(-1, 1, 2) * np.array([0.02, 0.2])
    return a.astype(np.float32)


def policy(env_id, obs, rs, mode, t):
    if mode == 'random':
        return rs.uniform(-1.5, 1.5, 2).astype(np.float32)   # also exercises the ctrl clip
    z = obs['zone_obs']
    pos = obs['obs'][1:3] * 3.0
    xy = z[:, 0:2] * 3.0
    d = np.linalg.norm(xy - pos, axis=1)
    if env_id == 'ColourMatch-v0':
        col = np.argmax(z[:, 2:5][:, ::-1], axis=1)       # rgb -> 0 Blue, 1 Green, 2 Red
        counts = [np.sum(col == c) for c in range(3)]
        goal = int(np.argmax(counts))
        cand = [i for i in range(len(z)) if col[i] != goal and z[i, 6] == 0]
        if not cand:
            return steer(obs, (0.0, 0.0), rs, 0.3)
        tgt = min(cand, key=lambda i: d[i])
    else:
        cand = [i for i in range(len(z)) if z[i, 2] == 0]    # cyan = unvisited
        if not cand:
            return np.zeros(2, dtype=np.float32)
        if env_id == 'PointTTSP-v0' and mode == 'deadline':
            tgt = min(cand, key=lambda i: z[i, 6])
        elif mode == 'idle' and t > 150:
            return np.zeros(2, dtype=np.float32)
        else:
            tgt = min(cand, key=lambda i: d[i])
    return steer(obs, xy[tgt], rs, 0.1)
