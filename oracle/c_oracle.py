"""ctypes face of oracle/crl_oracle.c (TEST INFRASTRUCTURE ONLY; see its header)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libcrl_oracle.so')
TASKS = {'PointTSP-v0': (0, 15, 2000), 'PointTTSP-v0': (1, 15, 2000), 'ColourMatch-v0': (2, 6, 2000),
         'PointTSP-v1': (0, 5, 1000), 'PointTTSP-v1': (1, 5, 1000),
         'PointTSP-v4': (0, 15, 1000), 'PointTSP-v5': (0, 15, 250)}
_lib = None
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int64)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, 'crl_oracle.c')
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
            subprocess.run(['make', '-s', '-C', HERE], check=True)
        L = ctypes.CDLL(LIB)
        L.oe_new.restype = ctypes.c_void_p
        L.oe_new.argtypes = [ctypes.c_int] * 3
        L.oe_free.argtypes = [ctypes.c_void_p]
        L.oe_seed.argtypes = [ctypes.c_void_p, ctypes.c_int64]
        L.oe_get_seed.restype = ctypes.c_int64
        L.oe_get_seed.argtypes = [ctypes.c_void_p]
        L.oe_reset.argtypes = [ctypes.c_void_p]
        L.oe_reset_layout.argtypes = [ctypes.c_void_p, dp, ctypes.c_double, dp, ip, ip]
        L.oe_set_state.argtypes = [ctypes.c_void_p, dp, dp]
        L.oe_get_state.argtypes = [ctypes.c_void_p, dp, dp]
        L.oe_get_layout.argtypes = [ctypes.c_void_p, dp, dp, dp, ip, ip]
        L.oe_get_task_state.argtypes = [ctypes.c_void_p, ip]
        L.oe_step.restype = ctypes.c_int
        L.oe_step.argtypes = [ctypes.c_void_p, dp, dp, ctypes.POINTER(ctypes.c_int)]
        L.oe_event.restype = ctypes.c_int
        L.oe_event.argtypes = [ctypes.c_void_p]
        L.oe_obs.argtypes = [ctypes.c_void_p, dp, dp]
        L.oe_substep.argtypes = [dp, dp, dp]
        L.oe_timed_rollout.restype = ctypes.c_double
        L.oe_timed_rollout.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ip]
        L.ph_reset.argtypes = ([ctypes.c_int] * 4 + [ctypes.c_int64] * 3 + [ctypes.c_uint32, ctypes.c_int64,
                               ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_float, ctypes.c_float]
                               + [ctypes.c_void_p] * 7)
        L.ph_action.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_void_p]
        L.ph_philox.argtypes = [ctypes.c_void_p] * 3
        L.ph_gamma_sample.restype = ctypes.c_double
        L.ph_gamma_sample.argtypes = [ctypes.c_int64, ctypes.c_double, ctypes.c_uint32, ctypes.c_uint32]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(dp)


def _i(a):
    return a.ctypes.data_as(ip)


class CEnv:
    """Same surface as oracle.zone_env.ZoneTaskEnv, backed by the C twin."""

    def __init__(self, env_id):
        self.task, self.N, self.num_steps = TASKS[env_id]
        self.Z = 6 if self.task == 0 else 7
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.oe_new(self.task, self.N, self.num_steps))

    def __del__(self):
        if getattr(self, 'h', None):
            self.L.oe_free(self.h)
            self.h = None

    def seed(self, s):
        self.L.oe_seed(self.h, int(s))

    def reset(self, layout=None):
        if layout is None:
            self.L.oe_reset(self.h)
        else:
            xy0 = np.ascontiguousarray(layout['xy0'], dtype=np.float64)
            zxy = np.ascontiguousarray(layout['zone_xy'], dtype=np.float64)
            tm = np.ascontiguousarray(layout.get('zone_max_steps', np.zeros(self.N)), dtype=np.int64)
            col = np.ascontiguousarray(layout.get('colours', np.zeros(self.N)), dtype=np.int64)
            self.L.oe_reset_layout(self.h, _d(xy0), float(layout['rot0']), _d(zxy), _i(tm), _i(col))
        return self.obs()

    def obs(self):
        o, z = np.zeros(8), np.zeros((self.N, self.Z))
        self.L.oe_obs(self.h, _d(o), _d(z))
        return {'obs': o, 'zone_obs': z}

    def step(self, action):
        a = np.ascontiguousarray(action, dtype=np.float64)
        r, g = ctypes.c_double(), ctypes.c_int()
        d = self.L.oe_step(self.h, _d(a), ctypes.byref(r), ctypes.byref(g))
        info = {'cost': 0}
        if g.value:
            info['goal_met'] = True
        return self.obs(), r.value, bool(d), info

    @property
    def event(self):
        return self.L.oe_event(self.h)

    def set_state(self, qpos, qvel):
        qp, qv = np.ascontiguousarray(qpos, dtype=np.float64), np.ascontiguousarray(qvel, dtype=np.float64)
        self.L.oe_set_state(self.h, _d(qp), _d(qv))

    def get_state(self):
        qp, qv = np.zeros(3), np.zeros(3)
        self.L.oe_get_state(self.h, _d(qp), _d(qv))
        return qp, qv

    def layout(self):
        xy0, rot0, zxy = np.zeros(2), ctypes.c_double(), np.zeros((self.N, 2))
        tm, col = np.zeros(self.N, dtype=np.int64), np.zeros(self.N, dtype=np.int64)
        self.L.oe_get_layout(self.h, _d(xy0), ctypes.byref(rot0), _d(zxy), _i(tm), _i(col))
        out = {'xy0': xy0, 'rot0': rot0.value, 'zone_xy': zxy}
        if self.task == 1:
            out['zone_max_steps'] = tm
        if self.task == 2:
            out['colours'] = col
        return out

    def task_state(self):
        s = np.zeros(4 + 3 * self.N, dtype=np.int64)
        self.L.oe_get_task_state(self.h, _i(s))
        N = self.N
        return {'steps': int(s[0]), 'done': bool(s[1]), 'event': int(s[2]), 'goal_dist': int(s[3]),
                'visited': s[4:4 + N].astype(bool), 'colours': s[4 + N:4 + 2 * N].copy(),
                'cooldown': s[4 + 2 * N:4 + 3 * N].copy()}


def substep(qpos, qvel, ctrl):
    qp, qv = np.array(qpos, dtype=np.float64), np.array(qvel, dtype=np.float64)
    c = np.ascontiguousarray(ctrl, dtype=np.float64)
    lib().oe_substep(_d(qp), _d(qv), _d(c))
    return qp, qv


def timed_rollout(env_id, threads, seconds):
    task, N, num_steps = TASKS[env_id]
    total = ctypes.c_int64()
    import time
    t0 = time.perf_counter()
    rate = lib().oe_timed_rollout(task, N, num_steps, threads, seconds, ctypes.byref(total))
    return rate, total.value, time.perf_counter() - t0


def philox_reset(env_id, seed_in, seed_mode=0, min_seed=0, max_seed=0, global_env=0, episode=0,
                 beta=(3.0, 1.5), keepouts=(0.4, 0.55), extent=3.0, fixed=None):
    """Design twin of the device reset (crl_reset / auto-reset) for one env.  ``fixed``: the
    float32 (1 + N, 4) table of CrlState.fixed_layout (hard instances), or None."""
    task, N, num_steps = TASKS[env_id]
    if fixed is not None:
        fixed = np.ascontiguousarray(fixed, dtype=np.float32).reshape(1 + N, 4)
    xy0, rot0, zxy = np.zeros(2, np.float32), np.zeros(1, np.float32), np.zeros((N, 2), np.float32)
    tm, col, after = np.zeros(N, np.int32), np.zeros(N, np.int32), np.zeros(1, np.int64)
    lib().ph_reset(task, N, num_steps, seed_mode, min_seed, max_seed, global_env, episode, seed_in,
                   beta[0], beta[1], keepouts[0], keepouts[1], extent,
                   None if fixed is None else fixed.ctypes.data, xy0.ctypes.data, rot0.ctypes.data, zxy.ctypes.data, tm.ctypes.data, col.ctypes.data,
                   after.ctypes.data)
    return {'xy0': xy0, 'rot0': float(rot0[0]), 'zone_xy': zxy, 'zone_max_steps': tm, 'colours': col,
            'seed_after': int(after[0])}


def philox_action(action_seed, global_env, step_index):
    a = np.zeros(2, np.float32)
    lib().ph_action(action_seed, global_env, step_index, a.ctypes.data)
    return a


def philox_gamma(seed, shape, zone=0, which=0):
    """One gamma(shape) draw of the device's task stream (design twin)."""
    return lib().ph_gamma_sample(seed, shape, zone, which)
