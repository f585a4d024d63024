"""`walled=True` (main/envs/zone_envs/ZoneEnvBase.py:39,55-62).  CPU part: the wall list of the
fixtures recorded from the REAL ZoneEnvBase, the oracle's restatement of it, the kernel-side contact (crl_core.cuh
wall_force, g++ build) against the oracle's generic solver.  PARITY UNPINNED for the contact itself (simplified MuJoCo
soft contact, see oracle/mj_point.py); what is pinned is the reference-owned plumbing."""
import ctypes
import glob
import os
import subprocess

import numpy as np
import pytest

from oracle import mj_point as mj
from oracle import zone_env as ze

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
WALLS = sorted(glob.glob(os.path.join(GOLDEN, 'walls_*.npz')))
WALL_RTOL = 2e-5        # per substep in contact, fp32 closed form vs fp64 generic solver (measured 8e-6)


def test_wall_list_of_the_real_class():
    assert len(WALLS) == 3
    for path in WALLS:
        g = dict(np.load(path))
        w = g['walls_locations']
        assert w.shape == (244, 2) and float(g['walls_size']) == ze.WALLS_SIZE == 0.1
        assert np.array_equal(w, np.array(ze.wall_locations(3), dtype=np.float64))        # the oracle's restatement
        assert np.max(np.abs(g['wall_layout'] - w)) <= 1e-9                             # placed within 1e-9 of the list
        # the square of half-width 3, every multiple of 0.1, the four corners twice
        assert np.all(np.abs(w).max(axis=1) == 3.0)
        assert len({(round(x, 6), round(y, 6)) for x, y in w}) == 240
        # the robot does reach the walls in these episodes and never gets through them
        pos = g['obs'][:, 1:3] * 3.0
        assert (np.abs(pos).max(axis=1) > 2.8).sum() > 100 and np.abs(pos).max() < 2.82


def test_config_carries_the_flag():
    from combinatorial_rl_tasks_b200.config import ENV_SPECS
    for k in ('PointTSP-v0', 'PointTTSP-v0', 'ColourMatch-v0'):
        assert ENV_SPECS['walled/' + k].walled and not ENV_SPECS[k].walled
        assert ENV_SPECS['walled/' + k].num_zones == ENV_SPECS[k].num_zones


@pytest.fixture(scope='module')
def hc():
    hcdir = os.path.join(ROOT, 'tests', 'hostcheck')
    so, src = os.path.join(hcdir, 'libhostcheck.so'), os.path.join(hcdir, 'hostcheck.cpp')
    core = os.path.join(ROOT, 'combinatorial_rl_tasks_b200', 'csrc', 'crl_core.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(['g++', '-O2', '-x', 'c++', '-shared', '-fPIC', '-o', so, src], check=True)
    L = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    L.hc_substeps_walled.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
    L.hc_substeps.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
    return L


def oracle_walls():
    return dict(p0=np.zeros(2), rot0=0.0, boxes=np.array(ze.wall_locations(3), dtype=np.float64), half=0.1)


def test_kernel_side_contact_against_the_oracle(hc):
    """One substep from identical states inside the contact band of one wall or of a corner: the closed form with
    merged rows (crl_core.cuh) vs the oracle's generic contact enumeration + projected Gauss-Seidel."""
    walls, rs = oracle_walls(), np.random.RandomState(0)
    fp, cs = ctypes.POINTER(ctypes.c_float), (ctypes.c_float * 2)()
    worst, touching = 0.0, 0
    for trial in range(1500):
        pen = rs.uniform(-0.002, 0.012)
        along = rs.uniform(-2.81, 2.81) if trial % 4 else rs.choice([-1, 1]) * (2.8 + rs.uniform(-0.002, 0.01))
        X, Y = [(2.8 + pen, along), (-2.8 - pen, along), (along, 2.8 + pen), (along, -2.8 - pen)][rs.randint(4)]
        st = np.array([X, Y, rs.uniform(-np.pi, np.pi), rs.uniform(-1.5, 1.5), rs.uniform(-1.5, 1.5), rs.uniform(-4, 4)],
                      dtype=np.float32)
        a = rs.uniform(-1.2, 1.2, 2).astype(np.float32)
        w = st.astype(np.float64)
        touching += bool(np.any(mj.wall_force(w[:3], w[3:], np.zeros(3), walls) != 0))
        q, v = mj.substep(w[:3], w[3:], a.astype(np.float64), walls=walls)
        out = st.copy()
        hc.hc_substeps_walled(out.ctypes.data_as(fp), float(a[0]), float(a[1]), 1, cs)
        ref = np.concatenate([q, v])
        worst = max(worst, float(np.max(np.abs(out - ref) / np.maximum(1.0, np.abs(ref)))))
    assert touching > 500
    assert worst <= WALL_RTOL, worst


def test_walls_change_nothing_away_from_them(hc):
    rs = np.random.RandomState(1)
    fp, cs = ctypes.POINTER(ctypes.c_float), (ctypes.c_float * 2)()
    for _ in range(300):
        st = np.array([rs.uniform(-2.79, 2.79), rs.uniform(-2.79, 2.79), rs.uniform(-3, 3), rs.uniform(-1.5, 1.5),
                       rs.uniform(-1.5, 1.5), rs.uniform(-4, 4)], dtype=np.float32)
        a, b = st.copy(), st.copy()
        hc.hc_substeps_walled(a.ctypes.data_as(fp), 0.7, -0.2, 1, cs)
        hc.hc_substeps(b.ctypes.data_as(fp), 0.7, -0.2, 1, cs)
        assert np.array_equal(a, b)


def test_the_wall_holds_and_dissipates():
    """Full throttle into a wall (oracle): penetration stays below 1.2 cm (critically damped, stiffness d / (dmax tc)^2),
    the robot ends up at rest against the inner face, and never gains speed from the contact."""
    walls = oracle_walls()
    q, v = np.array([2.0, 0.3, 0.0]), np.array([1.5, 0.0, 0.0])
    deepest, fastest = 0.0, 0.0
    for i in range(3000):
        q, v = mj.substep(q, v, np.array([1.0, 0.0]), walls=walls)
        deepest = max(deepest, q[0] - 2.8)
        fastest = max(fastest, abs(v[0]))
    assert 0.0 < deepest < 0.012, deepest
    assert fastest <= 1.5 + 1e-9
    assert abs(v[0]) < 1e-3 and 2.8 < q[0] < 2.8 + 2e-3        # pressed into the wall by the motor's 0.015 N
