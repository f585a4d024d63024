"""What could the one unpinned physics question change?  

The sphere/floor contact of point.xml sits at signed distance exactly 0.0.  The canonical reading (MuJoCo lists the
contact and EXCLUDES it: dist < includemargin is false) gives no constraint force; it cannot be checked against
MuJoCo here.  The term is isolated -- oracle/mj_point.py::constraint_force and its kernel-side twin
crl_core.cuh::constraint_acc<MODEL> (compile-time, -DCRL_CONTACT_MODEL) -- and this file measures what the
ALTERNATIVE reading (contact active, pyramidal friction cone) would do:

* the two twins agree per substep under BOTH readings (so a correction is a one-function change on each side);
* under the alternative the robot's terminal speed at full throttle is ~3 mm/s instead of 1.5 m/s -- and 1.5 is
  the reference's OWN velocity normaliser (``robot_velp / 1.5``, ZoneEnvBase.py:223): the reference's authors
  normalised by the no-contact model's terminal speed 0.3 * 0.05 / 0.01;
* replaying the recorded action sequences of the fixture episodes open loop, the alternative leaves the robot
  within centimetres of its start: no zone would ever be visited and every return would be 0.

So the risk is not a few percent of drift: the alternative describes a robot that does not move, which contradicts the
task itself (README gifs, published returns).  The numbers are also written by tools/contact_hypothesis.py to
profiles/r02_contact_hypothesis.json.
"""
import ctypes
import glob
import math
import os
import subprocess

import numpy as np
import pytest

from oracle import mj_point as mj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, 'tests', 'hostcheck')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='module')
def hc():
    so, src = os.path.join(HC, 'libhostcheck.so'), os.path.join(HC, 'hostcheck.cpp')
    core = os.path.join(ROOT, 'combinatorial_rl_tasks_b200', 'csrc', 'crl_core.cuh')
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.run(['g++', '-O2', '-x', 'c++', '-shared', '-fPIC', '-o', so, src], check=True)
    L = ctypes.CDLL(so)
    fp = ctypes.POINTER(ctypes.c_float)
    for f in (L.hc_substeps, L.hc_substeps_contact_active):
        f.argtypes = [fp, ctypes.c_float, ctypes.c_float, ctypes.c_int, fp]
    return L


@pytest.fixture
def contact_active():
    mj.CONTACT_MODEL = 'active'
    yield
    mj.CONTACT_MODEL = 'excluded'


def rollout(actions, model, frameskip=10):
    """Open-loop replay from rest at the origin with the oracle's substep under one contact reading."""
    old = mj.CONTACT_MODEL
    mj.CONTACT_MODEL = model
    try:
        q, v = np.zeros(3), np.zeros(3)
        path = [q[:2].copy()]
        for a in actions:
            ctrl = np.clip(a, -1, 1)
            for _ in range(frameskip):
                q, v = mj.substep(q, v, ctrl)
            path.append(q[:2].copy())
        return np.array(path), v
    finally:
        mj.CONTACT_MODEL = old


def test_kernel_twin_matches_oracle_under_both_readings(hc, contact_active):
    """constraint_acc<1> (crl_core.cuh, as the kernel would compile it with -DCRL_CONTACT_MODEL=1) against the
    oracle's 'active' hook, per substep from identical inputs; the canonical pair is tests/test_host_logic.py's."""
    rs = np.random.RandomState(1)
    fp = ctypes.POINTER(ctypes.c_float)
    cs = (ctypes.c_float * 2)()
    worst = 0.0
    for _ in range(500):
        st = np.array([rs.uniform(-2.5, 2.5), rs.uniform(-2.5, 2.5), rs.uniform(-np.pi, np.pi),
                       rs.uniform(-0.02, 0.02), rs.uniform(-0.02, 0.02), rs.uniform(-0.05, 0.05)], dtype=np.float32)
        a = rs.uniform(-1.3, 1.3, 2).astype(np.float32)
        w = st.astype(np.float64)
        q_ref, v_ref = mj.substep(w[:3], w[3:], a.astype(np.float64))
        out = st.copy()
        hc.hc_substeps_contact_active(out.ctypes.data_as(fp), float(a[0]), float(a[1]), 1, cs)
        ref = np.concatenate([q_ref, v_ref])
        worst = max(worst, float(np.max(np.abs(out - ref) / np.maximum(1.0, np.abs(ref)))))
    # the oracle solves the coupled 3x3 system, the kernel twin applies the (decoupled) acceleration form: the
    # off-diagonal coupling k/m ~ 2 % enters only the tiny constraint-induced accelerations
    assert worst <= 1e-4, worst


def test_terminal_speed_separates_the_two_readings():
    full = np.tile(np.array([1.0, 0.0]), (300, 1))                    # 6 s of full throttle, straight ahead
    _, v_free = rollout(full, 'excluded')
    _, v_drag = rollout(full, 'active')
    # no contact force: 0.3 * 0.05 / 0.01 = 1.5 m/s, which is exactly what ZoneEnvBase.py:223 divides robot_velp by
    assert abs(np.hypot(v_free[0], v_free[1]) - 1.5) < 2e-3
    # contact active: the drag b ~ 105 /s holds the robot at millimetres per second
    assert np.hypot(v_drag[0], v_drag[1]) < 5e-3


def test_fixture_action_sequences_under_the_alternative_reading():
    """How far the trajectories of the recorded episodes move between the two readings."""
    report = {}
    for path in sorted(glob.glob(os.path.join(GOLDEN, 'PointTSP_10000*_*.npz')))[:3]:
        g = dict(np.load(path))
        acts = g['actions'][:400]
        p_free, _ = rollout(acts, 'excluded')
        p_drag, _ = rollout(acts, 'active')
        reach_free = float(np.max(np.linalg.norm(p_free, axis=1)))
        reach_drag = float(np.max(np.linalg.norm(p_drag, axis=1)))
        report[os.path.basename(path)] = (reach_free, reach_drag)
        # the recorded episode itself (qpos of the fixture) is the canonical reading (to the servo chatter's
        # amplification of the last bit of the recorded actions)
        assert np.allclose(p_free[:10], g['qpos'][:10, :2], atol=1e-6)
        assert reach_drag < 0.05, report          # never leaves a 5 cm disc: zones are 0.2 m wide and >= 0.95 m away
        if 'idle' not in path:
            assert reach_free > 0.5, report
    assert report
