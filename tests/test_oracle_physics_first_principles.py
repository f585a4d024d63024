"""Independent check of the physics oracle's equations of motion (oracle/mj_point.py).

MuJoCo itself cannot be run here (physics parity is unpinned), so what CAN be
checked is checked from first principles: the joint-space inertia, the centrifugal bias and the
generalized actuator forces that the oracle restates from recall are re-derived symbolically --
Euler-Lagrange equations of the two-geom rigid body of point.xml (solid sphere r = 0.1 at the
body origin, solid box of half-size 0.05 at (0.1, 0, 0) in the body frame, density 1, slides x, y
along fixed axes followed by a hinge about z through the body origin) and the principle of
virtual work for the site motor and the hinge servo -- and compared with the oracle's functions
at random states.  What remains recalled and unpinned after this test: the XML constants
themselves, MuJoCo's semi-implicit Euler with implicit joint damping, and the exclusion of the
sphere/floor contact at distance == margin."""
import math

import numpy as np
import pytest

sp = pytest.importorskip('sympy')

from oracle import mj_point as mj  # noqa: E402


def derive():
    t = sp.symbols('t')
    x, y, th = (sp.Function(n)(t) for n in ('x', 'y', 'th'))
    rho, r, a, bx = sp.Integer(1), sp.Rational(1, 10), sp.Rational(1, 20), sp.Rational(1, 10)
    m_s = sp.Rational(4, 3) * sp.pi * r ** 3 * rho                       # solid sphere
    i_s = sp.Rational(2, 5) * m_s * r ** 2                               # about any axis through its centre
    m_b = (2 * a) ** 3 * rho                                             # solid cube, side 2a
    i_b = m_b * ((2 * a) ** 2 + (2 * a) ** 2) / 12                       # about z through its centre
    # positions of the two centres of mass in the frame of the (fixed) slide axes
    ps = sp.Matrix([x, y])
    pb = sp.Matrix([x + bx * sp.cos(th), y + bx * sp.sin(th)])
    vs, vb = ps.diff(t), pb.diff(t)
    w = th.diff(t)
    T = (m_s * vs.dot(vs) + i_s * w ** 2 + m_b * vb.dot(vb) + i_b * w ** 2) / 2   # planar motion, no potential term
    q = [x, y, th]
    qd = [v.diff(t) for v in q]
    qdd = [v.diff(t, 2) for v in q]
    eqs = [sp.expand(sp.diff(sp.diff(T, qd[i]), t) - sp.diff(T, q[i])) for i in range(3)]
    M = sp.Matrix(3, 3, lambda i, j: sp.simplify(sp.diff(eqs[i], qdd[j])))
    bias = sp.Matrix([sp.simplify(eqs[i] - sum(M[i, j] * qdd[j] for j in range(3))) for i in range(3)])
    # generalized forces by virtual work: force F along the body x axis applied AT the body origin
    # (the site), torque tau about z
    F, tau = sp.symbols('F tau')
    Jo = sp.Matrix(2, 3, lambda i, j: sp.diff(ps[i], q[j]))             # Jacobian of the site
    Q = Jo.T * sp.Matrix([F * sp.cos(th), F * sp.sin(th)]) + sp.Matrix([0, 0, tau])
    TH, W = sp.symbols('TH W')
    plain = lambda e: sp.simplify(e.subs(w, W).subs(th, TH))             # functions of t -> plain symbols
    assert all(not plain(e).has(x, y) and not plain(e).has(sp.Derivative) for e in list(M) + list(bias) + list(Q))
    return (sp.lambdify((TH, W), [plain(e) for e in M], 'math'), sp.lambdify((TH, W), [plain(e) for e in bias], 'math'),
            sp.lambdify((TH, F, tau), [plain(e) for e in Q], 'math'), float(m_s + m_b))


def test_equations_of_motion_from_first_principles():
    M_f, b_f, Q_f, mass = derive()
    assert abs(mass - mj.MASS) < 1e-15
    rs = np.random.RandomState(0)
    for _ in range(200):
        th, w = rs.uniform(-40, 40), rs.uniform(-6, 6)
        M = np.array(M_f(th, w), dtype=np.float64).reshape(3, 3)
        assert np.allclose(M, mj.mass_matrix(th), rtol=1e-12, atol=1e-18), (th, M, mj.mass_matrix(th))
        b = np.array(b_f(th, w), dtype=np.float64).reshape(3)
        assert np.allclose(b, mj.bias_force(th, np.array([0.3, -0.2, w])), rtol=1e-12, atol=1e-18)
        # actuator map: motor force f (already clamped) through gear 0.3 at the site, servo torque through gear 0.3
        u0, u1 = rs.uniform(-1.5, 1.5), rs.uniform(-1.5, 1.5)
        f_motor = min(max(min(max(u0, -1), 1), -0.05), 0.05)
        f_servo = min(max(min(max(u1, -1), 1) - 0.3 * w, -0.05), 0.05)
        Q = np.array(Q_f(th, 0.3 * f_motor, 0.3 * f_servo), dtype=np.float64).reshape(3)
        assert np.allclose(Q, mj.actuator_force(th, np.array([0.0, 0.0, w]), np.array([u0, u1])), rtol=1e-12, atol=1e-18)


def test_substep_is_the_damped_semi_implicit_euler_of_those_equations():
    """(M + h B) a = Q - bias - B v;  v' = v + h a;  q' = q + h v'  -- assembled here from the
    symbolic pieces and compared with oracle.substep."""
    M_f, b_f, Q_f, _ = derive()
    rs = np.random.RandomState(1)
    h, B = 0.002, np.diag([0.01, 0.01, 0.005])
    for _ in range(200):
        q = np.array([rs.uniform(-2, 2), rs.uniform(-2, 2), rs.uniform(-10, 10)])
        v = np.array([rs.uniform(-1.5, 1.5), rs.uniform(-1.5, 1.5), rs.uniform(-5, 5)])
        u = rs.uniform(-1.2, 1.2, 2) * (0.05 if rs.rand() < 0.3 else 1.0)
        f_motor = min(max(min(max(u[0], -1), 1), -0.05), 0.05)
        f_servo = min(max(min(max(u[1], -1), 1) - 0.3 * v[2], -0.05), 0.05)
        M = np.array(M_f(q[2], v[2]), dtype=np.float64).reshape(3, 3)
        rhs = np.array(Q_f(q[2], 0.3 * f_motor, 0.3 * f_servo), dtype=np.float64).reshape(3) \
            - np.array(b_f(q[2], v[2]), dtype=np.float64).reshape(3) - B @ v
        a = np.linalg.solve(M + h * B, rhs)
        v1 = v + h * a
        q1 = q + h * v1
        q_o, v_o = mj.substep(q, v, u)
        assert np.allclose(q1, q_o, rtol=1e-12, atol=1e-15) and np.allclose(v1, v_o, rtol=1e-11, atol=1e-15)
    # closed forms: terminal speed 0.3 * 0.05 / 0.01 = 1.5 m/s, terminal yaw rate from 0.3 (u - 0.3 w) = 0.005 w
    assert math.isclose(0.3 * 0.05 / 0.01, 1.5)
