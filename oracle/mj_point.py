"""CPU oracle, part 1: the MuJoCo 2.0 step for Safety Gym's ``xmls/point.xml``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it.

PARITY UNPINNED (physics).  The arithmetic restated here lives in third-party
code that is *not* under /root/reference and cannot be installed in this image:

* ``mujoco-py==2.0.2.9`` (reference ``requirements.txt:3``) wrapping the closed
  MuJoCo 2.0 binary -- ``mj_step`` / ``mj_forward``;
* ``safety-gym`` (un-pinned sibling checkout, reference ``README.md:36-37``,
  ``main/setup.sh:2``) -- ``safety_gym/xmls/point.xml`` and ``World.build``.

The reference's call sites into it are ``main/envs/zone_envs/ZoneEnvBase.py:94``
(``data.get_body_xpos``), ``:145`` (``sim.forward``), ``:220-224``
(``world.robot_pos / robot_vel``, ``get_body_xquat``, ``get_body_xvelr``) and
``Engine.step`` (``sim.step()`` x frameskip) reached from
``main/envs/TSP_env.py:49``.  What follows restates the published MuJoCo
algorithm for this one model:

model (point.xml): timestep 0.002; body ``robot`` at (x0, y0, 0.1) with quat
(cos(rot/2),0,0,sin(rot/2)); joints in order slide-x (damping .01), slide-y
(.01), hinge-z (.005); geoms sphere r=.1 and box half-size .05 at (.1,0,0),
density 1; actuators ``motor gear=.3 site=robot`` and ``velocity gear=.3 kv=1``
on the hinge, both ctrlrange +-1 and forcerange +-.05.

mj_step = mj_forward (position, velocity, actuation, acceleration, constraint
stages on the *current* state) followed by mj_Euler with implicit joint damping:
solve (M + h diag(B)) a = qfrc_smooth + qfrc_constraint, v += h a, q += h v.
The sphere/floor contact has signed distance exactly 0.0 == margin, which
MuJoCo lists but excludes (``dist < includemargin`` is false), so
``constraint_force`` returns zero; it is kept as an isolated hook.

tests/test_oracle_physics_first_principles.py re-derives the equations of motion used below
(mass matrix, centrifugal bias, generalized actuator forces, the damped semi-implicit Euler
step) symbolically from the model geometry and checks this file against them to 1e-12; that
removes transcription and derivation errors, not the dependence on the recalled model
constants, integrator choice and contact exclusion.

The mass matrix is assembled generically (composite rigid body of the two
geoms) and the 3x3 system is solved with ``numpy.linalg.solve`` on purpose: the
CUDA kernel uses an independently derived closed form, so agreement between the
two is a real check and not the same formula typed twice.
"""
import math

import numpy as np

# ---- model constants derived from point.xml (density 1) -------------------
TIMESTEP = 0.002
SPHERE_R = 0.1
BOX_HALF = 0.05
BOX_X = 0.1
M_SPHERE = 4.0 / 3.0 * math.pi * SPHERE_R ** 3
M_BOX = 8.0 * BOX_HALF ** 3
MASS = M_SPHERE + M_BOX
# centre of mass offset along body-x and inertia about the hinge (body z) axis
COM_X = M_BOX * BOX_X / MASS
I_HINGE = (0.4 * M_SPHERE * SPHERE_R ** 2
           + M_BOX / 3.0 * (BOX_HALF ** 2 + BOX_HALF ** 2)
           + M_BOX * BOX_X ** 2)
DAMPING = np.array([0.01, 0.01, 0.005])
GEAR_MOTOR = 0.3
GEAR_VELOCITY = 0.3
KV = 1.0
FORCE_LIMIT = 0.05
CTRL_LIMIT = 1.0
BODY_Z = 0.1


def quat_rotate(quat, vec):
    """v' = q v q*  (the two-product form MuJoCo's mju_rotVecQuat uses)."""
    w, x, y, z = quat
    t0 = -x * vec[0] - y * vec[1] - z * vec[2]
    t1 = w * vec[0] + y * vec[2] - z * vec[1]
    t2 = w * vec[1] + z * vec[0] - x * vec[2]
    t3 = w * vec[2] + x * vec[1] - y * vec[0]
    return np.array([
        -t0 * x + t1 * w - t2 * z + t3 * y,
        -t0 * y + t2 * w - t3 * x + t1 * z,
        -t0 * z + t3 * w - t1 * y + t2 * x,
    ])


def quat_mul(a, b):
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0],
    ])


def mass_matrix(theta):
    """Joint-space inertia of the sphere+box body for q = (x, y, theta).

    x, y are measured along the *initial* body axes (the slides precede the
    hinge in the kinematic chain, so they do not rotate with theta).
    """
    s, c = math.sin(theta), math.cos(theta)
    mc = MASS * COM_X
    return np.array([[MASS, 0.0, -mc * s],
                     [0.0, MASS, mc * c],
                     [-mc * s, mc * c, I_HINGE]])


def bias_force(theta, qvel):
    """qfrc_bias: centrifugal term of the off-axis COM (gravity has no
    generalized component: there is no z DOF and the hinge is vertical)."""
    s, c = math.sin(theta), math.cos(theta)
    mc = MASS * COM_X
    w2 = qvel[2] * qvel[2]
    return np.array([-mc * w2 * c, -mc * w2 * s, 0.0])


def actuator_force(theta, qvel, ctrl):
    """qfrc_actuator for the site motor and the hinge velocity servo."""
    u0 = min(max(ctrl[0], -CTRL_LIMIT), CTRL_LIMIT)
    u1 = min(max(ctrl[1], -CTRL_LIMIT), CTRL_LIMIT)
    f_motor = min(max(u0, -FORCE_LIMIT), FORCE_LIMIT)
    f_servo = KV * u1 - KV * (GEAR_VELOCITY * qvel[2])
    f_servo = min(max(f_servo, -FORCE_LIMIT), FORCE_LIMIT)
    s, c = math.sin(theta), math.cos(theta)
    # site wrench: GEAR_MOTOR * f along body-x at the hinge axis (no yaw arm)
    return np.array([GEAR_MOTOR * f_motor * c,
                     GEAR_MOTOR * f_motor * s,
                     GEAR_VELOCITY * f_servo])


# Which reading of the sphere/floor contact the hook below implements.  'excluded' is the canonical
# one; 'active' exists to put a NUMBER on what the unpinned question could change (tests/test_contact_hypothesis.py)
# and has its compile-time twin in csrc/crl_core.cuh (CRL_CONTACT_MODEL 1).
CONTACT_MODEL = 'excluded'
# MuJoCo defaults the active reading would run with: solimp d0 = 0.9 at zero penetration, solref (timeconst 0.02,
# dampratio 1) -> velocity gain b = 2 / (dmax * timeconst) with dmax = 0.95
CONTACT_IMPEDANCE = 0.9
CONTACT_B = 2.0 / (0.95 * 0.02)


def constraint_force(qpos, qvel, qfrc_smooth, h=TIMESTEP):
    """qfrc_constraint, the ONE isolated place where a contact enters the step.

    'excluded' (canonical): the sphere/floor contact has signed distance exactly 0.0 == margin; MuJoCo lists it and
    excludes it (``dist < includemargin`` is false) -> no force.
    'active' (the alternative of SURVEY A.3, simplified): a pyramidal friction cone whose normal row has a zero
    Jacobian for these three DOFs leaves opposing pairs of unilateral edges +-mu J_i per direction i (x, y, torsion).
    For a pair the regularised dual has the closed form f = -J_i^T d (J_i a_smooth + b J_i v) / A_ii with
    R = (1 - d) / d * mu^2 A_ii (mu cancels), i.e. every DOF loses the fraction d of (its smooth acceleration + b times
    its velocity): a = (1 - d) a_smooth - d b v.  (Decoupled A_ii; MuJoCo's diagApprox / impratio details cannot be
    checked here -- this is an order-of-magnitude model of that reading, not a restatement of the solver.)"""
    if CONTACT_MODEL == 'excluded':
        return np.zeros(3)
    A = mass_matrix(qpos[2]) + h * np.diag(DAMPING)
    a_smooth = np.linalg.solve(A, qfrc_smooth)
    return A @ (-CONTACT_IMPEDANCE * (a_smooth + CONTACT_B * qvel))


# ---- walls (ZoneEnvBase.py:55-62, `walled=True`): box geoms the robot's sphere can touch ------------------------
# PARITY UNPINNED, and SIMPLIFIED: MuJoCo's soft-contact model as recalled from its documentation ("Computation /
# Solver parameters"), default solref = (timeconst 0.02, dampratio 1) and solimp = (0.9, 0.95, 0.001, 0.5, 2), for the
# NORMAL row of each sphere-box contact only: no friction rows (pyramidal cone edges), no contacts of the `pointarrow`
# box geom, regulariser R_ii = (1 - d) / d x the exact A_ii instead of MuJoCo's diagApprox.  What is restated: contact
# detection (sphere centre to the nearest point of each box, dist < 0), impedance d(|dist|), reference acceleration
# a_ref = -b v - k dist with b = 2 / (dmax tc), k = d / (dmax^2 tc^2 dr^2), and the constrained minimisation
# f >= 0,  (A + R) f + J a_smooth - a_ref >= 0, complementary -- solved by projected Gauss-Seidel to convergence.
SOLREF_TC, SOLREF_DR = 0.02, 1.0
SOLIMP_D0, SOLIMP_DMAX, SOLIMP_WIDTH, SOLIMP_MID, SOLIMP_POWER = 0.9, 0.95, 0.001, 0.5, 2.0


def impedance(dist):
    """d(r) of solimp for penetration r = |dist| (MuJoCo 2.0's five-parameter sigmoid)."""
    x = min(abs(dist) / SOLIMP_WIDTH, 1.0)
    if x <= SOLIMP_MID:
        y = x ** SOLIMP_POWER / SOLIMP_MID ** (SOLIMP_POWER - 1.0)
    else:
        y = 1.0 - (1.0 - x) ** SOLIMP_POWER / (1.0 - SOLIMP_MID) ** (SOLIMP_POWER - 1.0)
    return SOLIMP_D0 + y * (SOLIMP_DMAX - SOLIMP_D0)


def sphere_box_contacts(centre_xy, boxes_xy, half):
    """Contacts of the robot's sphere (radius SPHERE_R, centre height == box centre height, so the geometry is planar)
    with axis-aligned boxes of half-size `half`: list of (normal (2,), dist < 0), normal from the box to the sphere."""
    out = []
    c = np.asarray(centre_xy, dtype=np.float64)
    near = np.abs(boxes_xy - c).max(axis=1) < half + SPHERE_R
    for b in boxes_xy[near]:
        nearest = np.minimum(np.maximum(c, b - half), b + half)
        d = c - nearest
        ln = math.sqrt(float(d @ d))
        if ln == 0.0:          # centre inside a box: never reached with these stiffnesses
            continue
        dist = ln - SPHERE_R
        if dist < 0.0:
            out.append((d / ln, dist))
    return out


def wall_force(qpos, qvel, qfrc_smooth, walls):
    """qfrc_constraint of the sphere-wall contacts.  `walls` = dict(p0 (2,), rot0, boxes (K, 2), half)."""
    c0, s0 = math.cos(walls['rot0']), math.sin(walls['rot0'])
    ax_x, ax_y = np.array([c0, s0]), np.array([-s0, c0])       # the slides' axes in the world
    centre = walls['p0'] + ax_x * qpos[0] + ax_y * qpos[1]
    contacts = sphere_box_contacts(centre, walls['boxes'], walls['half'])
    if not contacts:
        return np.zeros(3)
    J = np.array([[n @ ax_x, n @ ax_y, 0.0] for n, _ in contacts])    # the contact point is on the hinge axis' normal
    Minv = np.linalg.inv(mass_matrix(qpos[2]))
    A = J @ Minv @ J.T
    a0 = J @ (Minv @ qfrc_smooth)
    vel = J @ qvel
    d = np.array([impedance(dist) for _, dist in contacts])
    b = 2.0 / (SOLIMP_DMAX * SOLREF_TC)
    k = d / (SOLIMP_DMAX ** 2 * SOLREF_TC ** 2 * SOLREF_DR ** 2)
    aref = -b * vel - k * np.array([dist for _, dist in contacts])
    R = (1.0 - d) / d * np.diag(A)
    f = np.zeros(len(contacts))
    for _ in range(10000):                                      # projected Gauss-Seidel
        delta = 0.0
        for i in range(len(f)):
            resid = A[i] @ f + R[i] * f[i] + a0[i] - aref[i]
            fi = max(0.0, f[i] - resid / (A[i, i] + R[i]))
            delta = max(delta, abs(fi - f[i]))
            f[i] = fi
        if delta < 1e-15:
            break
    return J.T @ f


def substep(qpos, qvel, ctrl, h=TIMESTEP, walls=None):
    """One mj_step.  Returns new (qpos, qvel); inputs are not modified."""
    theta = qpos[2]
    M = mass_matrix(theta)
    passive = -DAMPING * qvel
    smooth = passive - bias_force(theta, qvel) + actuator_force(theta, qvel, ctrl)
    total = smooth + constraint_force(qpos, qvel, smooth)
    if walls is not None:
        total = total + wall_force(qpos, qvel, smooth, walls)
    qacc = np.linalg.solve(M + h * np.diag(DAMPING), total)
    qvel_new = qvel + h * qacc
    qpos_new = qpos + h * qvel_new
    return qpos_new, qvel_new


class _Model:
    """The slice of ``sim.model`` the reference touches (geom ids, rgba)."""

    def __init__(self, geom_names):
        self._geom_ids = {n: i for i, n in enumerate(geom_names)}
        self.geom_rgba = np.zeros((len(geom_names), 4))
        self.actuator_ctrlrange = np.array([[-1.0, 1.0], [-1.0, 1.0]])
        self.nu = 2

    def geom_name2id(self, name):
        return self._geom_ids[name]


class _Data:
    def __init__(self, sim):
        self._sim = sim
        self.qpos = np.zeros(3)
        self.qvel = np.zeros(3)
        self.ctrl = np.zeros(2)
        self.time = 0.0

    def get_body_xpos(self, name):
        return self._sim._xpos[name]

    def get_body_xquat(self, name):
        return self._sim._xquat[name]

    def get_body_xvelp(self, name):
        return self._sim._xvelp[name]

    def get_body_xvelr(self, name):
        return self._sim._xvelr[name]


class PointSim:
    """Stand-in for the ``MjSim`` World.build makes: Point robot + static bodies.

    ``static_bodies`` maps body name -> world position (3,), e.g. the zone
    cylinders added by ``ZoneEnvBase.build_world_config`` (ZoneEnvBase.py:124-140;
    contype = conaffinity = 0, so they never collide).
    """

    def __init__(self, robot_xy, robot_rot, static_bodies=None, geom_rgba=None, wall_boxes=None, wall_half=0.1):
        # wall_boxes: (K, 2) centres of the `walled=True` box geoms (ZoneEnvBase.py:55-62); None = no walls
        self.walls = None if wall_boxes is None or not len(wall_boxes) else {
            'p0': np.array([robot_xy[0], robot_xy[1]], dtype=np.float64), 'rot0': float(robot_rot),
            'boxes': np.asarray(wall_boxes, dtype=np.float64).reshape(-1, 2), 'half': float(wall_half)}
        self.robot_p0 = np.array([robot_xy[0], robot_xy[1], BODY_Z], dtype=np.float64)
        q0 = np.array([math.cos(robot_rot / 2.0), 0.0, 0.0, math.sin(robot_rot / 2.0)])
        self.robot_q0 = q0 / math.sqrt(float(q0 @ q0))
        self.static = {k: np.asarray(v, dtype=np.float64) for k, v in (static_bodies or {}).items()}
        names = ['floor', 'robot', 'pointarrow'] + list(self.static.keys())
        self.model = _Model(names)
        if geom_rgba:
            for k, v in geom_rgba.items():
                self.model.geom_rgba[self.model.geom_name2id(k)] = v
        self.data = _Data(self)
        self._xpos, self._xquat, self._xvelp, self._xvelr = {}, {}, {}, {}
        self.forward()

    def step(self):
        self.data.qpos, self.data.qvel = substep(self.data.qpos, self.data.qvel, self.data.ctrl, walls=self.walls)
        self.data.time += TIMESTEP

    def forward(self):
        """mj_kinematics + body velocities for the quantities the env reads."""
        q, v = self.data.qpos, self.data.qvel
        ax_x = quat_rotate(self.robot_q0, np.array([1.0, 0.0, 0.0]))
        ax_y = quat_rotate(self.robot_q0, np.array([0.0, 1.0, 0.0]))
        xpos = self.robot_p0 + ax_x * q[0] + ax_y * q[1]
        qloc = np.array([math.cos(q[2] / 2.0), 0.0, 0.0, math.sin(q[2] / 2.0)])
        xquat = quat_mul(self.robot_q0, qloc)
        xquat = xquat / math.sqrt(float(xquat @ xquat))
        self._xpos['robot'] = xpos
        self._xquat['robot'] = xquat
        # body origin lies on the hinge axis: no omega x r contribution
        self._xvelp['robot'] = ax_x * v[0] + ax_y * v[1]
        self._xvelr['robot'] = np.array([0.0, 0.0, v[2]])
        for k, p in self.static.items():
            self._xpos[k] = p
