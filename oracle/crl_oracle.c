/*
 * crl_oracle.c -- CPU oracle, C twin of oracle/{mj_point,sg_engine,zone_env}.py.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package links or loads this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs do.  PARITY UNPINNED for the physics and the layout sampler (they restate
 * un-vendored third-party code: MuJoCo 2.0 via mujoco-py==2.0.2.9 and safety-gym, see
 * oracle/mj_point.py); the task logic is pinned by tests/golden/*.npz, which were
 * recorded from the real reference task files (tests/golden/gen_golden.py), and this
 * file is checked against the same fixtures and against the Python oracle
 * (tests/test_c_oracle.py).
 *
 * Part 1  numpy legacy RandomState (MT19937 + the four distributions the path uses),
 *         so that reset(seed) reproduces the reference's draws: TTSP_env.py:19-21
 *         (beta), colour_match_env.py:57-68 (choice), Engine.sample_layout /
 *         random_rot / step's frameskip binomial [upstream] (uniform, binomial).
 * Part 2  fp64 Point-robot substep and the per-step order (Appendix B).
 * Part 3  a threaded random-action rollout used as the timed CPU stand-in for
 *         ParallelEnv over mujoco-py (penv.py:4-21).
 * Part 4  the DESIGN twin of the CUDA Philox reset: restates the product's own
 *         counter-based sampler (not the reference) so device resets can be checked.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ======================= Part 1: numpy legacy RandomState ====================== */
typedef struct {
  uint32_t key[624];
  int pos;
  int has_gauss;
  double gauss;
} MT;

static void mt_seed(MT* s, uint32_t seed) {
  /* RandomState(int) -> init_genrand */
  for (int i = 0; i < 624; ++i) {
    s->key[i] = seed;
    seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
  }
  s->pos = 624;
  s->has_gauss = 0;
  s->gauss = 0.0;
}

static void mt_gen(MT* s) {
  uint32_t* k = s->key;
  for (int i = 0; i < 624; ++i) {
    uint32_t y = (k[i] & 0x80000000u) | (k[(i + 1) % 624] & 0x7fffffffu);
    k[i] = k[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  s->pos = 0;
}

static uint32_t mt_u32(MT* s) {
  if (s->pos == 624) mt_gen(s);
  uint32_t y = s->key[s->pos++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

static double mt_double(MT* s) {
  const uint32_t a = mt_u32(s) >> 5, b = mt_u32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

static double mt_uniform(MT* s, double lo, double hi) { return lo + (hi - lo) * mt_double(s); }

static double mt_gauss(MT* s) {
  if (s->has_gauss) {
    s->has_gauss = 0;
    const double t = s->gauss;
    s->gauss = 0.0;
    return t;
  }
  double f, x1, x2, r2;
  do {
    x1 = 2.0 * mt_double(s) - 1.0;
    x2 = 2.0 * mt_double(s) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(-2.0 * log(r2) / r2);
  s->gauss = f * x1;
  s->has_gauss = 1;
  return f * x2;
}

static double mt_std_exponential(MT* s) { return -log(1.0 - mt_double(s)); }

static double mt_std_gamma(MT* s, double shape) {
  if (shape == 1.0) return mt_std_exponential(s);
  if (shape < 1.0) {
    for (;;) {
      const double U = mt_double(s), V = mt_std_exponential(s);
      if (U <= 1.0 - shape) {
        const double X = pow(U, 1.0 / shape);
        if (X <= V) return X;
      } else {
        const double Y = -log((1.0 - U) / shape), X = pow(1.0 - shape + shape * Y, 1.0 / shape);
        if (X <= V + Y) return X;
      }
    }
  }
  const double b = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * b);
  for (;;) {
    double X, V;
    do {
      X = mt_gauss(s);
      V = 1.0 + c * X;
    } while (V <= 0.0);
    V = V * V * V;
    const double U = mt_double(s);
    if (U < 1.0 - 0.0331 * (X * X) * (X * X)) return b * V;
    if (log(U) < 0.5 * X * X + b * (1.0 - V + log(V))) return b * V;
  }
}

static double mt_beta(MT* s, double a, double b) {
  if (a <= 1.0 && b <= 1.0) {
    for (;;) { /* Johnk */
      const double U = mt_double(s), V = mt_double(s);
      const double X = pow(U, 1.0 / a), Y = pow(V, 1.0 / b), XpY = X + Y;
      if (XpY <= 1.0 && XpY > 0.0) return X / XpY;
    }
  }
  const double Ga = mt_std_gamma(s, a), Gb = mt_std_gamma(s, b);
  return Ga / (Ga + Gb);
}

/* randint(0, n) for small n: masked rejection on 32-bit words */
static uint32_t mt_randint(MT* s, uint32_t n) {
  uint32_t rng = n - 1u, mask = rng, v;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  if (rng == 0u) return 0u;
  while ((v = mt_u32(s) & mask) > rng) {}
  return v;
}

/* binomial(n, 1.0): the inversion branch on q = 0 draws one double and returns n */
static int mt_binomial_p1(MT* s, int n) { (void)mt_double(s); return n; }

/* ============================ Part 2: the env ================================= */
#define MAXZ 16
enum { T_TSP = 0, T_TTSP = 1, T_CM = 2 };

static const double H = 0.002, PI = 3.14159265358979323846;
#define M_SPHERE (4.0 / 3.0 * PI * 0.1 * 0.1 * 0.1)
#define M_BOX (8.0 * 0.05 * 0.05 * 0.05)
#define MASS (M_SPHERE + M_BOX)
#define COMX (M_BOX * 0.1 / MASS)
#define IHINGE (0.4 * M_SPHERE * 0.01 + M_BOX / 3.0 * (0.0025 + 0.0025) + M_BOX * 0.01)

typedef struct {
  int task, N, Z, num_steps;
  int64_t seed;
  int has_rs;
  MT rs;
  double xy0[2], rot0, ax_x[3], ax_y[3], q0[4];
  double zone_xy[MAXZ][2];
  int visited[MAXZ], colours[MAXZ], cooldown[MAXZ];
  int64_t zone_max_steps[MAXZ];
  int goal_dist, steps, done, event;
  double qpos[3], qvel[3];
  /* kinematics of the last forward() */
  double xpos[2], xquat[4], xvelp[2], xvelr;
} OEnv;

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void quat_rotate(const double* q, const double* v, double* out) {
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double t0 = -x * v[0] - y * v[1] - z * v[2];
  const double t1 = w * v[0] + y * v[2] - z * v[1];
  const double t2 = w * v[1] + z * v[0] - x * v[2];
  const double t3 = w * v[2] + x * v[1] - y * v[0];
  out[0] = -t0 * x + t1 * w - t2 * z + t3 * y;
  out[1] = -t0 * y + t2 * w - t3 * x + t1 * z;
  out[2] = -t0 * z + t3 * w - t1 * y + t2 * x;
}

static void forward(OEnv* e) {
  e->xpos[0] = e->xy0[0] + e->ax_x[0] * e->qpos[0] + e->ax_y[0] * e->qpos[1];
  e->xpos[1] = e->xy0[1] + e->ax_x[1] * e->qpos[0] + e->ax_y[1] * e->qpos[1];
  const double cw = cos(e->qpos[2] / 2.0), sw = sin(e->qpos[2] / 2.0);
  double w = e->q0[0] * cw - e->q0[3] * sw, z = e->q0[0] * sw + e->q0[3] * cw;
  const double n = sqrt(w * w + z * z);
  e->xquat[0] = w / n; e->xquat[1] = 0.0; e->xquat[2] = 0.0; e->xquat[3] = z / n;
  e->xvelp[0] = e->ax_x[0] * e->qvel[0] + e->ax_y[0] * e->qvel[1];
  e->xvelp[1] = e->ax_x[1] * e->qvel[0] + e->ax_y[1] * e->qvel[1];
  e->xvelr = e->qvel[2];
}

/* 3x3 solve, Gaussian elimination with partial pivoting */
static void solve3(double A[3][3], double b[3], double x[3]) {
  for (int c = 0; c < 3; ++c) {
    int p = c;
    for (int r = c + 1; r < 3; ++r) if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
    if (p != c) {
      for (int k = 0; k < 3; ++k) { double t = A[c][k]; A[c][k] = A[p][k]; A[p][k] = t; }
      double t = b[c]; b[c] = b[p]; b[p] = t;
    }
    for (int r = c + 1; r < 3; ++r) {
      const double f = A[r][c] / A[c][c];
      for (int k = c; k < 3; ++k) A[r][k] -= f * A[c][k];
      b[r] -= f * b[c];
    }
  }
  for (int r = 2; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < 3; ++k) s -= A[r][k] * x[k];
    x[r] = s / A[r][r];
  }
}

/* one mj_step: forward dynamics on the current state, implicit-damping Euler */
void oe_substep(double* qpos, double* qvel, const double* ctrl) {
  const double th = qpos[2], s = sin(th), c = cos(th), mc = MASS * COMX;
  const double damp[3] = {0.01, 0.01, 0.005};
  double A[3][3] = {{MASS, 0.0, -mc * s}, {0.0, MASS, mc * c}, {-mc * s, mc * c, IHINGE}};
  const double w2 = qvel[2] * qvel[2];
  const double bias[3] = {-mc * w2 * c, -mc * w2 * s, 0.0};
  const double u0 = clampd(ctrl[0], -1.0, 1.0), u1 = clampd(ctrl[1], -1.0, 1.0);
  const double fm = clampd(u0, -0.05, 0.05);
  const double fs = clampd(1.0 * u1 - 1.0 * (0.3 * qvel[2]), -0.05, 0.05);
  const double act[3] = {0.3 * fm * c, 0.3 * fm * s, 0.3 * fs};
  double rhs[3], acc[3];
  for (int i = 0; i < 3; ++i) {
    rhs[i] = -damp[i] * qvel[i] - bias[i] + act[i]; /* + constraint force: none */
    A[i][i] += H * damp[i];
  }
  solve3(A, rhs, acc);
  for (int i = 0; i < 3; ++i) qvel[i] += H * acc[i];
  for (int i = 0; i < 3; ++i) qpos[i] += H * qvel[i];
}

static int hamming_to_goal(const int* col, int n) {
  int nb = 0, ng = 0, nr = 0;
  for (int i = 0; i < n; ++i) { nb += col[i] == 0; ng += col[i] == 1; nr += col[i] == 2; }
  const int b = 2 * ng + nr, g = 2 * nr + nb, r = 2 * nb + ng;
  int m = b < g ? b : g;
  return m < r ? m : r;
}

void* oe_new(int task, int n_zones, int num_steps) {
  OEnv* e = (OEnv*)calloc(1, sizeof(OEnv));
  e->task = task; e->N = n_zones; e->Z = task == T_TSP ? 6 : 7; e->num_steps = num_steps;
  e->done = 1;
  return e;
}
void oe_free(void* p) { free(p); }
void oe_seed(void* p, int64_t seed) { ((OEnv*)p)->seed = seed; }
int64_t oe_get_seed(void* p) { return ((OEnv*)p)->seed; }

static void install_layout(OEnv* e, const double* xy0, double rot0, const double* zone_xy) {
  e->xy0[0] = xy0[0]; e->xy0[1] = xy0[1]; e->rot0 = rot0;
  double q[4] = {cos(rot0 / 2.0), 0.0, 0.0, sin(rot0 / 2.0)};
  const double n = sqrt(q[0] * q[0] + q[3] * q[3]);
  q[0] /= n; q[3] /= n;
  memcpy(e->q0, q, sizeof(q));
  const double ex[3] = {1.0, 0.0, 0.0}, ey[3] = {0.0, 1.0, 0.0};
  quat_rotate(q, ex, e->ax_x);
  quat_rotate(q, ey, e->ax_y);
  for (int i = 0; i < e->N; ++i) {
    e->zone_xy[i][0] = zone_xy[2 * i]; e->zone_xy[i][1] = zone_xy[2 * i + 1];
    e->visited[i] = 0; e->cooldown[i] = 0;
  }
  if (e->task == T_CM) e->goal_dist = hamming_to_goal(e->colours, e->N);
  memset(e->qpos, 0, sizeof(e->qpos));
  memset(e->qvel, 0, sizeof(e->qvel));
  e->steps = 0; e->done = 0; e->event = 0;
  forward(e);
}

/* Engine.sample_layout with the legacy stream: robot (keepout .4) then N zones (.55) */
static void sample_layout_mt(MT* rs, int N, double* xy0, double* rot0, double* zone_xy) {
  double px[MAXZ + 1], py[MAXZ + 1], pk[MAXZ + 1];
  for (int attempt = 0; attempt < 10000; ++attempt) {
    int ok_layout = 1;
    for (int k = 0; k <= N && ok_layout; ++k) {
      const double keep = k == 0 ? 0.4 : 0.55, lo = -3.0 + keep, hi = 3.0 - keep;
      int found = 0;
      for (int t = 0; t < 100 && !found; ++t) {
        const double x = mt_uniform(rs, lo, hi), y = mt_uniform(rs, lo, hi);
        int valid = 1;
        for (int q = 0; q < k; ++q) {
          const double dx = x - px[q], dy = y - py[q];
          if (sqrt(dx * dx + dy * dy) < pk[q] + 0.0 + keep) { valid = 0; break; }
        }
        if (valid) { px[k] = x; py[k] = y; pk[k] = keep; found = 1; }
      }
      ok_layout = found;
    }
    if (ok_layout) break;
  }
  xy0[0] = px[0]; xy0[1] = py[0];
  *rot0 = mt_uniform(rs, 0.0, 2.0 * PI);
  for (int i = 0; i < N; ++i) {
    zone_xy[2 * i] = px[i + 1]; zone_xy[2 * i + 1] = py[i + 1];
    (void)mt_uniform(rs, 0.0, 2.0 * PI); /* ZoneEnvBase.py:132: one cosmetic rot per zone */
  }
}

/* TimedTSPEnv.reset / ColourMatchEnv.reset -> TSPEnv.reset -> Engine.reset */
void oe_reset(void* p) {
  OEnv* e = (OEnv*)p;
  if (e->task == T_TTSP) {
    MT t; mt_seed(&t, (uint32_t)e->seed);
    for (int i = 0; i < e->N; ++i) e->zone_max_steps[i] = (int64_t)(mt_beta(&t, 3.0, 1.5) * e->num_steps);
  } else if (e->task == T_CM) {
    MT t; mt_seed(&t, (uint32_t)e->seed);
    for (int i = 0; i < e->N; ++i) e->colours[i] = (int)mt_randint(&t, 3u);
  }
  e->seed += 1;
  mt_seed(&e->rs, (uint32_t)e->seed);
  e->has_rs = 1;
  double xy0[2], rot0, zxy[2 * MAXZ];
  sample_layout_mt(&e->rs, e->N, xy0, &rot0, zxy);
  install_layout(e, xy0, rot0, zxy);
}

void oe_reset_layout(void* p, const double* xy0, double rot0, const double* zone_xy,
                     const int64_t* zone_max_steps, const int64_t* colours) {
  OEnv* e = (OEnv*)p;
  for (int i = 0; i < e->N; ++i) {
    if (e->task == T_TTSP && zone_max_steps) e->zone_max_steps[i] = zone_max_steps[i];
    if (e->task == T_CM && colours) e->colours[i] = (int)colours[i];
  }
  e->has_rs = 0;
  install_layout(e, xy0, rot0, zone_xy);
}

void oe_set_state(void* p, const double* qpos, const double* qvel) {
  OEnv* e = (OEnv*)p;
  memcpy(e->qpos, qpos, 3 * sizeof(double));
  memcpy(e->qvel, qvel, 3 * sizeof(double));
  forward(e);
}
void oe_get_state(void* p, double* qpos, double* qvel) {
  OEnv* e = (OEnv*)p;
  memcpy(qpos, e->qpos, 3 * sizeof(double));
  memcpy(qvel, e->qvel, 3 * sizeof(double));
}
void oe_get_layout(void* p, double* xy0, double* rot0, double* zone_xy, int64_t* zone_max_steps, int64_t* colours) {
  OEnv* e = (OEnv*)p;
  xy0[0] = e->xy0[0]; xy0[1] = e->xy0[1]; *rot0 = e->rot0;
  for (int i = 0; i < e->N; ++i) {
    zone_xy[2 * i] = e->zone_xy[i][0]; zone_xy[2 * i + 1] = e->zone_xy[i][1];
    zone_max_steps[i] = e->zone_max_steps[i]; colours[i] = e->colours[i];
  }
}
/* task state: steps, done, event, goal_dist, then visited[N], colours[N], cooldown[N] */
void oe_get_task_state(void* p, int64_t* out) {
  OEnv* e = (OEnv*)p;
  out[0] = e->steps; out[1] = e->done; out[2] = e->event; out[3] = e->goal_dist;
  for (int i = 0; i < e->N; ++i) {
    out[4 + i] = e->visited[i]; out[4 + e->N + i] = e->colours[i]; out[4 + 2 * e->N + i] = e->cooldown[i];
  }
}

/* one env.step(); returns done.  Order: SURVEY.md Appendix B. */
int oe_step(void* p, const double* action, double* reward_out, int* goal_met_out) {
  OEnv* e = (OEnv*)p;
  const int N = e->N;
  if (e->task == T_CM)
    for (int i = 0; i < N; ++i) if (e->cooldown[i] > 0) e->cooldown[i] -= 1;
  const double ctrl[2] = {clampd(action[0], -1.0, 1.0), clampd(action[1], -1.0, 1.0)};
  if (e->has_rs) (void)mt_binomial_p1(&e->rs, 10);
  int fired = -1;
  for (int i = 0; i < N && fired < 0; ++i) {
    const int eligible = e->task == T_CM ? (e->cooldown[i] == 0) : !e->visited[i];
    if (!eligible) continue;
    const double dx = e->zone_xy[i][0] - e->xpos[0], dy = e->zone_xy[i][1] - e->xpos[1];
    volatile double xx = dx * dx, yy = dy * dy;
    if (sqrt(xx + yy) <= 0.2) fired = i;
  }
  if (fired >= 0) {
    if (e->task == T_CM) { e->colours[fired] = (e->colours[fired] + 1) % 3; e->cooldown[fired] = 150; }
    else e->visited[fired] = 1;
  }
  for (int k = 0; k < 10; ++k) oe_substep(e->qpos, e->qvel, ctrl);
  forward(e);
  int ev = 0, goal;
  if (e->task == T_CM) {
    if (fired >= 0) { const int nd = hamming_to_goal(e->colours, N); ev = e->goal_dist - nd; e->goal_dist = nd; }
    goal = e->goal_dist == 0;
  } else {
    ev = fired >= 0 ? 1 : 0;
    goal = 1;
    for (int i = 0; i < N; ++i) goal = goal && e->visited[i];
  }
  e->event = ev;
  double reward = (double)ev;
  if (goal) { reward += (e->num_steps - e->steps) * 0.01; e->done = 1; }
  e->steps += 1;
  if (e->steps >= e->num_steps) e->done = 1;
  if (e->task == T_TTSP && !e->done)
    for (int i = 0; i < N; ++i) {
      const double t = e->visited[i] ? 1.0 : (double)(e->zone_max_steps[i] - e->steps) / (double)e->num_steps;
      if (t <= 0.0) e->done = 1;
    }
  *reward_out = reward;
  *goal_met_out = goal;
  return e->done;
}

int oe_event(void* p) { return ((OEnv*)p)->event; }

void oe_obs(void* p, double* obs, double* zone_obs) {
  OEnv* e = (OEnv*)p;
  forward(e);
  obs[0] = 1.0 - (double)e->steps / (double)e->num_steps;
  obs[1] = e->xpos[0] / 3.0; obs[2] = e->xpos[1] / 3.0;
  const float q0 = (float)e->xquat[0], q3 = (float)e->xquat[3];
  const float a = q0 * q0, b = q3 * q3;
  obs[3] = (double)(float)(a - b);
  const float two_q0 = 2.0f * q0;
  obs[4] = (double)(float)(two_q0 * q3);
  obs[5] = e->xvelp[0] / 1.5; obs[6] = e->xvelp[1] / 1.5; obs[7] = e->xvelr / 3.0;
  for (int i = 0; i < e->N; ++i) {
    double* z = zone_obs + i * e->Z;
    z[0] = e->zone_xy[i][0] / 3.0; z[1] = e->zone_xy[i][1] / 3.0;
    if (e->task == T_CM) {
      z[2] = e->colours[i] == 2; z[3] = e->colours[i] == 1; z[4] = e->colours[i] == 0;
      z[6] = (double)((float)e->cooldown[i] / 150.0f);
    } else {
      z[2] = e->visited[i] ? 1.0 : 0.0; z[3] = 1.0; z[4] = e->visited[i] ? 0.0 : 1.0;
      if (e->task == T_TTSP)
        z[6] = e->visited[i] ? 1.0 : (double)(e->zone_max_steps[i] - e->steps) / (double)e->num_steps;
    }
    z[5] = 0.25;
  }
}

/* ===================== Part 3: timed threaded rollout ========================== */
typedef struct {
  int task, N, num_steps;
  int64_t first_seed;
  double seconds;
  int64_t steps_done;
  uint64_t rng;
} Work;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static double xs_uniform(uint64_t* s) {
  uint64_t x = *s;
  x ^= x << 13; x ^= x >> 7; x ^= x << 17;
  *s = x;
  return (double)(x >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}

static void* rollout_thread(void* arg) {
  Work* w = (Work*)arg;
  OEnv* e = (OEnv*)oe_new(w->task, w->N, w->num_steps);
  /* one env per thread, as penv.py runs one env per process; per-episode seeds cycle
     over 100 maps like make_train_env's FixedSeedsWrapper(min_seed=1, max_seed=100) */
  int64_t ep = 0;
  double obs[8], zobs[MAXZ * 7];
  oe_seed(e, w->first_seed + (ep++ % 100));
  oe_reset(e);
  oe_obs(e, obs, zobs);
  const double t_end = now_s() + w->seconds;
  int64_t n = 0;
  for (;;) {
    for (int k = 0; k < 256; ++k) {
      const double a[2] = {xs_uniform(&w->rng), xs_uniform(&w->rng)};
      double r; int g;
      const int d = oe_step(e, a, &r, &g);
      if (d) { oe_seed(e, w->first_seed + (ep++ % 100)); oe_reset(e); }
      oe_obs(e, obs, zobs);   /* the reference builds the full observation every step */
      ++n;
    }
    if (now_s() >= t_end) break;
  }
  w->steps_done = n;
  oe_free(e);
  return NULL;
}

/* returns env-steps/s over all threads; *total = steps executed */
double oe_timed_rollout(int task, int n_zones, int num_steps, int threads, double seconds, int64_t* total) {
  if (threads < 1) threads = 1;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  Work* w = (Work*)calloc(threads, sizeof(Work));
  const double t0 = now_s();
  for (int i = 0; i < threads; ++i) {
    w[i].task = task; w[i].N = n_zones; w[i].num_steps = num_steps; w[i].first_seed = 1 + 1000 * i;
    w[i].seconds = seconds; w[i].rng = 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1);
    pthread_create(&th[i], NULL, rollout_thread, &w[i]);
  }
  int64_t sum = 0;
  for (int i = 0; i < threads; ++i) { pthread_join(th[i], NULL); sum += w[i].steps_done; }
  const double dt = now_s() - t0;
  *total = sum;
  free(th); free(w);
  return (double)sum / dt;
}

/* ============ Part 4: design twin of the CUDA Philox reset (not reference) ====== */
static void philox4x32(const uint32_t ctr_in[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}
void ph_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32(ctr, key[0], key[1], out); }

enum { TAG_LAYOUT = 1, TAG_ROT = 2, TAG_TASK = 3, TAG_SEED = 4, TAG_ACTION = 5 };

static void ph_draw(int64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t tag, uint32_t out[4]) {
  const uint32_t ctr[4] = {c0, c1, c2, tag};
  philox4x32(ctr, (uint32_t)(uint64_t)seed, (uint32_t)((uint64_t)seed >> 32), out);
}
static float ph_u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
static double ph_u01d(uint32_t hi, uint32_t lo) {
  const uint64_t v = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
  return ((double)v + 1.0) * (1.0 / 9007199254740992.0);
}
static double ph_gamma(int64_t seed, double a, uint32_t zone, uint32_t which) {
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (uint32_t it = 0;; ++it) {
    uint32_t r0[4], r1[4];
    ph_draw(seed, it, zone, which, TAG_TASK, r0);
    ph_draw(seed, it, zone, which + 2u, TAG_TASK, r1);
    const double u1 = ph_u01d(r0[0], r0[1]), u2 = ph_u01d(r0[2], r0[3]), u3 = ph_u01d(r1[0], r1[1]);
    const double x = sqrt(-2.0 * log(u1)) * cos(2.0 * PI * u2);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    if (log(u3) < 0.5 * x * x + d - d * v + d * log(v) || it > 64u) return d * v;
  }
}

/* gamma(k/2), k a small integer: -log of the product of floor(k/2) uniforms, plus for odd
 * k the Box-Muller half-square -log(U) cos^2(2 pi V).  Uniform j = half j&1 of block j>>1. */
static double ph_gamma_half_integer(int64_t seed, int k, uint32_t zone, uint32_t which) {
  const int m = k >> 1, n_u = m + ((k & 1) ? 2 : 0);
  double prod = 1.0, U = 1.0, V = 0.0;
  for (int j = 0; j < n_u; j += 2) {
    uint32_t r[4];
    ph_draw(seed, (uint32_t)(j >> 1), zone, which, TAG_TASK, r);
    const double ua = ph_u01d(r[0], r[1]), ub = ph_u01d(r[2], r[3]);
    if (j < m) prod *= ua; else if (j == m) U = ua; else V = ua;
    if (j + 1 < m) prod *= ub; else if (j + 1 == m) U = ub; else if (j + 1 < n_u) V = ub;
  }
  double g = -log(prod);
  if (k & 1) {
    const double c = cos(2.0 * PI * V);
    g += -log(U) * (c * c);
  }
  return g;
}
static double ph_gamma_draw(int64_t seed, double a, uint32_t zone, uint32_t which) {
  const double t = 2.0 * a;
  const int k = (int)t;
  if ((double)k == t && k >= 1 && k <= 16) return ph_gamma_half_integer(seed, k, zone, which);
  return ph_gamma(seed, a, zone, which);
}
/* one gamma draw of the device stream, for distribution tests */
double ph_gamma_sample(int64_t seed, double a, uint32_t zone, uint32_t which) { return ph_gamma_draw(seed, a, zone, which); }

/* The product's reset for one env, sequentially.  seed_in = CrlState.seed[e] before the
 * reset (seed_mode 0) ; for seed_mode 1 the seed is first re-drawn in [min_seed, max_seed]
 * from Philox(key = global env index, counter = episode). */
void ph_reset(int task, int N, int num_steps, int seed_mode, int64_t min_seed, int64_t max_seed,
              int64_t global_env, uint32_t episode, int64_t seed_in, double beta_a, double beta_b,
              float robot_keepout, float zone_keepout, float extent,
              const float* fixed /* CrlState.fixed_layout: float4[1 + N] or NULL */,
              float* xy0, float* rot0, float* zone_xy, int32_t* tmax, int32_t* colours, int64_t* seed_after) {
  int64_t seed = seed_in;
  if (seed_mode == 1) {
    const uint64_t span = (uint64_t)(max_seed - min_seed) + 1ull;
    uint32_t r[4];
    ph_draw(global_env, episode, 0u, 0u, TAG_SEED, r);
    const uint64_t v = ((uint64_t)r[0] << 32) | r[1];
    seed = min_seed + (int64_t)(span ? (v % span) : v);
  }
  for (int i = 0; i < N; ++i) {
    tmax[i] = 0; colours[i] = 0;
    if (task == T_TTSP) {
      const double ga = ph_gamma_draw(seed, beta_a, (uint32_t)i, 0u), gb = ph_gamma_draw(seed, beta_b, (uint32_t)i, 1u);
      int t = (int)((ga / (ga + gb)) * (double)num_steps);
      tmax[i] = t < 0 ? 0 : (t > 65535 ? 65535 : t);
    }
    if (task == T_CM) {
      uint32_t c = 3u;
      for (uint32_t it = 0; c == 3u && it < 64u; ++it) {
        uint32_t r[4];
        ph_draw(seed, it, (uint32_t)i, 0u, TAG_TASK, r);
        uint32_t bits = r[0];
        for (int q = 0; q < 16 && c == 3u; ++q, bits >>= 2) c = bits & 3u;
      }
      colours[i] = c == 3u ? 0 : (int32_t)c;
    }
  }
  seed += 1;
  float px[MAXZ + 1], py[MAXZ + 1];
  for (uint32_t attempt = 0; attempt < 10000u; ++attempt) {
    int ok_layout = 1;
    for (int k = 0; k <= N && ok_layout; ++k) {
      const float keep = k == 0 ? robot_keepout : zone_keepout;
      const float lo = -extent + keep, span = (extent - keep) - lo;
      int found = 0;
      const int pinned = fixed && (fixed[4 * k + 3] == 1.f || fixed[4 * k + 3] == 3.f);
      for (int j = 0; j < (pinned ? 1 : 100) && !found; ++j) {
        uint32_t r[4];
        /* try j = half (j & 1) of Philox block (j >> 1, k, attempt) */
        ph_draw(seed, (uint32_t)(j >> 1), (uint32_t)k, attempt, TAG_LAYOUT, r);
        volatile float mx = span * ph_u01(r[(j & 1) ? 2 : 0]), my = span * ph_u01(r[(j & 1) ? 3 : 1]);
        const float x = pinned ? fixed[4 * k] : lo + mx, y = pinned ? fixed[4 * k + 1] : lo + my;
        int valid = 1;
        for (int q = 0; q < k; ++q) {
          const float need = (q == 0 ? robot_keepout : zone_keepout) + keep;
          const float dx = x - px[q], dy = y - py[q];
          /* as the device forms it: d2 = fma(dy, dy, fl(dx * dx)), one rounding for the second product and the sum
             (packed FFMA2 in the background sampler, __fmaf_rn in the warp-cooperative one) */
          volatile float xx = dx * dx, nn = need * need;
          const float d2 = fmaf(dy, dy, xx);
          valid = valid && (d2 >= nn);
        }
        if (valid) { px[k] = x; py[k] = y; found = 1; }
      }
      ok_layout = found;
    }
    if (ok_layout) break;
  }
  uint32_t rr[4];
  ph_draw(seed, 0u, 0u, 0u, TAG_ROT, rr);
  xy0[0] = px[0]; xy0[1] = py[0];
  *rot0 = 6.2831855f * ph_u01(rr[0]);
  if (fixed && fixed[3] >= 2.f) *rot0 = fixed[2];
  for (int i = 0; i < N; ++i) { zone_xy[2 * i] = px[i + 1]; zone_xy[2 * i + 1] = py[i + 1]; }
  *seed_after = seed;
}

/* U(-1,1)^2 action of global env `ge` at step `step_index` (crl_step with actions == NULL) */
void ph_action(uint64_t action_seed, uint32_t ge, uint64_t step_index, float* a) {
  const uint32_t ctr[4] = {ge, (uint32_t)step_index, (uint32_t)(step_index >> 32), TAG_ACTION};
  uint32_t r[4];
  philox4x32(ctr, (uint32_t)action_seed, (uint32_t)(action_seed >> 32), r);
  a[0] = 2.f * ph_u01(r[0]) - 1.f;
  a[1] = 2.f * ph_u01(r[1]) - 1.f;
}
