// crl_core.cuh -- per-env arithmetic of the fused step, shared by the sm_100a
// kernels (crl_kernels.cu) and by a g++ build that tests/ uses to check the logic
// without a GPU (tests/hostcheck).  Nothing here touches memory layout.
//
// What it computes and the reference lines it stands in for:
//   substeps()      Engine.step's frameskip loop of sim.step() on xmls/point.xml
//                   (MuJoCo 2.0 mj_step, implicit-damping Euler)
//   inside_zone()   dist_xy(zone) <= zones_size, main/envs/TSP_env.py:59-60,
//                   colour_match_env.py:111-112, evaluated exactly as numpy does
//   hamming()       colour_match_env.py:38-55
//   philox4x32()    counter-based RNG for layouts, timeouts, colours and actions
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CRL_HD __host__ __device__ __forceinline__
#define CRL_HD_COLD inline __host__ __device__ __noinline__        /* rare paths: keep them out of the callers' register allocation */
#else
#define CRL_HD inline
#define CRL_HD_COLD inline
#endif

namespace crl {

// ---- point.xml, density 1 ---------------------------------
constexpr double kPi = 3.14159265358979323846;
constexpr double kH = 0.002;
constexpr double kMSphere = 4.0 / 3.0 * kPi * 0.1 * 0.1 * 0.1;
constexpr double kMBox = 8.0 * 0.05 * 0.05 * 0.05;
constexpr double kMass = kMSphere + kMBox;
constexpr double kCom = kMBox * 0.1 / kMass;
constexpr double kIHinge = 0.4 * kMSphere * 0.01 + kMBox / 3.0 * (0.0025 + 0.0025) + kMBox * 0.01;
constexpr double kDampLin = 0.01;
constexpr double kDampYaw = 0.005;
constexpr double kGear = 0.3;
constexpr double kForceLimit = 0.05;
// closed form of (M + h B) a = tau for M = [[m,0,-ks],[0,m,kc],[-ks,kc,I]]:
//   a_th = (tau_th - (k b / m')(s vx - c vy)) / D,  D = I' - k^2 / m'
//   a_x  = (tau_x + k s a_th) / m',  a_y = (tau_y - k c a_th) / m'
constexpr double kK = kMass * kCom;
constexpr double kMp = kMass + kH * kDampLin;
constexpr double kIp = kIHinge + kH * kDampYaw;
constexpr double kD = kIp - kK * kK / kMp;

struct Body {
  float X, Y, phi;   // world position and heading
  float vx, vy, w;   // world velocity and yaw rate
};

// ---- the contact term, isolated (twin of oracle/mj_point.py::constraint_force) ----------------
// Model 0 (canonical): the sphere/floor contact sits at signed distance exactly 0.0 == margin: MuJoCo lists it
// and excludes it, no constraint force.  Model 1: the alternative reading (contact active, pyramidal cone whose
// normal row vanishes for these DOFs): every DOF loses the fraction d of (its smooth acceleration + b x its
// velocity), d = 0.9 (solimp at zero penetration), b = 2 / (0.95 x 0.02) (solref).  A compile-time switch
// (-DCRL_CONTACT_MODEL=1), so that a correction -- should a MuJoCo 2.0 trace ever pin the question -- is this one
// function; tests/test_contact_hypothesis.py measures what the switch changes (the robot's terminal speed falls
// from 1.5 m/s, the reference's own velocity normaliser, ZoneEnvBase.py:223, to 3 mm/s).
#ifndef CRL_CONTACT_MODEL
#define CRL_CONTACT_MODEL 0
#endif
constexpr double kContactImpedance = 0.9;
constexpr double kContactB = 2.0 / (0.95 * 0.02);

template <int MODEL>
CRL_HD void constraint_acc(float& ax, float& ay, float& ath, float vx, float vy, float w) {
  if (MODEL == 1) {
    const float d = (float)kContactImpedance, b = (float)kContactB;
    ax = (1.f - d) * ax - d * b * vx;
    ay = (1.f - d) * ay - d * b * vy;
    ath = (1.f - d) * ath - d * b * w;
  }
}

CRL_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// ---- walls (`walled=True`, ZoneEnvBase.py:55-62): twin of oracle/mj_point.py::wall_force -------------------------
// 244 box geoms of half-size 0.1 centred on the square of half-width `extent` = 3 at every multiple of 0.1 (both rows and
// both columns list their corner box: the corners hold two boxes, as in the reference's list).  The robot's sphere
// (radius 0.1, centre at the boxes' centre height, so the geometry is planar) touches a box when the distance from its
// centre to the nearest point of the box is below the radius.  MuJoCo's soft contact as recalled (default solref 0.02 / 1,
// solimp 0.9 / 0.95 / 0.001 / 0.5 / 2), NORMAL rows only -- frictionless, the pointarrow box ignored, R_ii = (1 - d) / d
// A_ii: PARITY UNPINNED and simplified, see the oracle's header.  Contacts with the same normal and depth (the face
// contacts of one wall: two or three overlapping boxes) are one row with a multiplicity.  Returns the generalised
// constraint force (Fx, Fy) in the world frame; it has no yaw component (the contact normal passes through the hinge).
constexpr double kWallHalf = 0.1, kSphereR = 0.1, kWallPitch = 0.1;
constexpr double kSolTc = 0.02, kSolD0 = 0.9, kSolDmax = 0.95, kSolWidth = 0.001;
constexpr double kD0 = kIHinge - kK * kK / kMass;             // I - k^2 / m of the bare mass matrix
constexpr int kMaxWallRows = 12;

CRL_HD float wall_impedance(float dist) {
  const float x = fminf(fabsf(dist) * (float)(1.0 / kSolWidth), 1.f);
  const float y = x <= 0.5f ? 2.f * x * x : 1.f - 2.f * (1.f - x) * (1.f - x);     // midpoint 0.5, power 2
  return (float)kSolD0 + y * (float)(kSolDmax - kSolD0);
}

struct WallForce { float fx, fy; };

// (X, Y): sphere centre; (c, s) = (cos, sin) of the heading; (vx, vy): velocity; (tx, ty, tth): the smooth generalised
// force (actuators + centrifugal bias + damping) of this substep.  By value on purpose (see profiles/r02_notes.md on
// what a by-reference Env does to the step kernel's stack frame).
CRL_HD_COLD WallForce wall_force(float X, float Y, float c, float s, float vx, float vy, float tx, float ty, float tth,
                            float extent) {
  float nx[kMaxWallRows], ny[kMaxWallRows], dist[kMaxWallRows], mult[kMaxWallRows];
  int rows = 0;
  const float half = (float)kWallHalf, R = (float)kSphereR, pitch = (float)kWallPitch;
  const int jmax = (int)(extent / pitch + 0.5f);
#pragma unroll 1
  for (int wall = 0; wall < 4; ++wall) {
    // wall 0 / 1: boxes at (+-extent, j pitch); wall 2 / 3: boxes at (j pitch, +-extent)
    const float sign = (wall & 1) ? -1.f : 1.f;
    const float across = wall < 2 ? X : Y, along = wall < 2 ? Y : X;
    if (sign * across <= extent - half - R - 1e-3f) continue;              // this wall's inner face is out of reach
    const int jc = (int)rintf(along / pitch);
#pragma unroll 1
    for (int dj = -2; dj <= 2; ++dj) {
      const int j = jc + dj;
      if (j < -jmax || j > jmax) continue;
      const float bc_across = sign * extent, bc_along = (float)j * pitch;
      const float da = across - clampf(across, bc_across - half, bc_across + half);
      const float dl = along - clampf(along, bc_along - half, bc_along + half);
      const float len = sqrtf(da * da + dl * dl);
      if (len <= 0.f || len >= R) continue;
      const float n_across = da / len, n_along = dl / len;
      const float ex = wall < 2 ? n_across : n_along, ey = wall < 2 ? n_along : n_across;
      const float d = len - R;
      int hit = -1;
      for (int r = 0; r < rows; ++r)
        if (nx[r] == ex && ny[r] == ey && dist[r] == d) hit = r;
      if (hit >= 0) { mult[hit] += 1.f; continue; }
      if (rows < kMaxWallRows) { nx[rows] = ex; ny[rows] = ey; dist[rows] = d; mult[rows] = 1.f; ++rows; }
    }
  }
  WallForce out{0.f, 0.f};
  if (rows == 0) return out;
  // u_r = M^-1 (nx, ny, 0) with the bare mass matrix (closed form); A_rq = n_r . u_q; a0_r = u_r . tau
  const float k = (float)kK, inv_m = (float)(1.0 / kMass), inv_mD0 = (float)(1.0 / (kMass * kD0));
  float ux[kMaxWallRows], uy[kMaxWallRows], rhs[kMaxWallRows], reg[kMaxWallRows], f[kMaxWallRows];
  const float b = (float)(2.0 / (kSolDmax * kSolTc));
  for (int r = 0; r < rows; ++r) {
    const float ut = k * (s * nx[r] - c * ny[r]) * inv_mD0;
    ux[r] = (nx[r] + k * s * ut) * inv_m;
    uy[r] = (ny[r] - k * c * ut) * inv_m;
    const float a0 = ux[r] * tx + uy[r] * ty + ut * tth;
    const float d = wall_impedance(dist[r]);
    const float kk = d * (float)(1.0 / (kSolDmax * kSolDmax * kSolTc * kSolTc));
    const float aref = -b * (nx[r] * vx + ny[r] * vy) - kk * dist[r];
    rhs[r] = aref - a0;
    reg[r] = (1.f - d) / d * (nx[r] * ux[r] + ny[r] * uy[r]);
    f[r] = 0.f;
  }
  // projected Gauss-Seidel on  sum_q mult_q A_rq f_q + reg_r f_r = rhs_r,  f >= 0  (f = the force of ONE contact of row r)
#pragma unroll 1
  for (int it = 0; it < 200; ++it) {
    float delta = 0.f, scale = 0.f;
    for (int r = 0; r < rows; ++r) {
      float acc = reg[r] * f[r] - rhs[r];
      for (int q = 0; q < rows; ++q) acc += mult[q] * (nx[r] * ux[q] + ny[r] * uy[q]) * f[q];
      const float arr = mult[r] * (nx[r] * ux[r] + ny[r] * uy[r]) + reg[r];
      const float fr = fmaxf(0.f, f[r] - acc / arr);
      delta = fmaxf(delta, fabsf(fr - f[r]));
      scale = fmaxf(scale, fr);
      f[r] = fr;
    }
    if (delta <= 1e-7f * scale) break;
  }
  for (int r = 0; r < rows; ++r) {
    out.fx += mult[r] * f[r] * nx[r];
    out.fy += mult[r] * f[r] * ny[r];
  }
  return out;
}

// n MuJoCo substeps with constant ctrl.  (c, s) = (cos phi, sin phi) is carried in
// registers and advanced by the exact rotation of h*w each substep (|h w| < 0.01, so
// a 5th-order series is exact to fp32), instead of n full-range sincosf calls.
// Returns the final (c, s), renormalised, for the observation.
template <int CONTACT = CRL_CONTACT_MODEL, bool WALLS = false>
CRL_HD void substeps(Body& b, float a0, float a1, int n, float& c_out, float& s_out, float extent = 3.f) {
  const float h = (float)kH;
  const float bl = (float)kDampLin, bt = (float)kDampYaw, g = (float)kGear;
  const float k = (float)kK, inv_m = (float)(1.0 / kMp), inv_D = (float)(1.0 / kD);
  const float kb_m = (float)(kK * kDampLin / kMp);
  const float Fm = g * clampf(clampf(a0, -1.f, 1.f), -(float)kForceLimit, (float)kForceLimit);
  const float u1 = clampf(a1, -1.f, 1.f);
  float s, c;
  sincosf(b.phi, &s, &c);
  float X = b.X, Y = b.Y, phi = b.phi, vx = b.vx, vy = b.vy, w = b.w;
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const float servo = clampf(u1 - g * w, -(float)kForceLimit, (float)kForceLimit);
    const float tau_th = g * servo - bt * w;
    const float a_th = (tau_th - kb_m * (s * vx - c * vy)) * inv_D;
    const float drive = k * w * w + Fm;           // centrifugal + motor, along heading
    const float tau_x = drive * c - bl * vx;
    const float tau_y = drive * s - bl * vy;
    const float ka = k * a_th;
    float ax = (tau_x + ka * s) * inv_m;
    float ay = (tau_y - ka * c) * inv_m;
    float a_th_c = a_th;
    constraint_acc<CONTACT>(ax, ay, a_th_c, vx, vy, w);   // model 0: nothing (compiled out)
    if (WALLS) {
      // a wall can only be touched from within one radius of its inner face
      if (fmaxf(fabsf(X), fabsf(Y)) > extent - (float)(kWallHalf + kSphereR)) {
        const WallForce wf = wall_force(X, Y, c, s, vx, vy, tau_x, tau_y, tau_th, extent);
        // a += (M + hB)^-1 (Fx, Fy, 0), the same closed form as above
        const float ut = k * (s * wf.fx - c * wf.fy) * inv_m * inv_D;
        ax += (wf.fx + k * s * ut) * inv_m;
        ay += (wf.fy - k * c * ut) * inv_m;
        a_th_c += ut;
      }
    }
    vx += h * ax;
    vy += h * ay;
    w += h * a_th_c;
    X += h * vx;
    Y += h * vy;
    const float d = h * w;
    phi += d;
    const float d2 = d * d;
    const float sd = d * (1.f + d2 * (-1.f / 6.f + d2 * (1.f / 120.f)));
    const float cd = 1.f + d2 * (-0.5f + d2 * (1.f / 24.f));
    const float cn = c * cd - s * sd;
    s = s * cd + c * sd;
    c = cn;
  }
  b.X = X; b.Y = Y; b.phi = phi; b.vx = vx; b.vy = vy; b.w = w;
  // c^2 + s^2 = 1 + e with |e| ~ 1e-6 after n rotations: 1/sqrt(1+e) = 1.5 - 0.5(1+e) + O(e^2)
  const float r = 1.5f - 0.5f * (c * c + s * s);
  c_out = c * r;
  s_out = s * r;
}

// heading kept in [-pi, pi]: one conditional suffices (|dphi| per env step << pi).
CRL_HD float wrap_pi(float phi) {
  const float two_pi_hi = 6.2831855f;             // fl32(2 pi)
  const float two_pi_lo = -1.7484555e-07f;        // 2 pi - fl32(2 pi)
  const float pi = 3.14159265f;
  if (phi > pi) phi = (phi - two_pi_hi) - two_pi_lo;
  else if (phi < -pi) phi = (phi + two_pi_hi) + two_pi_lo;
  return phi;
}

// Largest double t with sqrt(t) <= r: then  sqrt(d2) <= r  <=>  d2 <= t  exactly,
// because IEEE sqrt is correctly rounded and monotone.  Host only.
inline double sqrt_threshold(double r) {
  double t = r * r;
  while (sqrt(nextafter(t, INFINITY)) <= r) t = nextafter(t, INFINITY);
  while (sqrt(t) > r) t = nextafter(t, -INFINITY);
  return t;
}

// n / d correctly rounded (== IEEE float division) for integer-valued n and the two
// divisors the observation uses (num_steps, max_cooldown): q0 = RN(n y), y = RN(1/d),
// one exact-remainder correction (Markstein).  div_const_ok() proves it on the host, for
// the divisor in force, over every numerator the kernels can form; a configuration whose
// divisor failed the proof is refused (CRL_ERR_CONFIG) -- none does: every divisor in
// [1, 65534] passes for every numerator in [-65535, 65535] (checked exhaustively), so the
// device path carries no fallback branch.
struct DivConst { float d, y; int exact; };

CRL_HD float div_const(float n, const DivConst& k) {
#if defined(__CUDA_ARCH__)
  const float q0 = __fmul_rn(n, k.y);
  return __fmaf_rn(__fmaf_rn(-q0, k.d, n), k.y, q0);
#else
  if (!k.exact) return n / k.d;
  const float q0 = n * k.y;
  return fmaf(fmaf(-q0, k.d, n), k.y, q0);
#endif
}

inline DivConst make_div_const(int d, int n_lo, int n_hi) {
  DivConst k{(float)d, 1.0f / (float)d, 1};
  for (int n = n_lo; n <= n_hi && k.exact; ++n) {
    volatile float fn = (float)n, fd = (float)d;
    const float want = fn / fd;                    // IEEE division
    const float q0 = (float)n * k.y;
    const float got = fmaf(fmaf(-q0, k.d, (float)n), k.y, q0);
    if (!(got == want)) k.exact = 0;
  }
  return k;
}

// numpy: sqrt(sum(square(zone - robot))) <= size.  fl(fl(dx*dx) + fl(dy*dy)) with no
// fused multiply-add, compared against the pre-searched squared threshold.
CRL_HD bool inside_zone(float X, float Y, float zx, float zy, double thresh2) {
  const double dx = (double)zx - (double)X;
  const double dy = (double)zy - (double)Y;
#if defined(__CUDA_ARCH__)
  const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
#else
  volatile double xx = dx * dx;
  volatile double yy = dy * dy;
  const double d2 = xx + yy;
#endif
  return d2 <= thresh2;
}

// fp32 screen: true if the exact test could possibly pass (d2 within 1e-4 of r^2).
CRL_HD bool near_zone(float X, float Y, float zx, float zy, float r2_guard) {
  const float dx = zx - X, dy = zy - Y;
  return dx * dx + dy * dy <= r2_guard;
}

// colour_match_env.py:38-55 on 2-bit colour codes packed from bit 0 (0 B, 1 G, 2 R).
CRL_HD int popcount32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

CRL_HD int hamming(uint32_t colours, int n) {
  const uint32_t field = n >= 16 ? 0x55555555u : ((1u << (2 * n)) - 1u) & 0x55555555u;
  const uint32_t lo = colours & field, hi = (colours >> 1) & field;
  const int ng = popcount32(lo & ~hi), nr = popcount32(hi & ~lo);
  const int nb = popcount32(field & ~(lo | hi));
  const int to_b = 2 * ng + nr, to_g = 2 * nr + nb, to_r = 2 * nb + ng;
  int m = to_b < to_g ? to_b : to_g;
  return m < to_r ? m : to_r;
}

// ---- Philox4x32-10 (Salmon et al. 2011, the Random123 constants) -----------
struct U4 { uint32_t x, y, z, w; };

CRL_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  hi = __umulhi(a, b);
  lo = a * b;
#else
  const uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

CRL_HD U4 philox4x32(U4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mulhilo(0xD2511F53u, ctr.x, h0, l0);
    mulhilo(0xCD9E8D57u, ctr.z, h1, l1);
    U4 n;
    n.x = h1 ^ ctr.y ^ k0;
    n.y = l1;
    n.z = h0 ^ ctr.w ^ k1;
    n.w = l0;
    ctr = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}

// 24-bit uniform in [0, 1)
CRL_HD float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// 53-bit uniform in (0, 1]: safe under log()
CRL_HD double u01d(uint32_t hi, uint32_t lo) {
  const uint64_t v = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
  return ((double)v + 1.0) * (1.0 / 9007199254740992.0);
}

// counter word 3 tags: which draw family a Philox block belongs to
enum : uint32_t { kTagLayout = 1, kTagRot = 2, kTagTask = 3, kTagSeed = 4, kTagAction = 5 };

}  // namespace crl
