// g++ build of combinatorial_rl_tasks_b200/csrc/crl_core.cuh for CPU-side checks of the
// per-env arithmetic (closed-form substep, exact zone predicate, Hamming, Philox)
// against the oracle.  Test aid only: never loaded by the product package.
#include "../../combinatorial_rl_tasks_b200/csrc/crl_core.cuh"

extern "C" {

// state: X, Y, phi, vx, vy, w (world frame); returns cos/sin too
void hc_substeps(float* st, float a0, float a1, int n, float* cs) {
  crl::Body b{st[0], st[1], st[2], st[3], st[4], st[5]};
  crl::substeps(b, a0, a1, n, cs[0], cs[1]);
  st[0] = b.X; st[1] = b.Y; st[2] = b.phi; st[3] = b.vx; st[4] = b.vy; st[5] = b.w;
}
// the same with the ALTERNATIVE contact reading (crl_core.cuh constraint_acc<1>; never what the product compiles)
void hc_substeps_contact_active(float* st, float a0, float a1, int n, float* cs) {
  crl::Body b{st[0], st[1], st[2], st[3], st[4], st[5]};
  crl::substeps<1>(b, a0, a1, n, cs[0], cs[1]);
  st[0] = b.X; st[1] = b.Y; st[2] = b.phi; st[3] = b.vx; st[4] = b.vy; st[5] = b.w;
}
// substeps with the walls of a `walled=True` env (crl_core.cuh::wall_force), extent 3
void hc_substeps_walled(float* st, float a0, float a1, int n, float* cs) {
  crl::Body b{st[0], st[1], st[2], st[3], st[4], st[5]};
  crl::substeps<CRL_CONTACT_MODEL, true>(b, a0, a1, n, cs[0], cs[1], 3.f);
  st[0] = b.X; st[1] = b.Y; st[2] = b.phi; st[3] = b.vx; st[4] = b.vy; st[5] = b.w;
}
float hc_wrap_pi(float phi) { return crl::wrap_pi(phi); }
double hc_sqrt_threshold(double r) { return crl::sqrt_threshold(r); }
int hc_inside_zone(float X, float Y, float zx, float zy, double t2) { return crl::inside_zone(X, Y, zx, zy, t2); }
int hc_hamming(uint32_t colours, int n) { return crl::hamming(colours, n); }
// div_const over [n_lo, n_hi]: returns the host proof's verdict; out[i] = n / d as the kernel forms it
int hc_div_const(int d, int n_lo, int n_hi, float* out) {
  const crl::DivConst k = crl::make_div_const(d, n_lo, n_hi);
  for (int n = n_lo; n <= n_hi; ++n) out[n - n_lo] = crl::div_const((float)n, k);
  return k.exact;
}
void hc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  crl::U4 r = crl::philox4x32(crl::U4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
