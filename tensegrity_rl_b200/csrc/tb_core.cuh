// tb_core.cuh -- bar-lane tensegrity physics for sm_100a: THREE LANES PER ENV (one per bar), TEN ENVS PER WARP.
//
// What is computed is MuJoCo 2.3.7's mj_forward / mj_step for the 3-bar model (the reference reaches
// it through gym's MujocoEnv.do_simulation, tr_env.py:346,812, tensegrity_env.py:297): free-joint kinematics ->
// 9 two-site spatial tendons -> collision (plane / height field / bar-bar, MPR where MuJoCo uses libccd) -> condim-6
// elliptic contact rows -> Newton with exact line search -> implicitfast -> advance.
//
// Mapping.  A bar's 13-number pose/velocity, its warm start, its 6x6 Newton block and every per-bar vector of the
// solver live in the REGISTERS of the bar's lane for all frame_skip substeps; lanes exchange only what crosses bars
// (tendon end points, world-frame twists, contact wrenches) through a ~1.9 KB per-env slice of shared memory and
// 3-lane shuffle sums.  Contacts are owned by a lane (floor contacts by the bar's lane, bar-bar contacts by lane 2 of
// the env) and kept in that lane's local memory, which the L1 caches at its actual use instead of a static worst-case
// shared-memory reservation.  Control flow is WARP-UNIFORM: every data-dependent loop (Newton iterations, line-search
// evaluations, second forward pass after a bad acceleration) runs while ANY env of the warp needs it, with the other
// envs predicated off, so the 10 envs of a warp share every fetched instruction and never need a partial barrier.
#pragma once
#include "tb_mpr.h"

namespace tb {

constexpr int G = 3;        // lanes per env
constexpr int EPW = 10;     // envs per warp (lanes 30, 31 idle)
constexpr int MAXCL = 16;   // contacts one lane can own (a bar lying flat on the floor: 14; local memory, cached as used)
constexpr int KHAND = 3;    // bar-bar contacts per bar pair handed to the env's lane 2 per round
constexpr int MAXH = 12;    // bar-bar contacts one pair can have
constexpr int ZONE_TOP = 0, ZONE_BOTTOM = 1, ZONE_MIDDLE = 2;

template <typename real>
struct Con {
  real frame[9];
  real r1[3], r2[3];   // contact point relative to the centre of body 1 / body 2 (world axes)
  real aref[6], jar[6], jv[6], force[6];
  real su[6];
  real D0, wcoef, ca, cb;
  real U0, V0, UU, UV, VV, q0, q1, q2;
  int b1, b2;          // bar index 0..2, or -1 for the world (b1 only)
  int zone, pad;
};

template <typename real> struct HandCon { real dist, pos[3], nrm[3]; int b1, b2; };

// per-env slice of shared memory
template <typename real>
struct EnvSh {
  // home of the env state between physics calls (always double: the HBM record's precision)
  double qpos[NQ], qvel[NV], warm[NV], ctrl[NACT], act[NACT];
  // world-frame exchange, rewritten by every pass; xpos / xmat / sph / tlen double as the "stale" kinematics the
  // reference's observation reads (positions lag qpos by one substep)
  real xpos[9], xmat[27], vw[18], sph[18], tlen[NTEN], actdot[NACT];
  union {
    real site[NEND * 3];                       // tendon end points (tendon stage)
    HandCon<real> hand[3][KHAND];              // bar-bar contacts found by the pair lanes (collision stage)
    struct { real xv[NV], fx[NV], Hg[2 * 21]; } sol;  // solver: vector under J, wrenches of lane-2 contacts, blocks 0/1
    real cfrc[24];                             // after the last pass
  } u;
  int nhand[3];
  int nact, overflow, bad, niter, nls, nmpr;
  real barforce;
  // env layer (lane 0 of the env)
  double action[NACT], draws[NDRAW + 2];
};

struct LaneCtx { int lane, grp, bar, base; bool valid; };
TB_FN LaneCtx make_lane() {
  LaneCtx L;
  L.lane = simt_lane();
  L.valid = L.lane < G * EPW;
  L.grp = L.valid ? L.lane / G : EPW - 1;
  L.bar = L.valid ? L.lane % G : L.lane - G * EPW;
  L.base = L.grp * G;
  return L;
}
template <typename real> TB_FN real sum3(real v, int base) {
  real a = shfl(v, base), b = shfl(v, base + 1), c = shfl(v, base + 2);
  return (a + b) + c;
}
TB_FN int isum3(int v, int base) { return shfl(v, base) + shfl(v, base + 1) + shfl(v, base + 2); }
TB_FN bool grp_any(bool p, int base) { return ((ballot(p) >> base) & 7u) != 0; }

template <typename real> struct BarState { real x[3], q[4], v[6], warm[6]; };

// ------------------------------------------------------------------ contact helpers
template <typename real> TB_FN void make_frame(real* f) {
  normalize3(f);
  f[3] = f[4] = f[5] = 0;
  if (f[1] < real(0.5) && f[1] > real(-0.5)) f[4] = 1; else f[5] = 1;
  real t = dot3(f, f + 3);
  addscl3(f + 3, f, -t);
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
// out = J x, x given as world-frame twists per bar (lin, world angular) in xv
template <typename real> TB_FN void con_mulJ(const Con<real>& c, const real* xv, real* out) {
  real rel[3], relw[3], t[3];
  const real* w2 = xv + 6 * c.b2 + 3;
  cross3(t, w2, c.r2);
  add3(rel, xv + 6 * c.b2, t);
  copy3(relw, w2);
  if (c.b1 >= 0) {
    const real* w1 = xv + 6 * c.b1 + 3;
    real u1[3];
    cross3(t, w1, c.r1);
    add3(u1, xv + 6 * c.b1, t);
    sub3(rel, rel, u1); sub3(relw, relw, w1);
  }
  for (int a = 0; a < 3; a++) { out[a] = dot3(c.frame + 3 * a, rel); out[3 + a] = dot3(c.frame + 3 * a, relw); }
}
// world wrench (force F, torque T) of the contact force: side 2 receives (F, r2 x F + T), side 1 the opposite about r1
template <typename real> TB_FN void con_wrench(const Con<real>& c, real* F, real* T) {
  for (int k = 0; k < 3; k++) {
    F[k] = c.frame[k] * c.force[0] + c.frame[3 + k] * c.force[1] + c.frame[6 + k] * c.force[2];
    T[k] = c.frame[k] * c.force[3] + c.frame[3 + k] * c.force[4] + c.frame[6 + k] * c.force[5];
  }
}
template <typename real> TB_FN real impedance(const ModelT<real>& m, real pos) {
  const real MINIMP = real(0.0001), MAXIMP = real(0.9999);
  real d0 = clampr(m.solimp[0], MINIMP, MAXIMP), dw = clampr(m.solimp[1], MINIMP, MAXIMP);
  real width = tmax(Lim<real>::MINVAL, m.solimp[2]), mid = clampr(m.solimp[3], MINIMP, MAXIMP), power = tmax(real(1), m.solimp[4]);
  if (d0 == dw || width <= Lim<real>::MINVAL) return real(0.5) * (d0 + dw);
  real x = tdiv(tabs(pos), width), y;
  if (x >= 1) return dw;
  if (x == 0) return d0;
  if (power == 1) y = x;
  else if (power == 2) y = (x <= mid) ? trcp(mid) * (x * x) : 1 - trcp(1 - mid) * ((1 - x) * (1 - x));
  else if (x <= mid) y = (real)((1 / pow((double)mid, (double)power - 1)) * pow((double)x, (double)power));
  else y = (real)(1 - (1 / pow(1 - (double)mid, (double)power - 1)) * pow(1 - (double)x, (double)power));
  return d0 + y * (dw - d0);
}
// mj_constraintUpdate for one elliptic contact at c.jar: returns its cost; full: also force, zone, Hessian weights
template <typename real> TB_FN real con_update(Con<real>& c, const ModelT<real>& m, bool full) {
  real U[6], T = 0, mu = m.mu;
  U[0] = c.jar[0] * mu;
  for (int j = 1; j < 6; j++) { U[j] = c.jar[j] * m.fr[j - 1]; T += U[j] * U[j]; }
  real N = U[0];
  T = tsqrt(T);
  if (N >= mu * T || (T <= 0 && N >= 0)) {
    if (full) { for (int j = 0; j < 6; j++) c.force[j] = 0; c.zone = ZONE_TOP; c.wcoef = c.ca = c.cb = 0; }
    return 0;
  }
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
    real s = 0;
    for (int j = 0; j < 6; j++) {
      real D = c.D0 * m.dscale[j];
      s += real(0.5) * D * c.jar[j] * c.jar[j];
      if (full) c.force[j] = -D * c.jar[j];
    }
    if (full) { c.zone = ZONE_BOTTOM; c.wcoef = c.D0; c.ca = c.cb = 0; }
    return s;
  }
  real Dm = c.D0 * m.inv_mu2, NT = N - mu * T;
  if (full) {
    real invT = trcp(T);
    c.force[0] = -Dm * NT * mu;
    real kap = mu * mu - mu * N * invT;
    c.su[0] = 0;
    for (int j = 1; j < 6; j++) {
      c.force[j] = -c.force[0] * invT * U[j] * m.fr[j - 1];
      c.su[j] = m.fr[j - 1] * U[j] * invT;
    }
    c.ca = Dm; c.cb = Dm * kap; c.wcoef = c.cb; c.zone = ZONE_MIDDLE;
  }
  return real(0.5) * Dm * NT * NT;
}
// cost and its first two derivatives along the search direction at step a, for one contact
template <typename real> TB_FN void con_ls(const Con<real>& k, const ModelT<real>& m, real a, real& cost, real& d0, real& d1) {
  real mu = m.mu;
  real N = k.U0 + a * k.V0, Tsqr = k.UU + a * (2 * k.UV + a * k.VV);
  bool bottom = false;
  if (Tsqr <= 0) { if (N < 0) bottom = true; }
  else {
    real T = tsqrt(Tsqr);
    if (N >= mu * T) {}
    else if (mu * N + T <= 0) bottom = true;
    else {
      real invT = trcp(T);
      real N1 = k.V0, T1 = (k.UV + a * k.VV) * invT;
      real T2 = k.VV * invT - (k.UV + a * k.VV) * T1 * (invT * invT);
      real NT = N - mu * T, Dm = k.D0 * m.inv_mu2;
      cost += real(0.5) * Dm * NT * NT;
      d0 += Dm * NT * (N1 - mu * T1);
      d1 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NT * (-mu * T2));
    }
  }
  if (bottom) { cost += a * a * k.q2 + a * k.q1 + k.q0; d0 += 2 * a * k.q2 + k.q1; d1 += 2 * k.q2; }
}

// rows of one side's 6x6 Jacobian block in the bar's own dof coordinates (lin world, ang body-local):
// translational row a = (s f_a, s R^T (r x f_a)), rotational row a = (0, s R^T f_a)
template <typename real> struct SideJ { real lin[3][3], ang[3][3], rot[3][3]; };
template <typename real> TB_FN void side_rows(const Con<real>& c, real s, const real* r, const real* R, SideJ<real>& J) {
  for (int a = 0; a < 3; a++) {
    const real* f = c.frame + 3 * a;
    real t[3], w[3];
    cross3(t, r, f);
    mulMTV(w, R, t);
    for (int k = 0; k < 3; k++) { J.lin[a][k] = s * f[k]; J.ang[a][k] = s * w[k]; }
    mulMTV(w, R, f);
    for (int k = 0; k < 3; k++) J.rot[a][k] = s * w[k];
  }
}
// 6-vector of row r of a side
template <typename real> TB_FN void side_row6(const SideJ<real>& J, int r, real* j) {
  if (r < 3) { for (int k = 0; k < 3; k++) { j[k] = J.lin[r][k]; j[3 + k] = J.ang[r][k]; } }
  else { for (int k = 0; k < 3; k++) { j[k] = 0; j[3 + k] = J.rot[r - 3][k]; } }
}
// H (packed lower 6x6) += w j j^T
template <typename real> TB_FN void rank1_sym(real* H, real w, const real* j) {
  TB_UNROLL
  for (int i = 0, e = 0; i < 6; i++) {
    real wi = w * j[i];
    TB_UNROLL
    for (int k = 0; k <= i; k++, e++) H[e] += wi * j[k];
  }
}
// X (6x6 row-major) += w a b^T
template <typename real> TB_FN void rank1_gen(real* X, real w, const real* a, const real* b) {
  TB_UNROLL
  for (int i = 0; i < 6; i++) {
    real wi = w * a[i];
    TB_UNROLL
    for (int k = 0; k < 6; k++) X[6 * i + k] += wi * b[k];
  }
}
// cone vectors of a middle-zone contact side: bvec = sum_j su_j J_j, avec = mu (J_0 - bvec)
template <typename real> TB_FN void side_cone(const Con<real>& c, const ModelT<real>& m, const SideJ<real>& J, real* avec, real* bvec) {
  for (int k = 0; k < 6; k++) bvec[k] = 0;
  for (int r = 1; r < 6; r++) {
    real j[6];
    side_row6(J, r, j);
    for (int k = 0; k < 6; k++) bvec[k] += c.su[r] * j[k];
  }
  real j0[6];
  side_row6(J, 0, j0);
  for (int k = 0; k < 6; k++) avec[k] = m.mu * (j0[k] - bvec[k]);
}
// diagonal-block contribution of one contact side
template <typename real> TB_FN void side_hessian(const Con<real>& c, const ModelT<real>& m, const SideJ<real>& J, real* H) {
  const real* wt = m.wtab[c.zone == ZONE_MIDDLE ? 1 : 0];
  for (int r = 0; r < 6; r++) {
    real w = c.wcoef * wt[r];
    if (w != 0) { real j[6]; side_row6(J, r, j); rank1_sym(H, w, j); }
  }
  if (c.zone == ZONE_MIDDLE) {
    real a[6], b[6];
    side_cone(c, m, J, a, b);
    rank1_sym(H, c.ca, a);
    rank1_sym(H, -c.cb, b);
  }
}

// ---- 6x6 block kernels of the Newton solve (packed lower triangles: entry (i, k) at i (i + 1) / 2 + k; full blocks
// row-major).  Everything is unrolled over static indices, so blocks that are plain local variables stay in registers.
// A = L D L^T in place: strict lower part = unit L, diagonal = d; dinv = 1 / d
template <typename real> TB_FN void blk_ldl(real* A, real* dinv) {
  TB_UNROLL
  for (int j = 0; j < 6; j++) {
    real s = A[j * (j + 1) / 2 + j];
    TB_UNROLL
    for (int k = 0; k < j; k++) {
      real t = A[j * (j + 1) / 2 + k];
      TB_UNROLL
      for (int q = 0; q < k; q++) t -= A[j * (j + 1) / 2 + q] * A[q * (q + 1) / 2 + q] * A[k * (k + 1) / 2 + q];
      real l = t * dinv[k];
      A[j * (j + 1) / 2 + k] = l;
      s -= l * t;
    }
    s = tmax(s, Lim<real>::MINVAL);
    A[j * (j + 1) / 2 + j] = s;
    dinv[j] = trcp(s);
  }
}
template <typename real> TB_FN void blk_fwd(const real* L, real* x) {   // x <- L^-1 x
  TB_UNROLL
  for (int i = 1; i < 6; i++) {
    real t = x[i];
    TB_UNROLL
    for (int k = 0; k < i; k++) t -= L[i * (i + 1) / 2 + k] * x[k];
    x[i] = t;
  }
}
template <typename real> TB_FN void blk_bwd(const real* L, real* x) {   // x <- L^-T x
  TB_UNROLL
  for (int i = 4; i >= 0; i--) {
    real t = x[i];
    TB_UNROLL
    for (int k = i + 1; k < 6; k++) t -= L[k * (k + 1) / 2 + i] * x[k];
    x[i] = t;
  }
}
// X <- X L^-T D^-1   (X a full 6x6 block below the diagonal block L D L^T)
template <typename real> TB_FN void blk_trsm(real* X, const real* L, const real* dinv) {
  TB_UNROLL
  for (int r = 0; r < 6; r++) {
    real u[6];
    TB_UNROLL
    for (int k = 0; k < 6; k++) {
      real t = X[6 * r + k];
      TB_UNROLL
      for (int q = 0; q < k; q++) t -= u[q] * L[k * (k + 1) / 2 + q];
      u[k] = t;
    }
    TB_UNROLL
    for (int k = 0; k < 6; k++) X[6 * r + k] = u[k] * dinv[k];
  }
}
// C (packed) -= A diag(d) A^T
template <typename real> TB_FN void blk_syrk(real* C, const real* A, const real* Ld) {
  TB_UNROLL
  for (int i = 0; i < 6; i++) {
    real ad[6];
    TB_UNROLL
    for (int k = 0; k < 6; k++) ad[k] = A[6 * i + k] * Ld[k * (k + 1) / 2 + k];
    TB_UNROLL
    for (int j = 0; j <= i; j++) {
      real sacc = 0;
      TB_UNROLL
      for (int k = 0; k < 6; k++) sacc += ad[k] * A[6 * j + k];
      C[i * (i + 1) / 2 + j] -= sacc;
    }
  }
}
// C (full) -= A diag(d) B^T
template <typename real> TB_FN void blk_gemm(real* C, const real* A, const real* Ld, const real* B) {
  TB_UNROLL
  for (int i = 0; i < 6; i++) {
    real ad[6];
    TB_UNROLL
    for (int k = 0; k < 6; k++) ad[k] = A[6 * i + k] * Ld[k * (k + 1) / 2 + k];
    TB_UNROLL
    for (int j = 0; j < 6; j++) {
      real sacc = 0;
      TB_UNROLL
      for (int k = 0; k < 6; k++) sacc += ad[k] * B[6 * j + k];
      C[6 * i + j] -= sacc;
    }
  }
}
template <typename real> TB_FN void blk_gemv_sub(real* y, const real* A, const real* x) {   // y -= A x
  TB_UNROLL
  for (int i = 0; i < 6; i++) {
    real t = y[i];
    TB_UNROLL
    for (int k = 0; k < 6; k++) t -= A[6 * i + k] * x[k];
    y[i] = t;
  }
}
template <typename real> TB_FN void blk_gemvT_sub(real* y, const real* A, const real* x) {   // y -= A^T x
  TB_UNROLL
  for (int i = 0; i < 6; i++) {
    real t = y[i];
    TB_UNROLL
    for (int k = 0; k < 6; k++) t -= A[6 * k + i] * x[k];
    y[i] = t;
  }
}

// ------------------------------------------------------------------ collision pieces
template <typename real>
TB_NOINL void plane_cylinder_points(const ModelT<real>& m, const real* pos2, const real* axis_in, real radius, real half,
                                    const real* xaxis, int& cnt, real dist[4], real pts[4][3]) {
  const real* normal = m.fnormal;
  real axis[3], vec[3];
  copy3(axis, axis_in);
  real prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { scl3(axis, axis, real(-1)); prjaxis = -prjaxis; }
  sub3(vec, pos2, m.fpos);
  real dist0 = dot3(vec, normal);
  scl3(vec, axis, prjaxis); sub3(vec, vec, normal);
  real len_sqr = dot3(vec, vec);
  if (len_sqr >= Lim<real>::MINVAL * Lim<real>::MINVAL) scl3(vec, vec, radius / tsqrt(len_sqr));
  else scl3(vec, xaxis, radius);
  real prjvec = dot3(vec, normal);
  scl3(axis, axis, half); prjaxis *= half;
  cnt = 0;
  if (dist0 + prjaxis + prjvec <= 0) {
    dist[cnt] = dist0 + prjaxis + prjvec;
    add3(pts[cnt], pos2, vec); add3(pts[cnt], pts[cnt], axis); addscl3(pts[cnt], normal, -dist[cnt] * real(0.5));
    cnt++;
  } else return;
  if (dist0 - prjaxis + prjvec <= 0) {
    dist[cnt] = dist0 - prjaxis + prjvec;
    add3(pts[cnt], pos2, vec); sub3(pts[cnt], pts[cnt], axis); addscl3(pts[cnt], normal, -dist[cnt] * real(0.5));
    cnt++;
  }
  real prjvec1 = -prjvec * real(0.5);
  if (dist0 + prjaxis + prjvec1 <= 0) {
    real vec1[3];
    cross3(vec1, vec, axis); normalize3(vec1); scl3(vec1, vec1, radius * tsqrt(real(3)) * real(0.5));
    for (int s = 0; s < 2; s++) {
      dist[cnt] = dist0 + prjaxis + prjvec1;
      add3(pts[cnt], pos2, axis); addscl3(pts[cnt], vec1, s ? real(-1) : real(1)); addscl3(pts[cnt], vec, real(-0.5));
      addscl3(pts[cnt], normal, -dist[cnt] * real(0.5));
      cnt++;
    }
  }
}
// squared distance between segments p1 +- a1, p2 +- a2 (a = half-axis vectors)
template <typename real> TB_FN real segseg_dist2(const real* p1, const real* a1, const real* p2, const real* a2) {
  real r[3]; sub3(r, p1, p2);
  real A = dot3(a1, a1), E = dot3(a2, a2), Bq = dot3(a1, a2), C = dot3(a1, r), F = dot3(a2, r);
  real den = A * E - Bq * Bq, s = 0, t;
  if (den > real(1e-30)) s = clampr(tdiv(Bq * F - C * E, den), real(-1), real(1));
  t = tdiv(Bq * s + F, E);
  if (t < real(-1)) { t = real(-1); s = clampr(tdiv(-Bq - C, A), real(-1), real(1)); }
  else if (t > real(1)) { t = real(1); s = clampr(tdiv(Bq - C, A), real(-1), real(1)); }
  real d[3] = {r[0] + s * a1[0] - t * a2[0], r[1] + s * a1[1] - t * a2[1], r[2] + s * a1[2] - t * a2[2]};
  return dot3(d, d);
}
template <typename real> TB_FN real ptseg_dist2(const real* c, const real* p, const real* a) {
  real r[3]; sub3(r, c, p);
  real t = clampr(tdiv(dot3(r, a), dot3(a, a)), real(-1), real(1));
  addscl3(r, a, -t);
  return dot3(r, r);
}

// appends an active contact to the lane's list (dist >= 0 gives no rows: includemargin 0)
template <typename real>
TB_FN void add_contact(Con<real>* con, int& ncon, int& overflow, int b1, int b2, real dist, const real* pos, const real* normal,
                       const real* xpos) {
  if (!(dist < 0)) return;
  if (ncon >= MAXCL) { overflow = 1; return; }
  Con<real>& c = con[ncon++];
  c.b1 = b1; c.b2 = b2;
  copy3(c.frame, normal);
  make_frame(c.frame);
  sub3(c.r2, pos, xpos + 3 * b2);
  if (b1 >= 0) sub3(c.r1, pos, xpos + 3 * b1); else { c.r1[0] = c.r1[1] = c.r1[2] = 0; }
  c.aref[0] = dist;   // parked until the rows are built
}

// height-field prism k of row r (vertices n = k, k+1, k+2 of the strip; c = cmin + n/2, odd n -> row r, even -> r + 1)
template <typename real> TB_FN void hf_prism(const ModelT<real>& m, int r, int cmin, int k, CObj<real>& o) {
  for (int j = 0; j < 3; j++) {
    int n = k + j, c = cmin + n / 2, rr = r + ((n & 1) ? 0 : 1);
    o.px[j] = m.hdx * c - m.hsize[0]; o.py[j] = m.hdy * rr - m.hsize[1];
    o.pz[j] = (real)m.hdata[rr * m.ncol + c] * m.hsize[2];
  }
  o.pbase = -m.hsize[3];
}
// Conservative cull before MPR: a prism lies on or below the plane through its three top vertices, so a geom whose
// support point towards that plane stays above it (by more than a rounding slack) cannot touch the prism.
template <typename real>
TB_FN bool hf_above_top_plane(const CObj<real>& pr, int gtype, const real* pos, const real* R, real rad, real hl) {
  real e1[3] = {pr.px[1] - pr.px[0], pr.py[1] - pr.py[0], pr.pz[1] - pr.pz[0]};
  real e2[3] = {pr.px[2] - pr.px[0], pr.py[2] - pr.py[0], pr.pz[2] - pr.pz[0]};
  real n[3];
  cross3(n, e1, e2);
  if (n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  real nn = dot3(n, n);
  if (!(nn > 0)) return false;
  real reach;
  if (gtype == GEOM_SPHERE) reach = rad * tsqrt(nn);
  else {
    real na = n[0] * R[2] + n[1] * R[5] + n[2] * R[8];
    reach = hl * tabs(na) + rad * tsqrt(tmax(real(0), nn - na * na));
  }
  real top0[3] = {pr.px[0], pr.py[0], pr.pz[0]};
  real sep = dot3(n, pos) - reach - dot3(n, top0);
  return sep > real(1e-9) * tsqrt(nn);
}

// ------------------------------------------------------------------ one physics pass
// mj_forward (integ = false) or mj_step (integ = true) for the envs of the warp that are `on`.  The bar state B is
// the lane's registers; con / ncon (lane-local) hold the contacts of the last pass on return.
template <typename real>
TB_FN void phys(BarState<real>& B, EnvSh<real>& S, const ModelT<real>& m, const LaneCtx& L, bool on, bool integ,
                Con<real>* con, int& ncon) {
  const int b = L.bar, base = L.base;
  const real MINV = Lim<real>::MINVAL;
  auto reset_data = [&]() {   // mj_resetData on this lane's bar
    for (int k = 0; k < 3; k++) B.x[k] = m.qpos0[7 * b + k];
    for (int k = 0; k < 4; k++) B.q[k] = m.qpos0[7 * b + 3 + k];
    for (int k = 0; k < 6; k++) { B.v[k] = 0; B.warm[k] = 0; }
    if (b == 0) for (int k = 0; k < NACT; k++) { S.ctrl[k] = 0; S.act[k] = 0; }
  };
  if (integ) {   // mj_checkPos / mj_checkVel
    bool bad = false;
    for (int k = 0; k < 3; k++) bad |= is_bad(B.x[k]);
    for (int k = 0; k < 4; k++) bad |= is_bad(B.q[k]);
    for (int k = 0; k < 6; k++) bad |= is_bad(B.v[k]);
    bad = grp_any(bad && on, base);
    if (bad && on) { if (b == 0) S.bad |= 1; reset_data(); }
    wsync();
  }
  real R[9], fsm[6], asm_[6], Dblk[21], qacc[6], fcon[6];
  bool pass_on = on;
  for (int pass = 0; pass < 2; pass++) {
  // ---------------- position stage
  if (pass_on) {
    normalize4(B.q);
    quat2mat(R, B.q);
    for (int k = 0; k < 3; k++) { S.xpos[3 * b + k] = B.x[k]; S.vw[6 * b + k] = B.v[k]; }
    for (int k = 0; k < 9; k++) S.xmat[9 * b + k] = R[k];
    real ww[3];
    mulMV(ww, R, B.v + 3);
    for (int k = 0; k < 3; k++) S.vw[6 * b + 3 + k] = ww[k];
    for (int i = 0; i < 2; i++) {   // end-cap centres s(2b), s(2b+1): geoms 1, 2 of the bar
      real c[3];
      mulMV(c, R, m.gpos[5 * b + 1 + i]);
      for (int k = 0; k < 3; k++) S.sph[3 * (2 * b + i) + k] = B.x[k] + c[k];
    }
    for (int n = 0; n < m.nends[b]; n++) {
      int end = m.ends[b][n];
      real c[3];
      mulMV(c, R, m.tsite[end]);
      for (int k = 0; k < 3; k++) S.u.site[3 * end + k] = B.x[k] + c[k];
    }
  }
  wsync();
  // ---------------- tendons: length, velocity, spring-damper / actuator force, smooth force, damping block
  if (pass_on) {
    real f[6] = {0, 0, 0, 0, 0, 0};
    for (int e = 0; e < 21; e++) Dblk[e] = 0;
    TB_UNROLL1
    for (int n = 0; n < m.nends[b]; n++) {
      const int end = m.ends[b][n], t = end >> 1, s = end & 1, oe = end ^ 1, ob = m.tbody[oe];
      real pown[3], poth[3], dir[3];
      copy3(pown, S.u.site + 3 * end); copy3(poth, S.u.site + 3 * oe);
      if (s) sub3(dir, pown, poth); else sub3(dir, poth, pown);
      real len = normalize3(dir);
      real ro[3], co[3], cw[3], Jw[3];
      sub3(ro, pown, B.x); cross3(co, ro, dir); mulMTV(Jw, R, co);
      real vown = dot3(dir, B.v) + dot3(Jw, B.v + 3);
      sub3(ro, poth, S.xpos + 3 * ob); cross3(cw, ro, dir);
      real voth = dot3(dir, S.vw + 6 * ob) + dot3(cw, S.vw + 6 * ob + 3);
      real vel = s ? (vown - voth) : (voth - vown);
      real frc = 0, Bt = -m.tdamp[t];
      if (m.tk[t] > 0) {
        if (len > m.tls[t][1]) frc = m.tk[t] * (m.tls[t][1] - len);
        else if (len < m.tls[t][0]) frc = m.tk[t] * (m.tls[t][0] - len);
      }
      frc -= m.tdamp[t] * vel;
      const int a = m.ten_act[t];
      if (a >= 0) {
        real ctrl = (real)S.ctrl[a], input;
        if (m.ctrllimited) ctrl = clampr(ctrl, m.ctrlrange[0], m.ctrlrange[1]);
        real act = (real)S.act[a];
        if (m.dyntype) { if (s == 0) S.actdot[a] = (ctrl - act) / tmax(MINV, m.dynprm0); input = act; }
        else { if (s == 0) S.actdot[a] = 0; input = ctrl; }
        real fa = m.gain * input + m.bias[0] + m.bias[1] * len + m.bias[2] * vel;
        bool clamped = false;
        if (m.forcelimited) {
          clamped = (fa <= m.forcerange[0] || fa >= m.forcerange[1]);
          fa = clampr(fa, m.forcerange[0], m.forcerange[1]);
        }
        frc += fa;
        if (m.bias[2] != 0 && ((m.flags & 1u) || !clamped)) Bt += m.bias[2];
      }
      if (s == 0) S.tlen[t] = len;
      real sg = s ? frc : -frc;
      real j[6] = {dir[0], dir[1], dir[2], Jw[0], Jw[1], Jw[2]};
      for (int k = 0; k < 6; k++) f[k] += sg * j[k];
      rank1_sym(Dblk, Bt, j);
    }
    // qfrc_smooth = passive + actuator - bias ; qacc_smooth = M^-1 qfrc_smooth
    const real* I = m.inertia[b];
    const real* w = B.v + 3;
    for (int k = 0; k < 3; k++) {
      fsm[k] = f[k] + m.M[6 * b + k] * m.grav[k];
      int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      fsm[3 + k] = f[3 + k] - (w[k1] * (I[k2] * w[k2]) - w[k2] * (I[k1] * w[k1]));
    }
    for (int k = 0; k < 6; k++) asm_[k] = fsm[k] * m.invM[6 * b + k];
  }
  wsync();   // the tendon end points are dead: the union now carries the bar-bar hand-off
  // ---------------- collision
  if (pass_on) ncon = 0;
  int overflow = 0, nmpr = 0, nh = 0;
  HandCon<real> found[MAXH];   // bar-bar contacts of this lane's pair
  if (pass_on) {
    if (m.floor_type == 0) {
      TB_UNROLL1
      for (int g = 0; g < 5; g++) {
        const int Gi = 5 * b + g;
        real c[3], tmp[3];
        mulMV(c, R, m.gpos[Gi]); add3(c, c, B.x);
        sub3(tmp, c, m.fpos);
        real cdist = dot3(tmp, m.fnormal);
        if (m.gtype[Gi] == GEOM_SPHERE) {
          real r = m.gsize[Gi][0];
          if (cdist <= r) {
            real dist = cdist - r, pos[3];
            copy3(pos, c); addscl3(pos, m.fnormal, -dist / 2 - r);
            add_contact(con, ncon, overflow, -1, b, dist, pos, m.fnormal, S.xpos);
          }
        } else if (cdist <= m.gbound[Gi]) {
          real axis[3] = {R[2], R[5], R[8]}, xaxis[3] = {R[0], R[3], R[6]}, dist[4], pts[4][3];
          int cnt = 0;
          plane_cylinder_points(m, c, axis, m.gsize[Gi][0], m.gsize[Gi][1], xaxis, cnt, dist, pts);
          for (int k = 0; k < cnt; k++) add_contact(con, ncon, overflow, -1, b, dist[k], pts[k], m.fnormal, S.xpos);
        }
      }
    } else {
      // height field (frame axis-aligned at fpos): per geom, the prisms under its AABB whose top reaches the AABB's
      // bottom (MuJoCo's test) and that the geom can reach (conservative cull), each through MPR
      TB_UNROLL1
      for (int g = 0; g < 5; g++) {
        const int Gi = 5 * b + g;
        real gc[3], pos[3];
        mulMV(gc, R, m.gpos[Gi]); add3(gc, gc, B.x);
        sub3(pos, gc, m.fpos);
        const real r = m.gsize[Gi][0], hl = m.gsize[Gi][1], rb = m.gbound[Gi];
        bool ok = true;
        for (int i = 0; i < 2; i++) if (m.hsize[i] < pos[i] - rb || -m.hsize[i] > pos[i] + rb) ok = false;
        if (m.hsize[2] < pos[2] - rb || -m.hsize[3] > pos[2] + rb) ok = false;
        real ext[3];
        if (m.gtype[Gi] == GEOM_SPHERE) ext[0] = ext[1] = ext[2] = r;
        else for (int i = 0; i < 3; i++) {   // AABB half extents of a cylinder = what the +-axis support queries return
          real az = R[3 * i + 2];
          ext[i] = r * tsqrt(tmax(real(0), 1 - az * az)) + hl * tabs(az);
        }
        real xmin = pos[0] - ext[0], xmax = pos[0] + ext[0], ymin = pos[1] - ext[1], ymax = pos[1] + ext[1];
        real zmin = pos[2] - ext[2], zmax = pos[2] + ext[2];
        if (xmin > m.hsize[0] || xmax < -m.hsize[0] || ymin > m.hsize[1] || ymax < -m.hsize[1] || zmin > m.hsize[2] || zmax < -m.hsize[3]) ok = false;
        if (!ok) continue;
        int cmin = (int)tfloor((xmin + m.hsize[0]) / (2 * m.hsize[0]) * (m.ncol - 1));
        int cmax = (int)tceil((xmax + m.hsize[0]) / (2 * m.hsize[0]) * (m.ncol - 1));
        int rmin = (int)tfloor((ymin + m.hsize[1]) / (2 * m.hsize[1]) * (m.nrow - 1));
        int rmax = (int)tceil((ymax + m.hsize[1]) / (2 * m.hsize[1]) * (m.nrow - 1));
        if (cmin < 0) cmin = 0;
        if (cmax > m.ncol - 1) cmax = m.ncol - 1;
        if (rmin < 0) rmin = 0;
        if (rmax > m.nrow - 1) rmax = m.nrow - 1;
        const int per_row = 2 * (cmax - cmin + 1) - 2;
        if (per_row <= 0 || rmax <= rmin) continue;
        CObj<real> o2;
        o2.type = m.gtype[Gi]; copy3(o2.pos, pos); o2.size[0] = r; o2.size[1] = hl;
        for (int k = 0; k < 9; k++) o2.R[k] = R[k];
        TB_UNROLL1
        for (int rr = rmin; rr < rmax; rr++) {
          TB_UNROLL1
          for (int k = 0; k < per_row; k++) {
            CObj<real> o1;
            o1.type = 100;
            hf_prism(m, rr, cmin, k, o1);
            if (!(o1.pz[0] >= zmin || o1.pz[1] >= zmin || o1.pz[2] >= zmin)) continue;
            if (hf_above_top_plane(o1, o2.type, pos, R, r, hl)) continue;
            real depth = 0, dir[3] = {0, 0, 1}, cp[3] = {0, 0, 0};
            nmpr++;
            bool hit = mpr_penetration(o1, o2, m.mpr_tol, m.mpr_iterations, &depth, dir, cp);
            if (hit && ccd_vec_is_origin(dir)) hit = false;
            if (hit) {
              add3(cp, cp, m.fpos);
              if ((m.flags & 4u) && o2.type == GEOM_SPHERE) {
                real nn[3]; sub3(nn, gc, cp);
                if (tsqrt(dot3(nn, nn)) > MINV) { normalize3(nn); copy3(dir, nn); }
              }
              add_contact(con, ncon, overflow, -1, b, -depth, cp, dir, S.xpos);
            }
          }
        }
      }
    }
    // bar-bar: this lane takes pair p = its bar index: (0,1), (0,2), (1,2).  25 geom pairs: bounding spheres, analytic
    // capsule bound (conservative: MPR reports penetration only for intersecting shapes), then sphere-sphere / MPR
    {
      const int p = b, pb1 = p == 2 ? 1 : 0, pb2 = p == 0 ? 1 : 2;
      const real *X1 = S.xpos + 3 * pb1, *X2 = S.xpos + 3 * pb2, *R1 = S.xmat + 9 * pb1, *R2 = S.xmat + 9 * pb2;
      // bar-level cull: the bars' bounding capsules (axis segment of the whole bar, largest radius)
      TB_UNROLL1
      for (int i = 0; i < 25; i++) {
        int g1 = 5 * pb1 + i / 5, g2 = 5 * pb2 + i % 5;
        int cb1 = pb1, cb2 = pb2;
        const real *Ra = R1, *Rb = R2;
        real c1[3], c2[3];
        mulMV(c1, R1, m.gpos[g1]); add3(c1, c1, X1);
        mulMV(c2, R2, m.gpos[g2]); add3(c2, c2, X2);
        real dd[3]; sub3(dd, c1, c2);
        real bs = m.gbound[g1] + m.gbound[g2];
        real d2c = dot3(dd, dd);
        if (d2c > bs * bs) continue;   // MuJoCo's own bounding-sphere test
        int t1 = m.gtype[g1], t2 = m.gtype[g2];
        real rs = m.gsize[g1][0] + m.gsize[g2][0] + real(1e-6);
        real a1[3] = {R1[2] * m.gsize[g1][1], R1[5] * m.gsize[g1][1], R1[8] * m.gsize[g1][1]};
        real a2[3] = {R2[2] * m.gsize[g2][1], R2[5] * m.gsize[g2][1], R2[8] * m.gsize[g2][1]};
        real d2;
        if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) d2 = d2c;
        else if (t1 == GEOM_SPHERE) d2 = ptseg_dist2(c1, c2, a2);
        else if (t2 == GEOM_SPHERE) d2 = ptseg_dist2(c2, c1, a1);
        else d2 = segseg_dist2(c1, a1, c2, a2);
        if (!(d2 <= rs * rs)) continue;
        if (t1 > t2) {   // lower geom type first
          int ti = g1; g1 = g2; g2 = ti; ti = cb1; cb1 = cb2; cb2 = ti;
          for (int k = 0; k < 3; k++) { real tv = c1[k]; c1[k] = c2[k]; c2[k] = tv; }
          const real* tp = Ra; Ra = Rb; Rb = tp;
        }
        bool hit;
        real dist = 1, pos[3] = {0, 0, 0}, nrm[3] = {1, 0, 0};
        nmpr++;
        if (m.gtype[g2] == GEOM_SPHERE) {
          sub3(nrm, c2, c1);
          real len = normalize3(nrm), r1 = m.gsize[g1][0];
          dist = len - r1 - m.gsize[g2][0];
          hit = dist <= 0;
          copy3(pos, c1); addscl3(pos, nrm, r1 + dist / 2);
        } else {
          CObj<real> o1, o2;
          o1.type = m.gtype[g1]; copy3(o1.pos, c1); o1.size[0] = m.gsize[g1][0]; o1.size[1] = m.gsize[g1][1];
          o2.type = m.gtype[g2]; copy3(o2.pos, c2); o2.size[0] = m.gsize[g2][0]; o2.size[1] = m.gsize[g2][1];
          for (int k = 0; k < 9; k++) { o1.R[k] = Ra[k]; o2.R[k] = Rb[k]; }
          real depth;
          hit = mpr_penetration(o1, o2, m.mpr_tol, m.mpr_iterations, &depth, nrm, pos);
          if (hit && ccd_vec_is_origin(nrm)) hit = false;
          dist = -depth;
          if (hit && (m.flags & 4u) && o1.type == GEOM_SPHERE) {
            real nn[3]; sub3(nn, pos, c1);
            if (tsqrt(dot3(nn, nn)) > MINV) { normalize3(nn); copy3(nrm, nn); }
          }
        }
        if (hit && dist < 0) {
          if (nh < MAXH) {
            HandCon<real>& hc = found[nh++];
            hc.dist = dist; copy3(hc.pos, pos); copy3(hc.nrm, nrm); hc.b1 = cb1; hc.b2 = cb2;
          } else overflow = 1;
        }
      }
    }
  }
  // hand-off to lane 2, KHAND contacts per pair and round (one round unless a pair has more than KHAND contacts)
  bool coupled = false;   // does this env have a bar-bar contact (lane 2 owns them)
  TB_UNROLL1
  for (int round = 0;; round++) {
    const int lo = round * KHAND;
    if (pass_on) {
      int cnt = nh - lo; cnt = cnt < 0 ? 0 : (cnt > KHAND ? KHAND : cnt);
      for (int k = 0; k < cnt; k++) S.u.hand[b][k] = found[lo + k];
      S.nhand[b] = cnt;
    }
    wsync();
    if (pass_on && b == 2) {
      for (int p = 0; p < 3; p++)
        for (int k = 0; k < S.nhand[p]; k++) {
          const HandCon<real>& hc = S.u.hand[p][k];
          add_contact(con, ncon, overflow, hc.b1, hc.b2, hc.dist, hc.pos, hc.nrm, S.xpos);
          coupled = true;
        }
    }
    if (!any(pass_on && nh > lo + KHAND)) break;
    wsync();
  }
  coupled = grp_any(coupled, base);
  {
    int nact = isum3(pass_on ? ncon : 0, base), ov = isum3(pass_on ? overflow : 0, base), nm = isum3(pass_on ? nmpr : 0, base);
    if (pass_on && b == 0) { S.nact = nact; if (ov) S.overflow = 1; S.nmpr += nm; }
  }
  // rows: velocity, impedance, reference acceleration
  if (pass_on) {
    TB_UNROLL1
    for (int n = 0; n < ncon; n++) {
      Con<real>& c = con[n];
      real dist = c.aref[0], vel[6];
      con_mulJ(c, S.vw, vel);
      real imp = impedance(m, dist);
      real tran = (c.b1 >= 0 ? m.invw_tran[c.b1] : real(0)) + m.invw_tran[c.b2];
      c.D0 = trcp(tmax(MINV, tdiv(1 - imp, imp) * tran));
      for (int r = 0; r < 6; r++) c.aref[r] = -m.B * vel[r] - (r ? real(0) : m.K * imp * dist);
    }
  }
  wsync();   // the hand-off slots are dead: the union now carries the solver exchange
  // ---------------- mj_fwdConstraint: warm-start choice + Newton
  const int nact_env = isum3(pass_on ? ncon : 0, base);
  bool act = pass_on && nact_env > 0;
  const real* Mb = m.M + 6 * b;
  // publishes a per-bar 6-vector (lin world, ang body-local) as a world-frame twist
  auto publish = [&](const real* a, bool doit) {
    if (doit) {
      real w[3];
      mulMV(w, R, a + 3);
      for (int k = 0; k < 3; k++) { S.u.sol.xv[6 * b + k] = a[k]; S.u.sol.xv[6 * b + 3 + k] = w[k]; }
    }
  };
  auto jar_from_xv = [&]() {
    TB_UNROLL1
    for (int n = 0; n < ncon; n++) {
      real o[6];
      con_mulJ(con[n], S.u.sol.xv, o);
      for (int r = 0; r < 6; r++) con[n].jar[r] = o[r] - con[n].aref[r];
    }
  };
  auto cost_only = [&](const real* a) -> real {   // this lane's share of the cost at a (jar current)
    real s = 0;
    TB_UNROLL1
    for (int n = 0; n < ncon; n++) s += con_update(con[n], m, false);
    for (int k = 0; k < 6; k++) { real d = a[k] - asm_[k]; s += real(0.5) * Mb[k] * d * d; }
    return s;
  };
  bool use_smooth = false;
  if (any(act)) {
    publish(asm_, act); wsync();
    real csm = 0, cws = 0;
    if (act) { jar_from_xv(); csm = cost_only(asm_); }
    wsync();
    publish(B.warm, act); wsync();
    if (act) { jar_from_xv(); cws = cost_only(B.warm); }
    csm = sum3(csm, base); cws = sum3(cws, base);
    use_smooth = cws > csm;
    wsync();
    if (any(act && use_smooth)) {
      publish(asm_, act && use_smooth); wsync();
      if (act && use_smooth) jar_from_xv();
      wsync();
    }
  }
  for (int k = 0; k < 6; k++) { qacc[k] = (act && !use_smooth) ? B.warm[k] : asm_[k]; fcon[k] = 0; }
  real grad[6], search[6] = {0, 0, 0, 0, 0, 0};
  real cost = 0, oldcost = 0;
  int iter = 0, nls = 0;
  bool first = true;
  TB_UNROLL1
  for (;;) {
    if (!any(act)) break;
    // ---- cost, forces, gradient at qacc
    real cpart = 0, Fw[3] = {0, 0, 0}, Tw[3] = {0, 0, 0};
    if (act && coupled) for (int k = 0; k < 6; k++) S.u.sol.fx[6 * b + k] = 0;
    if (any(act && coupled)) wsync();
    if (act) {
      TB_UNROLL1
      for (int n = 0; n < ncon; n++) {
        Con<real>& c = con[n];
        cpart += con_update(c, m, true);
        if (c.zone == ZONE_TOP) continue;
        real F[3], T[3], t[3];
        con_wrench(c, F, T);
        cross3(t, c.r2, F);
        if (c.b1 < 0) { for (int k = 0; k < 3; k++) { Fw[k] += F[k]; Tw[k] += t[k] + T[k]; } }
        else {   // bar-bar (this is lane 2): through shared memory to the two bars' lanes
          real* f2 = S.u.sol.fx + 6 * c.b2; real* f1 = S.u.sol.fx + 6 * c.b1;
          for (int k = 0; k < 3; k++) { f2[k] += F[k]; f2[3 + k] += t[k] + T[k]; }
          cross3(t, c.r1, F);
          for (int k = 0; k < 3; k++) { f1[k] -= F[k]; f1[3 + k] -= t[k] + T[k]; }
        }
      }
      for (int k = 0; k < 6; k++) { real d = qacc[k] - asm_[k]; cpart += real(0.5) * Mb[k] * d * d; }
    }
    if (any(act && coupled)) wsync();
    real gn = 0;
    if (act) {
      if (coupled) for (int k = 0; k < 3; k++) { Fw[k] += S.u.sol.fx[6 * b + k]; Tw[k] += S.u.sol.fx[6 * b + 3 + k]; }
      real tl[3];
      mulMTV(tl, R, Tw);
      for (int k = 0; k < 3; k++) { fcon[k] = Fw[k]; fcon[3 + k] = tl[k]; }
      for (int k = 0; k < 6; k++) { grad[k] = Mb[k] * (qacc[k] - asm_[k]) - fcon[k]; gn += grad[k] * grad[k]; }
    }
    real newcost = sum3(cpart, base);
    gn = sum3(gn, base);
    if (act) {
      oldcost = cost; cost = newcost;
      if (!first && (m.solscale * (oldcost - cost) < m.tol || gn < m.gradtol * m.gradtol)) act = false;   // converged
    }
    first = false;
    const bool go = act && iter < m.iterations;
    if (!go) act = false;
    if (!any(go)) break;
    // ---- Hessian blocks
    real H[21];
    if (go) {
      for (int i = 0, e = 0; i < 6; i++) for (int k = 0; k <= i; k++, e++) H[e] = (i == k) ? Mb[i] : real(0);
      TB_UNROLL1
      for (int n = 0; n < ncon; n++) {
        const Con<real>& c = con[n];
        if (c.zone == ZONE_TOP || c.b1 >= 0) continue;
        SideJ<real> J;
        side_rows(c, real(1), c.r2, R, J);
        side_hessian(c, m, J, H);
      }
    }
    if (any(go && coupled)) {
      // coupled env: lanes 0, 1 hand their diagonal block and gradient to lane 2, which adds the bar-bar contacts'
      // blocks, factorises the 18x18 system (structural zeros skipped) and returns the search direction
      if (go && coupled) {
        if (b < 2) for (int e = 0; e < 21; e++) S.u.sol.Hg[21 * b + e] = H[e];
        for (int k = 0; k < 6; k++) S.u.sol.fx[6 * b + k] = grad[k];
      }
      wsync();
      if (go && coupled && b == 2) {
        // blocks: D[bar] (packed, 21 each), then O[pair] (full 6x6: rows = the higher bar, cols = the lower one) for the
        // pairs (1,0), (2,0), (2,1)
        real Hb[3 * 21 + 3 * 36];
        real* const Dg = Hb; real* const Og = Hb + 63;
        bool cpl[3] = {false, false, false};
        for (int e = 0; e < 42; e++) Dg[e] = S.u.sol.Hg[e];
        for (int e = 0; e < 21; e++) Dg[42 + e] = H[e];
        TB_UNROLL1
        for (int n = 0; n < ncon; n++) {
          const Con<real>& c = con[n];
          if (c.zone == ZONE_TOP || c.b1 < 0) continue;
          SideJ<real> J1, J2;
          side_rows(c, real(-1), c.r1, S.xmat + 9 * c.b1, J1);
          side_rows(c, real(1), c.r2, S.xmat + 9 * c.b2, J2);
          const int hi = c.b1 > c.b2 ? c.b1 : c.b2, lo = c.b1 > c.b2 ? c.b2 : c.b1, p = hi == 1 ? 0 : (lo == 0 ? 1 : 2);
          real* X = Og + 36 * p;
          if (!cpl[p]) { for (int e = 0; e < 36; e++) X[e] = 0; cpl[p] = true; }
          side_hessian(c, m, J1, Dg + 21 * c.b1);
          side_hessian(c, m, J2, Dg + 21 * c.b2);
          const SideJ<real>& Jh = c.b2 > c.b1 ? J2 : J1;   // rows: the higher bar
          const SideJ<real>& Jl = c.b2 > c.b1 ? J1 : J2;
          const real* wt = m.wtab[c.zone == ZONE_MIDDLE ? 1 : 0];
          for (int r = 0; r < 6; r++) {
            real w = c.wcoef * wt[r];
            if (w != 0) { real jh[6], jl[6]; side_row6(Jh, r, jh); side_row6(Jl, r, jl); rank1_gen(X, w, jh, jl); }
          }
          if (c.zone == ZONE_MIDDLE) {
            real ah[6], bh[6], al[6], bl[6];
            side_cone(c, m, Jh, ah, bh); side_cone(c, m, Jl, al, bl);
            rank1_gen(X, c.ca, ah, al);
            rank1_gen(X, -c.cb, bh, bl);
          }
        }
        // block LDL^T in the order 0, 1, 2; structurally absent blocks are skipped (fill-in only in pair (2,1))
        real dinv[NV], x[NV];
        for (int i = 0; i < NV; i++) x[i] = S.u.sol.fx[i];
        real *D0 = Dg, *D1 = Dg + 21, *D2 = Dg + 42, *O10 = Og, *O20 = Og + 36, *O21 = Og + 72;
        blk_ldl(D0, dinv);
        if (cpl[0]) { blk_trsm(O10, D0, dinv); blk_syrk(D1, O10, D0); }
        if (cpl[1]) { blk_trsm(O20, D0, dinv); blk_syrk(D2, O20, D0); }
        if (cpl[0] && cpl[1]) {
          if (!cpl[2]) { for (int e = 0; e < 36; e++) O21[e] = 0; cpl[2] = true; }
          blk_gemm(O21, O20, D0, O10);
        }
        blk_ldl(D1, dinv + 6);
        if (cpl[2]) { blk_trsm(O21, D1, dinv + 6); blk_syrk(D2, O21, D1); }
        blk_ldl(D2, dinv + 12);
        blk_fwd(D0, x);
        if (cpl[0]) blk_gemv_sub(x + 6, O10, x);
        blk_fwd(D1, x + 6);
        if (cpl[1]) blk_gemv_sub(x + 12, O20, x);
        if (cpl[2]) blk_gemv_sub(x + 12, O21, x + 6);
        blk_fwd(D2, x + 12);
        for (int i = 0; i < NV; i++) x[i] *= dinv[i];
        blk_bwd(D2, x + 12);
        if (cpl[2]) blk_gemvT_sub(x + 6, O21, x + 12);
        blk_bwd(D1, x + 6);
        if (cpl[0]) blk_gemvT_sub(x, O10, x + 6);
        if (cpl[1]) blk_gemvT_sub(x, O20, x + 12);
        blk_bwd(D0, x);
        for (int i = 0; i < NV; i++) S.u.sol.xv[i] = -x[i];
      }
      wsync();
      if (go && coupled) for (int k = 0; k < 6; k++) search[k] = S.u.sol.xv[6 * b + k];
      wsync();
    }
    if (go && !coupled) {   // this bar's own 6x6 block: LDL^T and solve in registers
      real dinv[6], x[6];
      for (int k = 0; k < 6; k++) x[k] = grad[k];
      blk_ldl(H, dinv);
      blk_fwd(H, x);
      for (int k = 0; k < 6; k++) x[k] *= dinv[k];
      blk_bwd(H, x);
      for (int k = 0; k < 6; k++) search[k] = -x[k];
    }
    // ---- exact line search along search: mj_solPrimal's bracketing search as a per-env state machine.  Every tick
    // evaluates cost / slope / curvature at ONE step size per env (3-lane sums), then each env advances its own
    // bracketing logic, so the envs of a warp stay in lock step whatever their individual search sequences are.
    real snorm = 0, gs = 0, qG1 = 0, qG2 = 0, gauss = 0;
    if (go) for (int k = 0; k < 6; k++) {
      real sk = search[k], d = qacc[k] - asm_[k];
      snorm += sk * sk; gs += grad[k] * sk;
      qG1 += sk * (Mb[k] * qacc[k]) - fsm[k] * sk;
      qG2 += real(0.5) * sk * (Mb[k] * sk);
      gauss += real(0.5) * Mb[k] * d * d;
    }
    snorm = sum3(snorm, base); gs = sum3(gs, base);
    snorm = tsqrt(snorm);
    publish(search, go); wsync();
    if (go) {
      TB_UNROLL1
      for (int n = 0; n < ncon; n++) {
        Con<real>& k = con[n];
        con_mulJ(k, S.u.sol.xv, k.jv);
        real q0 = 0, q1 = 0, q2 = 0, UU = 0, UV = 0, VV = 0;
        for (int j = 0; j < 6; j++) {
          real D = k.D0 * m.dscale[j], ja = k.jar[j], jv = k.jv[j];
          q0 += real(0.5) * D * ja * ja; q1 += D * ja * jv; q2 += real(0.5) * D * jv * jv;
          if (j > 0) { real U = ja * m.fr[j - 1], V = jv * m.fr[j - 1]; UU += U * U; UV += U * V; VV += V * V; }
        }
        k.q0 = q0; k.q1 = q1; k.q2 = q2;
        k.U0 = k.jar[0] * m.mu; k.V0 = k.jv[0] * m.mu; k.UU = UU; k.UV = UV; k.VV = VV;
      }
    }
    wsync();
    real alpha = 0;
    int evals = 1;
    {
      struct Pnt { real alpha, cost, d0, d1; };
      // W_*: waiting for the evaluation it requested; L_*: pure logic, resolved without an evaluation
      enum { W_P1 = 0, W_A, W_P1NEXT, W_MID, W_B1, W_B2, L_ACHECK, L_AFTERA, L_BCHECK, L_DOB2, L_ENDITER, L_FINAL, LS_DONE };
      const real gtol = m.tol * m.ls_tol * snorm * (m.meaninertia * NV);
      const int maxe = m.ls_iterations;
      Pnt p0, p1, p2, pmid, p1next, p2next, c0, c1, c2;
      p0.alpha = 0; p0.cost = cost; p0.d0 = gs; p0.d1 = -gs > 0 ? -gs : MINV;   // alpha = 0 is analytic (H search = -grad)
      p1 = p2 = pmid = p1next = p2next = c0 = c1 = c2 = p0;
      int st = LS_DONE, dirn = 1;
      bool p2update = false, b1 = false, b2 = false;
      real aeval = 0;
      auto newton = [&](const Pnt& p) { return p.alpha - tdiv(p.d0, p.d1); };
      auto bracket = [&](Pnt& p) {   // update_bracket against the candidates captured at the mid-point evaluation
        int flag = 0;
        const Pnt* cand[3] = {&c0, &c1, &c2};
        for (int i = 0; i < 3; i++) {
          if (p.d0 < 0 && cand[i]->d0 < 0 && p.d0 < cand[i]->d0) { p = *cand[i]; flag = 1; }
          else if (p.d0 > 0 && cand[i]->d0 > 0 && p.d0 > cand[i]->d0) { p = *cand[i]; flag = 2; }
        }
        return flag;
      };
      if (go && !(snorm < MINV)) { st = W_P1; aeval = newton(p0); }
      TB_UNROLL1
      for (;;) {
        const bool ev = st != LS_DONE;
        if (!any(ev)) break;
        real c_ = 0, d0_ = 0, d1_ = 0;
        if (ev) {
          TB_UNROLL1
          for (int n = 0; n < ncon; n++) con_ls(con[n], m, aeval, c_, d0_, d1_);
          c_ += aeval * aeval * qG2 + aeval * qG1 + gauss; d0_ += 2 * aeval * qG2 + qG1; d1_ += 2 * qG2;
        }
        Pnt r;
        r.alpha = aeval; r.cost = sum3(c_, base); r.d0 = sum3(d0_, base); r.d1 = sum3(d1_, base);
        if (r.d1 <= 0) r.d1 = MINV;
        if (ev) {
          evals++;
          switch (st) {
            case W_P1:
              p1 = r;
              if (p0.cost < p1.cost) p1 = p0;
              if (tabs(p1.d0) < gtol) { alpha = p1.alpha; st = LS_DONE; }
              else { dirn = p1.d0 < 0 ? 1 : -1; p2 = p1; p2update = false; st = L_ACHECK; }
              break;
            case W_A:
              p1 = r;
              if (tabs(p1.d0) < gtol) { alpha = p1.alpha; st = LS_DONE; } else st = L_ACHECK;
              break;
            case W_P1NEXT: p1next = r; st = L_BCHECK; break;
            case W_MID: {
              pmid = r;
              c0 = p1next; c1 = p2next; c2 = pmid;
              const Pnt* cand[3] = {&c0, &c1, &c2};
              int best = -1; real bestcost = 0;
              for (int i = 0; i < 3; i++)
                if (tabs(cand[i]->d0) < gtol && (best == -1 || cand[i]->cost < bestcost)) { bestcost = cand[i]->cost; best = i; }
              if (best >= 0) { alpha = cand[best]->alpha; st = LS_DONE; }
              else {
                b1 = bracket(p1) != 0;
                if (b1) { aeval = newton(p1); st = W_B1; } else st = L_DOB2;
              }
              break;
            }
            case W_B1: p1next = r; st = L_DOB2; break;
            case W_B2: p2next = r; st = L_ENDITER; break;
            default: break;
          }
          TB_UNROLL1
          for (;;) {   // logic transitions until the env requests an evaluation or finishes
            if (st == L_ACHECK) {
              if (p1.d0 * dirn <= -gtol && evals < maxe) { p2 = p1; p2update = true; aeval = newton(p1); st = W_A; }
              else st = L_AFTERA;
            } else if (st == L_AFTERA) {
              if (evals >= maxe || !p2update) { alpha = p1.alpha; st = LS_DONE; }
              else { p2next = p1; aeval = newton(p1); st = W_P1NEXT; }
            } else if (st == L_BCHECK) {
              if (evals < maxe) { aeval = real(0.5) * (p1.alpha + p2.alpha); st = W_MID; } else st = L_FINAL;
            } else if (st == L_DOB2) {
              b2 = bracket(p2) != 0;
              if (b2) { aeval = newton(p2); st = W_B2; } else st = L_ENDITER;
            } else if (st == L_ENDITER) {
              if (!b1 && !b2) { alpha = pmid.alpha; st = LS_DONE; } else st = L_BCHECK;
            } else if (st == L_FINAL) {
              if (p1.cost <= p2.cost && p1.cost < p0.cost) alpha = p1.alpha;
              else if (p2.cost <= p1.cost && p2.cost < p0.cost) alpha = p2.alpha;
              else alpha = 0;
              st = LS_DONE;
            } else break;
          }
        }
      }
    }
    // ---- move
    if (go) {
      nls += evals;
      if (alpha == 0) act = false;
      else {
        for (int k = 0; k < 6; k++) qacc[k] += alpha * search[k];
        TB_UNROLL1
        for (int n = 0; n < ncon; n++) for (int r = 0; r < 6; r++) con[n].jar[r] += alpha * con[n].jv[r];
        iter++;
      }
    }
  }
  if (pass_on) {
    if (nact_env > 0) {
      for (int k = 0; k < 6; k++) B.warm[k] = qacc[k];
      if (b == 0) { S.niter += iter; S.nls += nls; }
    } else for (int k = 0; k < 6; k++) B.warm[k] = asm_[k];
  }
  // ---------------- mj_checkAcc: a bad acceleration resets the env and repeats the forward pass
  if (!integ || pass == 1) break;
  {
    bool bad = false;
    for (int k = 0; k < 6; k++) bad |= is_bad(qacc[k]);
    bad = grp_any(bad && on, base);
    if (!any(bad)) break;
    pass_on = on && bad;
    if (pass_on) { if (b == 0) S.bad |= 4; reset_data(); }
    wsync();
  }
  }  // pass
  // ---------------- implicitfast + advance
  if (integ && on) {
    const real h = m.h;
    // (M - h D) x = qfrc_smooth + qfrc_constraint on the bar's 6x6 block (Cholesky, packed lower)
    real A[21], x[6];
    for (int e = 0; e < 21; e++) A[e] = -h * Dblk[e];
    for (int r = 0; r < 6; r++) { A[r * (r + 1) / 2 + r] += m.M[6 * b + r]; x[r] = fsm[r] + fcon[r]; }
    TB_UNROLL
    for (int j = 0; j < 6; j++) {
      real s = A[j * (j + 1) / 2 + j];
      TB_UNROLL
      for (int k = 0; k < j; k++) s -= A[j * (j + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
      real inv = trcp(tsqrt(s));
      A[j * (j + 1) / 2 + j] = inv;   // 1 / L_jj
      TB_UNROLL
      for (int i = j + 1; i < 6; i++) {
        real t = A[i * (i + 1) / 2 + j];
        TB_UNROLL
        for (int k = 0; k < j; k++) t -= A[i * (i + 1) / 2 + k] * A[j * (j + 1) / 2 + k];
        A[i * (i + 1) / 2 + j] = t * inv;
      }
    }
    TB_UNROLL
    for (int i = 0; i < 6; i++) {
      real t = x[i];
      TB_UNROLL
      for (int k = 0; k < i; k++) t -= A[i * (i + 1) / 2 + k] * x[k];
      x[i] = t * A[i * (i + 1) / 2 + i];
    }
    TB_UNROLL
    for (int i = 5; i >= 0; i--) {
      real t = x[i];
      TB_UNROLL
      for (int k = i + 1; k < 6; k++) t -= A[k * (k + 1) / 2 + i] * x[k];
      x[i] = t * A[i * (i + 1) / 2 + i];
    }
    for (int k = 0; k < 6; k++) B.v[k] += h * x[k];
    for (int k = 0; k < 3; k++) B.x[k] += h * B.v[k];
    real ax[3] = {B.v[3], B.v[4], B.v[5]}, qr[4], qn[4];
    real ang = h * normalize3(ax);
    if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
    else { real sn, cs; tsincos(ang * real(0.5), &sn, &cs); qr[0] = cs; qr[1] = ax[0] * sn; qr[2] = ax[1] * sn; qr[3] = ax[2] * sn; }
    normalize4(B.q);
    const real* a = B.q;
    qn[0] = a[0] * qr[0] - a[1] * qr[1] - a[2] * qr[2] - a[3] * qr[3];
    qn[1] = a[0] * qr[1] + a[1] * qr[0] + a[2] * qr[3] - a[3] * qr[2];
    qn[2] = a[0] * qr[2] - a[1] * qr[3] + a[2] * qr[0] + a[3] * qr[1];
    qn[3] = a[0] * qr[3] + a[1] * qr[2] - a[2] * qr[1] + a[3] * qr[0];
    for (int k = 0; k < 4; k++) B.q[k] = qn[k];
    if (m.dyntype && b == 0) for (int i = 0; i < NACT; i++) S.act[i] += (double)(h * S.actdot[i]);
  }
  wsync();
}

// mj_rnePostConstraint: cfrc_ext rows [torque; force] for world + 3 bars about the (stale) body positions, from the
// contacts of the last pass; also the total bar-bar contact force magnitude (run.py:155-161).  Leaves S.u.cfrc.
template <typename real>
TB_FN void cfrc_stage(EnvSh<real>& S, const ModelT<real>& m, const LaneCtx& L, bool on, const Con<real>* con, int ncon) {
  const int b = L.bar, base = L.base;
  // world row is taken about the mass-weighted centre of the three bars
  real com[3] = {0, 0, 0}, mt = 0;
  for (int bb = 0; bb < NBAR; bb++) { addscl3(com, S.xpos + 3 * bb, m.M[6 * bb]); mt += m.M[6 * bb]; }
  scl3(com, com, 1 / mt);
  real own[6] = {0, 0, 0, 0, 0, 0}, world[6] = {0, 0, 0, 0, 0, 0}, barf = 0;
  wsync();   // everyone has read xpos-independent union data; the union becomes cfrc
  if (on) for (int k = 0; k < 24; k++) if (k / 6 == b + 1) S.u.cfrc[k] = 0;
  wsync();
  if (on) {
    for (int n = 0; n < ncon; n++) {
      const Con<real>& c = con[n];
      real F[3], T[3], pos[3], r[3], tq[3];
      con_wrench(c, F, T);
      add3(pos, S.xpos + 3 * c.b2, c.r2);
      if (c.b1 < 0) {
        cross3(tq, c.r2, F);
        for (int k = 0; k < 3; k++) { own[k] += tq[k] + T[k]; own[3 + k] += F[k]; }
        sub3(r, pos, com); cross3(tq, r, F);
        for (int k = 0; k < 3; k++) { world[k] -= tq[k] + T[k]; world[3 + k] -= F[k]; }
      } else {
        barf += tsqrt(c.force[0] * c.force[0] + c.force[1] * c.force[1] + c.force[2] * c.force[2]);
      }
    }
  }
  for (int k = 0; k < 6; k++) world[k] = sum3(world[k], base);
  if (on) {
    for (int k = 0; k < 6; k++) S.u.cfrc[6 * (b + 1) + k] = own[k];
    if (b == 0) for (int k = 0; k < 6; k++) S.u.cfrc[k] = world[k];
    if (b == 2) S.barforce = barf;
  }
  wsync();
  if (on && b == 2) {   // bar-bar contacts: lane 2 adds them to the two bars' rows
    for (int n = 0; n < ncon; n++) {
      const Con<real>& c = con[n];
      if (c.b1 < 0) continue;
      real F[3], T[3], tq[3];
      con_wrench(c, F, T);
      cross3(tq, c.r2, F);
      for (int k = 0; k < 3; k++) { S.u.cfrc[6 * (c.b2 + 1) + k] += tq[k] + T[k]; S.u.cfrc[6 * (c.b2 + 1) + 3 + k] += F[k]; }
      cross3(tq, c.r1, F);
      for (int k = 0; k < 3; k++) { S.u.cfrc[6 * (c.b1 + 1) + k] -= tq[k] + T[k]; S.u.cfrc[6 * (c.b1 + 1) + 3 + k] -= F[k]; }
    }
  }
  wsync();
}

}  // namespace tb
