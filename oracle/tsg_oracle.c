/* tsg_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, not product code).  See tsg_oracle.h.
 *
 * Readable dense fp64 restatement of MuJoCo 2.3.7's mj_step for the 3-bar
 * tensegrity.  PARITY UNPINNED at the level of single-step
 * dynamics (no MuJoCo binary, no stored trajectory in the reference); kinematics,
 * tendon lengths and the observation layout are pinned to 1e-12 against the
 * `_last_obs` vectors stored in the reference's checkpoints
 * (tests/test_golden_last_obs.py).  Section comments name the MuJoCo stage restated.
 */
#include "tsg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MINVAL 1e-15 /* mjMINVAL */
#define MAXVAL 1e10  /* mjMAXVAL */
#define MINIMP 0.0001
#define MAXIMP 0.9999
#define NV TSG_NV

enum { ST_QUADRATIC = 0, ST_SATISFIED = 1, ST_CONE = 4 };

int tsgo_sizeof_data(void) { return (int)sizeof(TsgoData); }
int tsgo_sizeof_model(void) { return (int)sizeof(TsgModel); }

/* ------------------------------------------------------------------ math */
static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(double *r, const double *a, const double *b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static void sub3(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static void add3(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static void copy3(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static void scl3(double *r, const double *a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static void addscl3(double *r, const double *a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
static double norm3(const double *a) { return sqrt(dot3(a, a)); }
/* mju_normalize3 */
static double normalize3(double *a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { double s = 1 / n; a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
/* mju_normalize4 */
static void normalize4(double *q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > MINVAL) { double s = 1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
/* mju_rotVecMat: r = R v ; mju_rotVecMatT: r = R^T v  (row-major 3x3) */
static void mulMV(double *r, const double *R, const double *v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulMTV(double *r, const double *R, const double *v) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
/* mju_quat2Mat */
static void quat2mat(double *R, const double *q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2 * (q12 - q03); R[2] = 2 * (q13 + q02);
  R[3] = 2 * (q12 + q03); R[5] = 2 * (q23 - q01);
  R[6] = 2 * (q13 - q02); R[7] = 2 * (q23 + q01);
}
/* mju_mulQuat */
static void mulQuat(double *r, const double *a, const double *b) {
  double t0 = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double t1 = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double t2 = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double t3 = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = t0; r[1] = t1; r[2] = t2; r[3] = t3;
}
static int is_bad(double x) { return isnan(x) || x > MAXVAL || x < -MAXVAL; }

/* ------------------------------------------------------------------ reset */
void tsgo_reset_data(const TsgModel *m, TsgoData *d) { /* mj_resetData */
  memset(d, 0, sizeof(*d));
  memcpy(d->qpos, m->qpos0, sizeof(d->qpos));
}

/* ------------------------------------------------------------------ kinematics (mj_kinematics, mj_comPos) */
void tsgo_kinematics(const TsgModel *m, TsgoData *d) {
  double mtot = 0;
  d->com_world[0] = d->com_world[1] = d->com_world[2] = 0;
  for (int b = 0; b < TSG_NBAR; b++) {
    double *q = d->qpos + 7 * b;
    normalize4(q + 3); /* in place, as mj_kinematics does for free joints */
    copy3(d->xpos[b], q);
    memcpy(d->xquat[b], q + 3, 4 * sizeof(double));
    quat2mat(d->xmat[b], q + 3);
    for (int g = 0; g < TSG_NGEOM_BAR; g++) {
      double v[3], gq[4];
      mulMV(v, d->xmat[b], m->geom_pos[b][g]);
      add3(d->geom_xpos[b][g], d->xpos[b], v);
      mulQuat(gq, d->xquat[b], m->geom_quat[b][g]);
      quat2mat(d->geom_xmat[b][g], gq);
    }
    addscl3(d->com_world, d->xpos[b], m->body_mass[b]);
    mtot += m->body_mass[b];
  }
  scl3(d->com_world, d->com_world, 1 / mtot);
  for (int t = 0; t < TSG_NTEN; t++)
    for (int e = 0; e < 2; e++) {
      int b = m->ten_body[t][e];
      double v[3];
      mulMV(v, d->xmat[b], m->ten_site[t][e]);
      add3(d->site_xpos[t][e], d->xpos[b], v);
    }
}

/* dof-space Jacobian row of direction `a` for a point at offset r (world) from
 * the origin of free body b (mj_jac for a free joint: linear dofs are world
 * axes, angular dofs are body-local axes): row = [a, R^T (r x a)] */
static void jac_point_row(double *row6, const double *xmat, const double *r, const double *a) {
  double t[3];
  copy3(row6, a);
  cross3(t, r, a);
  mulMTV(row6 + 3, xmat, t);
}

/* ------------------------------------------------------------------ tendons (mj_tendon) */
void tsgo_tendon(const TsgModel *m, TsgoData *d) {
  for (int t = 0; t < TSG_NTEN; t++) {
    double dif[3];
    sub3(dif, d->site_xpos[t][1], d->site_xpos[t][0]);
    d->ten_length[t] = normalize3(dif);
    memset(d->ten_J[t], 0, sizeof(d->ten_J[t]));
    for (int e = 0; e < 2; e++) {
      int b = m->ten_body[t][e];
      double r[3], row[6], s = e ? 1.0 : -1.0;
      sub3(r, d->site_xpos[t][e], d->xpos[b]);
      jac_point_row(row, d->xmat[b], r, dif);
      for (int k = 0; k < 6; k++) d->ten_J[t][6 * b + k] += s * row[k];
    }
  }
}

/* ------------------------------------------------------------------ MPR (libccd ccdMPRPenetration as bundled by MuJoCo) */
#define CCD_EPS 2.220446049250313e-16 /* DBL_EPSILON */
#define OBJ_PRISM 100

typedef struct { int type; const double *pos, *mat, *size; } CObj;
typedef struct { double v[3], v1[3], v2[3]; } Supp;

static int ccd_is_zero(double x) { return fabs(x) < CCD_EPS; }
static int ccd_eq(double a_, double b_) {
  double ab = fabs(a_ - b_);
  if (ab < CCD_EPS) return 1;
  double a = fabs(a_), b = fabs(b_);
  return (b > a) ? (ab < CCD_EPS * b) : (ab < CCD_EPS * a);
}
static int ccd_vec_eq(const double *a, const double *b) { return ccd_eq(a[0], b[0]) && ccd_eq(a[1], b[1]) && ccd_eq(a[2], b[2]); }
static void ccd_normalize(double *v) { double s = 1.0 / sqrt(dot3(v, v)); v[0] *= s; v[1] *= s; v[2] *= s; }

/* mjccd_center */
static void obj_center(const CObj *o, double *c) {
  if (o->type == OBJ_PRISM) {
    c[0] = c[1] = c[2] = 0;
    for (int i = 0; i < 6; i++) { c[0] += o->size[3 * i]; c[1] += o->size[3 * i + 1]; c[2] += o->size[3 * i + 2]; }
    c[0] /= 6; c[1] /= 6; c[2] /= 6;
  } else copy3(c, o->pos);
}
/* mjccd_support (margin 0) */
static void obj_support(const CObj *o, const double *dir, double *out) {
  if (o->type == OBJ_PRISM) { /* vertex maximising dot(dir) in the hfield frame */
    int best = 0; double bd = dot3(o->size, dir);
    for (int i = 1; i < 6; i++) { double dd = dot3(o->size + 3 * i, dir); if (dd > bd) { bd = dd; best = i; } }
    copy3(out, o->size + 3 * best);
    return;
  }
  double ld[3], res[3];
  mulMTV(ld, o->mat, dir);
  if (o->type == TSG_GEOM_SPHERE) scl3(res, ld, o->size[0]);
  else { /* cylinder */
    double tmp = sqrt(ld[0] * ld[0] + ld[1] * ld[1]);
    if (tmp > MINVAL) { res[0] = ld[0] / tmp * o->size[0]; res[1] = ld[1] / tmp * o->size[0]; }
    else res[0] = res[1] = 0;
    res[2] = (ld[2] > 0 ? 1.0 : (ld[2] < 0 ? -1.0 : 0.0)) * o->size[1];
  }
  mulMV(out, o->mat, res);
  add3(out, out, o->pos);
}
static void mink_support(const CObj *o1, const CObj *o2, const double *dir, Supp *s) {
  double nd[3] = {-dir[0], -dir[1], -dir[2]};
  obj_support(o1, dir, s->v1);
  obj_support(o2, nd, s->v2);
  sub3(s->v, s->v1, s->v2);
}
static void portal_dir(const Supp *p, double *dir) {
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v); sub3(b, p[3].v, p[1].v);
  cross3(dir, a, b); ccd_normalize(dir);
}
static int portal_reach_tol(const Supp *p, const Supp *v4, const double *dir, double tol) {
  double dv1 = dot3(p[1].v, dir), dv2 = dot3(p[2].v, dir), dv3 = dot3(p[3].v, dir), dv4 = dot3(v4->v, dir);
  double d1 = dv4 - dv1, d2 = dv4 - dv2, d3 = dv4 - dv3;
  d1 = fmin(d1, d2); d1 = fmin(d1, d3);
  return ccd_eq(d1, tol) || d1 < tol;
}
static void expand_portal(Supp *p, const Supp *v4) {
  double v4v0[3];
  cross3(v4v0, v4->v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) { if (dot3(p[2].v, v4v0) > 0) p[1] = *v4; else p[3] = *v4; }
  else { if (dot3(p[3].v, v4v0) > 0) p[2] = *v4; else p[1] = *v4; }
}
static double point_seg_dist2(const double *P, const double *x0, const double *b, double *wit) {
  double d[3], a[3], t, w[3], e[3];
  sub3(d, b, x0); sub3(a, x0, P);
  t = -1.0 * dot3(a, d); t /= dot3(d, d);
  if (t < 0 || ccd_is_zero(t)) copy3(w, x0);
  else if (t > 1 || ccd_eq(t, 1.0)) copy3(w, b);
  else { scl3(w, d, t); add3(w, w, x0); }
  copy3(wit, w);
  sub3(e, w, P);
  return dot3(e, e);
}
/* ccdVec3PointTriDist2 */
static double point_tri_dist2(const double *P, const double *x0, const double *B, const double *C, double *wit) {
  double d1[3], d2[3], a[3];
  sub3(d1, B, x0); sub3(d2, C, x0); sub3(a, x0, P);
  double v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(a, d1), q = dot3(a, d2), r = dot3(d1, d2);
  double s, t, dist, dd = w * v - r * r;
  if (ccd_is_zero(dd)) s = t = -1;
  else { s = (q * r - w * p) / dd; t = (-s * r - q) / w; }
  if ((ccd_is_zero(s) || s > 0) && (ccd_eq(s, 1.0) || s < 1) && (ccd_is_zero(t) || t > 0) &&
      (ccd_eq(t, 1.0) || t < 1) && (ccd_eq(t + s, 1.0) || t + s < 1)) {
    double e[3];
    scl3(d1, d1, s); scl3(d2, d2, t);
    copy3(wit, x0); add3(wit, wit, d1); add3(wit, wit, d2);
    sub3(e, wit, P); dist = dot3(e, e);
  } else {
    double w2[3], dist2;
    dist = point_seg_dist2(P, x0, B, wit);
    dist2 = point_seg_dist2(P, x0, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
    dist2 = point_seg_dist2(P, B, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
  }
  return dist;
}
static void find_pos(const Supp *p, double *pos) {
  double dir[3], vec[3], b[4], sum, p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
  portal_dir(p, dir);
  cross3(vec, p[1].v, p[2].v); b[0] = dot3(vec, p[3].v);
  cross3(vec, p[3].v, p[2].v); b[1] = dot3(vec, p[0].v);
  cross3(vec, p[0].v, p[1].v); b[2] = dot3(vec, p[3].v);
  cross3(vec, p[2].v, p[1].v); b[3] = dot3(vec, p[0].v);
  sum = b[0] + b[1] + b[2] + b[3];
  if (ccd_is_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(vec, p[2].v, p[3].v); b[1] = dot3(vec, dir);
    cross3(vec, p[3].v, p[1].v); b[2] = dot3(vec, dir);
    cross3(vec, p[1].v, p[2].v); b[3] = dot3(vec, dir);
    sum = b[1] + b[2] + b[3];
  }
  double inv = 1.0 / sum;
  for (int i = 0; i < 4; i++) { addscl3(p1, p[i].v1, b[i]); addscl3(p2, p[i].v2, b[i]); }
  scl3(p1, p1, inv); scl3(p2, p2, inv);
  add3(pos, p1, p2); scl3(pos, pos, 0.5);
}
/* returns 0 portal found, 1 origin on v1, 2 origin on v0-v1 segment, -1 no intersection */
static int discover_portal(const CObj *o1, const CObj *o2, Supp *p) {
  double dir[3], va[3], vb[3], dot, zero[3] = {0, 0, 0};
  obj_center(o1, p[0].v1); obj_center(o2, p[0].v2);
  sub3(p[0].v, p[0].v1, p[0].v2);
  if (ccd_vec_eq(p[0].v, zero)) p[0].v[0] += CCD_EPS * 10.0;
  scl3(dir, p[0].v, -1.0); ccd_normalize(dir);
  mink_support(o1, o2, dir, &p[1]);
  dot = dot3(p[1].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return -1;
  cross3(dir, p[0].v, p[1].v);
  if (ccd_is_zero(dot3(dir, dir))) return ccd_vec_eq(p[1].v, zero) ? 1 : 2;
  ccd_normalize(dir);
  mink_support(o1, o2, dir, &p[2]);
  dot = dot3(p[2].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return -1;
  sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
  cross3(dir, va, vb); ccd_normalize(dir);
  if (dot3(dir, p[0].v) > 0) { Supp t = p[1]; p[1] = p[2]; p[2] = t; scl3(dir, dir, -1.0); }
  for (;;) {
    int cont = 0;
    mink_support(o1, o2, dir, &p[3]);
    dot = dot3(p[3].v, dir);
    if (ccd_is_zero(dot) || dot < 0) return -1;
    cross3(va, p[1].v, p[3].v); dot = dot3(va, p[0].v);
    if (dot < 0 && !ccd_is_zero(dot)) { p[2] = p[3]; cont = 1; }
    if (!cont) {
      cross3(va, p[3].v, p[2].v); dot = dot3(va, p[0].v);
      if (dot < 0 && !ccd_is_zero(dot)) { p[1] = p[3]; cont = 1; }
    }
    if (!cont) return 0;
    sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
    cross3(dir, va, vb); ccd_normalize(dir);
  }
}
static int refine_portal(const CObj *o1, const CObj *o2, Supp *p, double tol) {
  double dir[3], dot; Supp v4;
  for (;;) {
    portal_dir(p, dir);
    dot = dot3(dir, p[1].v);
    if (ccd_is_zero(dot) || dot > 0) return 0; /* portal encapsulates origin */
    mink_support(o1, o2, dir, &v4);
    dot = dot3(v4.v, dir);
    if (!(ccd_is_zero(dot) || dot > 0) || portal_reach_tol(p, &v4, dir, tol)) return -1;
    expand_portal(p, &v4);
  }
}
static void find_penetr(const CObj *o1, const CObj *o2, Supp *p, double tol, int max_iter,
                        double *depth, double *pdir, double *pos) {
  double dir[3], zero[3] = {0, 0, 0}; Supp v4; unsigned long it = 0;
  for (;;) {
    portal_dir(p, dir);
    mink_support(o1, o2, dir, &v4);
    if (portal_reach_tol(p, &v4, dir, tol) || it > (unsigned long)max_iter) {
      *depth = sqrt(point_tri_dist2(zero, p[1].v, p[2].v, p[3].v, pdir));
      if (ccd_is_zero(pdir[0]) && ccd_is_zero(pdir[1]) && ccd_is_zero(pdir[2])) copy3(pdir, dir);
      ccd_normalize(pdir);
      find_pos(p, pos);
      return;
    }
    expand_portal(p, &v4);
    it++;
  }
}
static int mpr_penetration(const CObj *o1, const CObj *o2, double tol, int max_iter,
                           double *depth, double *dir, double *pos) {
  Supp p[4];
  int res = discover_portal(o1, o2, p);
  if (res < 0) return -1;
  if (res == 1) { /* findPenetrTouch */
    *depth = 0; dir[0] = dir[1] = dir[2] = 0;
    add3(pos, p[1].v1, p[1].v2); scl3(pos, pos, 0.5);
  } else if (res == 2) { /* findPenetrSegment */
    add3(pos, p[1].v1, p[1].v2); scl3(pos, pos, 0.5);
    copy3(dir, p[1].v); *depth = sqrt(dot3(dir, dir)); ccd_normalize(dir);
  } else {
    if (refine_portal(o1, o2, p, tol) < 0) return -1;
    find_penetr(o1, o2, p, tol, max_iter, depth, dir, pos);
  }
  return 0;
}
int tsgo_mpr(int type1, const double *pos1, const double *mat1, const double *size1,
             int type2, const double *pos2, const double *mat2, const double *size2,
             double tolerance, int max_iterations, double *depth, double *dir, double *pos) {
  CObj a = {type1, pos1, mat1, size1}, b = {type2, pos2, mat2, size2};
  return mpr_penetration(&a, &b, tolerance, max_iterations, depth, dir, pos) == 0;
}

/* ------------------------------------------------------------------ collision (mj_collision + narrow phase) */
static TsgoContact *new_contact(TsgoData *d, int g1, int g2, int b1, int b2) {
  if (d->ncon >= TSGO_MAXCON) { d->con_overflow = 1; return 0; }
  TsgoContact *c = &d->contact[d->ncon++];
  memset(c, 0, sizeof(*c));
  c->geom1 = g1; c->geom2 = g2; c->body1 = b1; c->body2 = b2; c->efc_address = -1;
  return c;
}
/* mjc_PlaneSphere */
static void plane_sphere(const TsgModel *m, TsgoData *d, int b, int g) {
  const double *n = (const double[]){m->floor_mat[2], m->floor_mat[5], m->floor_mat[8]};
  const double *c = d->geom_xpos[b][g];
  double r = m->geom_size[b][g][0], tmp[3];
  sub3(tmp, c, m->floor_pos);
  double cdist = dot3(tmp, n);
  if (cdist > r) return; /* margin 0 */
  TsgoContact *con = new_contact(d, 0, 1 + 5 * b + g, 0, 1 + b);
  if (!con) return;
  con->dist = cdist - r;
  copy3(con->pos, c); addscl3(con->pos, n, -con->dist / 2 - r);
  copy3(con->frame, n);
}
/* mjc_PlaneCylinder (up to 4 contacts) */
static void plane_cylinder(const TsgModel *m, TsgoData *d, int b, int g) {
  double normal[3] = {m->floor_mat[2], m->floor_mat[5], m->floor_mat[8]};
  const double *mat2 = d->geom_xmat[b][g], *pos2 = d->geom_xpos[b][g];
  double axis[3] = {mat2[2], mat2[5], mat2[8]}, vec[3], radius = m->geom_size[b][g][0], half = m->geom_size[b][g][1];
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { scl3(axis, axis, -1); prjaxis = -prjaxis; }
  sub3(vec, pos2, m->floor_pos);
  double dist0 = dot3(vec, normal);
  scl3(vec, axis, prjaxis); sub3(vec, vec, normal);
  double len_sqr = dot3(vec, vec);
  if (len_sqr >= MINVAL * MINVAL) scl3(vec, vec, radius / sqrt(len_sqr));
  else { vec[0] = mat2[0] * radius; vec[1] = mat2[3] * radius; vec[2] = mat2[6] * radius; }
  double prjvec = dot3(vec, normal);
  scl3(axis, axis, half); prjaxis *= half;
  int g2 = 1 + 5 * b + g;
  TsgoContact *con;
  if (dist0 + prjaxis + prjvec <= 0) {
    if (!(con = new_contact(d, 0, g2, 0, 1 + b))) return;
    con->dist = dist0 + prjaxis + prjvec;
    add3(con->pos, pos2, vec); add3(con->pos, con->pos, axis); addscl3(con->pos, normal, -con->dist * 0.5);
    copy3(con->frame, normal);
  } else return;
  if (dist0 - prjaxis + prjvec <= 0) {
    if (!(con = new_contact(d, 0, g2, 0, 1 + b))) return;
    con->dist = dist0 - prjaxis + prjvec;
    add3(con->pos, pos2, vec); sub3(con->pos, con->pos, axis); addscl3(con->pos, normal, -con->dist * 0.5);
    copy3(con->frame, normal);
  }
  double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= 0) {
    double vec1[3];
    cross3(vec1, vec, axis); normalize3(vec1); scl3(vec1, vec1, radius * sqrt(3.0) * 0.5);
    for (int s = 0; s < 2; s++) {
      if (!(con = new_contact(d, 0, g2, 0, 1 + b))) return;
      con->dist = dist0 + prjaxis + prjvec1;
      add3(con->pos, pos2, axis); addscl3(con->pos, vec1, s ? -1.0 : 1.0); addscl3(con->pos, vec, -0.5);
      addscl3(con->pos, normal, -con->dist * 0.5);
      copy3(con->frame, normal);
    }
  }
}
/* mjc_SphereSphere */
static void sphere_sphere(const TsgModel *m, TsgoData *d, int b1, int g1, int b2, int g2) {
  const double *c1 = d->geom_xpos[b1][g1], *c2 = d->geom_xpos[b2][g2];
  double r1 = m->geom_size[b1][g1][0], r2 = m->geom_size[b2][g2][0], n[3];
  sub3(n, c2, c1);
  double len = normalize3(n);
  double dist = len - r1 - r2;
  if (dist > 0) return;
  TsgoContact *con = new_contact(d, 1 + 5 * b1 + g1, 1 + 5 * b2 + g2, 1 + b1, 1 + b2);
  if (!con) return;
  con->dist = dist;
  copy3(con->pos, c1); addscl3(con->pos, n, r1 + dist / 2);
  copy3(con->frame, n);
}
/* mjc_Convex (+ mjc_fixNormal for the sphere) ; (b1,g1) has the lower geom type */
static void convex_pair(const TsgModel *m, TsgoData *d, int b1, int g1, int b2, int g2) {
  CObj o1 = {m->geom_type[b1][g1], d->geom_xpos[b1][g1], d->geom_xmat[b1][g1], m->geom_size[b1][g1]};
  CObj o2 = {m->geom_type[b2][g2], d->geom_xpos[b2][g2], d->geom_xmat[b2][g2], m->geom_size[b2][g2]};
  double depth, dir[3], pos[3], zero[3] = {0, 0, 0};
  d->mpr_calls++;
  if (mpr_penetration(&o1, &o2, m->mpr_tolerance, m->mpr_iterations, &depth, dir, pos) != 0) return;
  if (ccd_vec_eq(dir, zero)) return;
  TsgoContact *con = new_contact(d, 1 + 5 * b1 + g1, 1 + 5 * b2 + g2, 1 + b1, 1 + b2);
  if (!con) return;
  con->dist = -depth;
  copy3(con->frame, dir);
  copy3(con->pos, pos);
  if ((m->flags & TSG_FLAG_FIXNORMAL) && o1.type == TSG_GEOM_SPHERE) {
    double n[3];
    sub3(n, pos, o1.pos);
    if (norm3(n) > MINVAL) { normalize3(n); copy3(con->frame, n); }
  }
}
/* mj_collideGeoms: type ordering, bounding-sphere filter, narrow phase */
static void collide_bar_geoms(const TsgModel *m, TsgoData *d, int b1, int g1, int b2, int g2) {
  double dif[3], bound = m->geom_rbound[b1][g1] + m->geom_rbound[b2][g2];
  sub3(dif, d->geom_xpos[b1][g1], d->geom_xpos[b2][g2]);
  if (dot3(dif, dif) > bound * bound) return; /* mj_filterSphere */
  int t1 = m->geom_type[b1][g1], t2 = m->geom_type[b2][g2];
  if (t1 > t2) { int tb = b1, tg = g1; b1 = b2; g1 = g2; b2 = tb; g2 = tg; t1 = t2; t2 = m->geom_type[b2][g2]; }
  if (t1 == TSG_GEOM_SPHERE && t2 == TSG_GEOM_SPHERE) sphere_sphere(m, d, b1, g1, b2, g2);
  else convex_pair(m, d, b1, g1, b2, g2);
}
/* mjc_ConvexHField: geom (b,g) against the height field */
static void hfield_geom(const TsgModel *m, TsgoData *d, int b, int g) {
  const double *size1 = m->hf_size, *pos1 = m->floor_pos, *mat1 = m->floor_mat;
  int nrow = m->hf_nrow, ncol = m->hf_ncol;
  double vec[3], pos[3], mat[9], r2 = m->geom_rbound[b][g];
  sub3(vec, d->geom_xpos[b][g], pos1); mulMTV(pos, mat1, vec);
  for (int i = 0; i < 2; i++) if (size1[i] < pos[i] - r2 || -size1[i] > pos[i] + r2) return;
  if (size1[2] < pos[2] - r2) return;
  if (-size1[3] > pos[2] + r2) return;
  /* geom frame expressed in the hfield frame: mat = mat1^T * mat2 */
  const double *mat2 = d->geom_xmat[b][g];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++)
    mat[3 * i + j] = mat1[i] * mat2[j] + mat1[3 + i] * mat2[3 + j] + mat1[6 + i] * mat2[6 + j];
  CObj o2 = {m->geom_type[b][g], pos, mat, m->geom_size[b][g]};
  double s[3], dirs[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
  double xmax, xmin, ymax, ymin, zmax, zmin;
  obj_support(&o2, dirs[0], s); xmax = s[0]; obj_support(&o2, dirs[1], s); xmin = s[0];
  obj_support(&o2, dirs[2], s); ymax = s[1]; obj_support(&o2, dirs[3], s); ymin = s[1];
  obj_support(&o2, dirs[4], s); zmax = s[2]; obj_support(&o2, dirs[5], s); zmin = s[2];
  if (xmin > size1[0] || xmax < -size1[0] || ymin > size1[1] || ymax < -size1[1] || zmin > size1[2] || zmax < -size1[3]) return;
  int cmin = (int)floor((xmin + size1[0]) / (2 * size1[0]) * (ncol - 1));
  int cmax = (int)ceil((xmax + size1[0]) / (2 * size1[0]) * (ncol - 1));
  int rmin = (int)floor((ymin + size1[1]) / (2 * size1[1]) * (nrow - 1));
  int rmax = (int)ceil((ymax + size1[1]) / (2 * size1[1]) * (nrow - 1));
  if (cmin < 0) cmin = 0;
  if (cmax > ncol - 1) cmax = ncol - 1;
  if (rmin < 0) rmin = 0;
  if (rmax > nrow - 1) rmax = nrow - 1;
  double dx = (2.0 * size1[0]) / (ncol - 1), dy = (2.0 * size1[1]) / (nrow - 1);
  double prism[18];
  CObj o1 = {OBJ_PRISM, 0, 0, prism};
  int dr[2] = {1, 0}, cnt = 0;
  prism[2] = prism[5] = prism[8] = -size1[3];
  for (int r = rmin; r < rmax; r++) {
    int nvert = 0;
    for (int c = cmin; c <= cmax; c++)
      for (int i = 0; i < 2; i++) {
        /* addVert: shift, then append (x, y, z) */
        double x = dx * c - size1[0], y = dy * (r + dr[i]) - size1[1];
        double z = (double)m->hf_data[(r + dr[i]) * ncol + c] * size1[2];
        prism[0] = prism[3]; prism[1] = prism[4]; prism[3] = prism[6]; prism[4] = prism[7];
        memcpy(prism + 9, prism + 12, 3 * sizeof(double)); memcpy(prism + 12, prism + 15, 3 * sizeof(double));
        prism[6] = x; prism[7] = y; prism[15] = x; prism[16] = y; prism[17] = z;
        nvert++;
        if (nvert > 2) {
          if (prism[11] < zmin && prism[14] < zmin && prism[17] < zmin) continue;
          double depth, dir[3], p[3], zero[3] = {0, 0, 0};
          d->mpr_calls++;
          if (mpr_penetration(&o1, &o2, m->mpr_tolerance, m->mpr_iterations, &depth, dir, p) == 0 && !ccd_vec_eq(dir, zero)) {
            TsgoContact *con = new_contact(d, 0, 1 + 5 * b + g, 0, 1 + b);
            if (!con) return;
            con->dist = -depth;
            mulMV(con->frame, mat1, dir);
            mulMV(con->pos, mat1, p); add3(con->pos, con->pos, pos1);
            if ((m->flags & TSG_FLAG_FIXNORMAL) && o2.type == TSG_GEOM_SPHERE) {
              double n[3]; /* geom2 is the sphere: normal = -(pos - centre) */
              sub3(n, d->geom_xpos[b][g], con->pos);
              if (norm3(n) > MINVAL) { normalize3(n); copy3(con->frame, n); }
            }
            if (++cnt >= 50) return; /* mjMAXCONPAIR */
          }
        }
      }
  }
}
/* mju_makeFrame */
static void make_frame(double *f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) {
    f[3] = f[4] = f[5] = 0;
    if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  }
  double t = dot3(f, f + 3);
  addscl3(f + 3, f, -t);
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
void tsgo_collision(const TsgModel *m, TsgoData *d) {
  d->ncon = 0; d->mpr_calls = 0; d->con_overflow = 0;
  /* body pairs in MuJoCo order: (world,bar0..2) then (bar_i,bar_j), geoms in model order */
  for (int b = 0; b < TSG_NBAR; b++)
    for (int g = 0; g < TSG_NGEOM_BAR; g++) {
      if (m->floor_type == TSG_FLOOR_PLANE) {
        /* plane bounding-sphere filter, then narrow phase */
        double n[3] = {m->floor_mat[2], m->floor_mat[5], m->floor_mat[8]}, v[3];
        sub3(v, d->geom_xpos[b][g], m->floor_pos);
        if (dot3(v, n) > m->geom_rbound[b][g]) continue;
        if (m->geom_type[b][g] == TSG_GEOM_SPHERE) plane_sphere(m, d, b, g); else plane_cylinder(m, d, b, g);
      } else hfield_geom(m, d, b, g);
    }
  for (int b1 = 0; b1 < TSG_NBAR; b1++)
    for (int b2 = b1 + 1; b2 < TSG_NBAR; b2++)
      for (int g1 = 0; g1 < TSG_NGEOM_BAR; g1++)
        for (int g2 = 0; g2 < TSG_NGEOM_BAR; g2++) collide_bar_geoms(m, d, b1, g1, b2, g2);
  for (int i = 0; i < d->ncon; i++) {
    TsgoContact *c = &d->contact[i];
    c->frame[3] = c->frame[4] = c->frame[5] = 0;
    make_frame(c->frame);
    c->exclude = (c->dist >= 0); /* includemargin = margin - gap = 0 */
  }
}

/* ------------------------------------------------------------------ constraints (mj_makeConstraint, mj_makeImpedance) */
static double impedance(const double *solimp, double pos) { /* getimpedance, margin 0 */
  double d0 = solimp[0], dw = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
  d0 = fmin(MAXIMP, fmax(MINIMP, d0)); dw = fmin(MAXIMP, fmax(MINIMP, dw));
  width = fmax(MINVAL, width); mid = fmin(MAXIMP, fmax(MINIMP, mid)); power = fmax(1, power);
  if (d0 == dw || width <= MINVAL) return 0.5 * (d0 + dw);
  double x = fabs(pos) / width, y;
  if (x >= 1) return dw;
  if (x == 0) return d0;
  if (power == 1) y = x;
  else if (x <= mid) y = (1 / pow(mid, power - 1)) * pow(x, power);
  else y = 1 - (1 / pow(1 - mid, power - 1)) * pow(1 - x, power);
  return d0 + y * (dw - d0);
}
void tsgo_make_constraint(const TsgModel *m, TsgoData *d) {
  d->nefc = 0;
  double dmax = fmin(MAXIMP, fmax(MINIMP, m->solimp[1]));
  double K = -m->solref[0] / (dmax * dmax), B = -m->solref[1] / dmax; /* direct (negative) solref */
  for (int i = 0; i < d->ncon; i++) {
    TsgoContact *c = &d->contact[i];
    c->efc_address = -1;
    if (c->exclude) continue;
    if (d->nefc + 6 > TSGO_MAXEFC) { d->con_overflow = 1; break; }
    int a0 = d->nefc;
    c->efc_address = a0;
    /* mj_jacDifPair at the contact point, rotated into the contact frame */
    for (int r = 0; r < 6; r++) memset(d->efc_J[a0 + r], 0, sizeof(d->efc_J[0]));
    for (int side = 0; side < 2; side++) {
      int body = side ? c->body2 : c->body1;
      if (body == 0) continue;
      int b = body - 1;
      double s = side ? 1.0 : -1.0, rr[3], row[6];
      sub3(rr, c->pos, d->xpos[b]);
      for (int ax = 0; ax < 3; ax++) {
        const double *a = c->frame + 3 * ax;
        jac_point_row(row, d->xmat[b], rr, a);
        for (int k = 0; k < 6; k++) d->efc_J[a0 + ax][6 * b + k] += s * row[k];
        mulMTV(row, d->xmat[b], a); /* rotational rows: [0, R^T a] */
        for (int k = 0; k < 3; k++) d->efc_J[a0 + 3 + ax][6 * b + 3 + k] += s * row[k];
      }
    }
    /* mj_diagApprox (elliptic): translational invweight of both bodies */
    double tran = 0;
    if (c->body1) tran += m->body_invweight0[c->body1 - 1][0];
    if (c->body2) tran += m->body_invweight0[c->body2 - 1][0];
    double imp = impedance(m->solimp, c->dist);
    double R0 = fmax(MINVAL, (1 - imp) / imp * tran);
    double R1 = R0 / m->impratio;
    for (int r = 0; r < 6; r++) {
      double vel = 0;
      for (int k = 0; k < NV; k++) vel += d->efc_J[a0 + r][k] * d->qvel[k];
      d->efc_vel[a0 + r] = vel;
      d->efc_pos[a0 + r] = r ? 0 : c->dist;
      double Rr = R0;
      if (r > 0) Rr = R1 * m->friction[0] * m->friction[0] / (m->friction[r - 1] * m->friction[r - 1]);
      d->efc_R[a0 + r] = Rr;
      d->efc_D[a0 + r] = 1 / Rr;
      /* mj_referenceConstraint: friction rows have K = 0 */
      d->efc_aref[a0 + r] = -B * vel - (r ? 0 : K * imp * c->dist);
    }
    d->nefc += 6;
  }
}

/* ------------------------------------------------------------------ smooth dynamics */
static void fwd_velocity_actuation(const TsgModel *m, TsgoData *d) {
  /* mj_fwdVelocity: ten_velocity, mj_passive, mj_rne(flg_acc=0) */
  memset(d->qfrc_passive, 0, sizeof(d->qfrc_passive));
  for (int t = 0; t < TSG_NTEN; t++) {
    double v = 0, frc = 0;
    for (int k = 0; k < NV; k++) v += d->ten_J[t][k] * d->qvel[k];
    d->ten_velocity[t] = v;
    if (m->ten_stiffness[t] > 0) {
      double L = d->ten_length[t];
      if (L > m->ten_lengthspring[t][1]) frc = m->ten_stiffness[t] * (m->ten_lengthspring[t][1] - L);
      else if (L < m->ten_lengthspring[t][0]) frc = m->ten_stiffness[t] * (m->ten_lengthspring[t][0] - L);
    }
    frc -= m->ten_damping[t] * v;
    for (int k = 0; k < NV; k++) d->qfrc_passive[k] += d->ten_J[t][k] * frc;
  }
  for (int b = 0; b < TSG_NBAR; b++) {
    const double *w = d->qvel + 6 * b + 3, *I = m->body_inertia[b];
    double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]}, gy[3];
    cross3(gy, w, Iw);
    for (int k = 0; k < 3; k++) {
      d->qfrc_bias[6 * b + k] = -m->body_mass[b] * m->gravity[k];
      d->qfrc_bias[6 * b + 3 + k] = gy[k];
    }
  }
  /* mj_fwdActuation */
  memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
  for (int i = 0; i < TSG_NACT; i++) {
    int t = m->act_tendon[i];
    double ctrl = d->ctrl[i], input;
    if (m->ctrllimited) ctrl = fmin(m->ctrlrange[1], fmax(m->ctrlrange[0], ctrl));
    if (m->act_dyntype == TSG_DYN_FILTER) {
      d->act_dot[i] = (ctrl - d->act[i]) / fmax(MINVAL, m->act_dynprm0);
      input = d->act[i];
    } else { d->act_dot[i] = 0; input = ctrl; }
    double f = m->act_gain * input + m->act_bias[0] + m->act_bias[1] * d->ten_length[t] + m->act_bias[2] * d->ten_velocity[t];
    if (m->forcelimited) f = fmin(m->forcerange[1], fmax(m->forcerange[0], f));
    d->actuator_force[i] = f;
    for (int k = 0; k < NV; k++) d->qfrc_actuator[k] += d->ten_J[t][k] * f;
  }
  /* mj_fwdAcceleration */
  for (int k = 0; k < NV; k++) {
    int b = k / 6, j = k % 6;
    double M = j < 3 ? m->body_mass[b] : m->body_inertia[b][j - 3];
    d->qfrc_smooth[k] = d->qfrc_passive[k] - d->qfrc_bias[k] + d->qfrc_actuator[k];
    d->qacc_smooth[k] = d->qfrc_smooth[k] / M;
  }
}
static double Mdiag(const TsgModel *m, int k) { int b = k / 6, j = k % 6; return j < 3 ? m->body_mass[b] : m->body_inertia[b][j - 3]; }

/* ------------------------------------------------------------------ Newton solver (mj_solNewton / mj_solPrimal) */
typedef struct {
  const TsgModel *m; TsgoData *d;
  double Ma[NV], Jaref[TSGO_MAXEFC], grad[NV], Mgrad[NV], search[NV], Mv[NV], Jv[TSGO_MAXEFC];
  double coneH[TSGO_MAXCON][36];
  double quad[TSGO_MAXEFC][3], quadGauss[3];
  double cU0[TSGO_MAXCON], cV0[TSGO_MAXCON], cUU[TSGO_MAXCON], cUV[TSGO_MAXCON], cVV[TSGO_MAXCON], cDm[TSGO_MAXCON];
  double cost, gauss, mu;
  int ncone, ls_evals;
} Primal;

/* mj_constraintUpdate: efc_force, efc_state, constraint cost, optional cone Hessians */
static double constraint_update(const TsgModel *m, TsgoData *d, const double *jar, double (*coneH)[36], int *ncone_out) {
  double s = 0, mu = m->friction[0] / sqrt(m->impratio);
  const double *fr = m->friction;
  int ncone = 0;
  for (int i = 0; i < d->nefc; i += 6) {
    const double *D = d->efc_D + i;
    double U[6], N, T = 0;
    U[0] = jar[i] * mu;
    for (int j = 1; j < 6; j++) { U[j] = jar[i + j] * fr[j - 1]; T += U[j] * U[j]; }
    N = U[0]; T = sqrt(T);
    if (N >= mu * T || (T <= 0 && N >= 0)) { /* top zone */
      for (int j = 0; j < 6; j++) { d->efc_force[i + j] = 0; d->efc_state[i + j] = ST_SATISFIED; }
    } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) { /* bottom zone */
      for (int j = 0; j < 6; j++) {
        d->efc_force[i + j] = -D[j] * jar[i + j]; d->efc_state[i + j] = ST_QUADRATIC;
        s += 0.5 * D[j] * jar[i + j] * jar[i + j];
      }
    } else { /* middle zone */
      double Dm = D[0] / (mu * mu * (1 + mu * mu)), NT = N - mu * T;
      s += 0.5 * Dm * NT * NT;
      d->efc_force[i] = -Dm * NT * mu;
      for (int j = 1; j < 6; j++) d->efc_force[i + j] = -d->efc_force[i] / T * U[j] * fr[j - 1];
      for (int j = 0; j < 6; j++) d->efc_state[i + j] = ST_CONE;
      ncone++;
      if (coneH) {
        double *H = coneH[i / 6], t;
        memset(H, 0, 36 * sizeof(double));
        t = -mu / T; H[0] = 1;
        for (int j = 1; j < 6; j++) H[j] = t * U[j];
        t = mu * N / (T * T * T);
        for (int k = 1; k < 6; k++) for (int j = k; j < 6; j++) H[6 * k + j] = t * U[j] * U[k];
        t = mu * mu - mu * N / T;
        for (int j = 1; j < 6; j++) H[7 * j] += t;
        for (int k = 0; k < 6; k++) for (int j = k; j < 6; j++)
          H[6 * k + j] *= Dm * (k == 0 ? mu : fr[k - 1]) * (j == 0 ? mu : fr[j - 1]);
        for (int k = 0; k < 6; k++) for (int j = k + 1; j < 6; j++) H[6 * j + k] = H[6 * k + j];
      }
    }
  }
  if (ncone_out) *ncone_out = ncone;
  return s;
}
static void primal_update_constraint(Primal *c) {
  TsgoData *d = c->d;
  double s = constraint_update(c->m, d, c->Jaref, c->coneH, &c->ncone);
  for (int k = 0; k < NV; k++) {
    double f = 0;
    for (int i = 0; i < d->nefc; i++) f += d->efc_J[i][k] * d->efc_force[i];
    d->qfrc_constraint[k] = f;
  }
  c->gauss = 0;
  for (int k = 0; k < NV; k++) c->gauss += 0.5 * (c->Ma[k] - d->qfrc_smooth[k]) * (d->qacc[k] - d->qacc_smooth[k]);
  c->cost = s + c->gauss;
}
/* MakeHessian + HessianCone + factor + solve: Mgrad = H^-1 grad */
static void primal_update_gradient(Primal *c) {
  TsgoData *d = c->d; const TsgModel *m = c->m;
  double H[NV][NV];
  for (int k = 0; k < NV; k++) c->grad[k] = c->Ma[k] - d->qfrc_smooth[k] - d->qfrc_constraint[k];
  memset(H, 0, sizeof(H));
  for (int k = 0; k < NV; k++) H[k][k] = Mdiag(m, k);
  for (int i = 0; i < d->nefc; i++)
    if (d->efc_state[i] == ST_QUADRATIC)
      for (int a = 0; a < NV; a++) {
        double ja = d->efc_J[i][a] * d->efc_D[i];
        if (ja != 0) for (int b = 0; b <= a; b++) H[a][b] += ja * d->efc_J[i][b];
      }
  for (int i = 0; i < d->nefc; i += 6)
    if (d->efc_state[i] == ST_CONE) {
      const double *C = c->coneH[i / 6];
      double CJ[6][NV];
      for (int r = 0; r < 6; r++) for (int k = 0; k < NV; k++) {
        double v = 0;
        for (int s = 0; s < 6; s++) v += C[6 * r + s] * d->efc_J[i + s][k];
        CJ[r][k] = v;
      }
      for (int a = 0; a < NV; a++) for (int b = 0; b <= a; b++) {
        double v = 0;
        for (int r = 0; r < 6; r++) v += d->efc_J[i + r][a] * CJ[r][b];
        H[a][b] += v;
      }
    }
  /* dense Cholesky H = L L^T (lower), then solve */
  for (int j = 0; j < NV; j++) {
    double s = H[j][j];
    for (int k = 0; k < j; k++) s -= H[j][k] * H[j][k];
    double piv = sqrt(fmax(s, MINVAL));
    H[j][j] = piv;
    for (int i = j + 1; i < NV; i++) {
      double t = H[i][j];
      for (int k = 0; k < j; k++) t -= H[i][k] * H[j][k];
      H[i][j] = t / piv;
    }
  }
  double y[NV];
  for (int i = 0; i < NV; i++) { double t = c->grad[i]; for (int k = 0; k < i; k++) t -= H[i][k] * y[k]; y[i] = t / H[i][i]; }
  for (int i = NV - 1; i >= 0; i--) { double t = y[i]; for (int k = i + 1; k < NV; k++) t -= H[k][i] * c->Mgrad[k]; c->Mgrad[i] = t / H[i][i]; }
}
typedef struct { double alpha, cost, deriv[2]; } LsPnt;
/* PrimalPrepare */
static void primal_prepare(Primal *c) {
  TsgoData *d = c->d; const TsgModel *m = c->m;
  double mu = m->friction[0] / sqrt(m->impratio);
  c->mu = mu;
  c->quadGauss[0] = c->gauss; c->quadGauss[1] = 0; c->quadGauss[2] = 0;
  for (int k = 0; k < NV; k++) {
    c->quadGauss[1] += c->search[k] * c->Ma[k] - d->qfrc_smooth[k] * c->search[k];
    c->quadGauss[2] += 0.5 * c->search[k] * c->Mv[k];
  }
  for (int i = 0; i < d->nefc; i++) {
    double D = d->efc_D[i], ja = c->Jaref[i], jv = c->Jv[i];
    c->quad[i][0] = 0.5 * D * ja * ja; c->quad[i][1] = D * ja * jv; c->quad[i][2] = 0.5 * D * jv * jv;
  }
  for (int i = 0; i < d->nefc; i += 6) { /* elliptic: accumulate rows into the first, cone quantities */
    int ci = i / 6;
    for (int j = 1; j < 6; j++) for (int k = 0; k < 3; k++) c->quad[i][k] += c->quad[i + j][k];
    double U[6], V[6];
    U[0] = c->Jaref[i] * mu; V[0] = c->Jv[i] * mu;
    for (int j = 1; j < 6; j++) { U[j] = c->Jaref[i + j] * m->friction[j - 1]; V[j] = c->Jv[i + j] * m->friction[j - 1]; }
    c->cU0[ci] = U[0]; c->cV0[ci] = V[0]; c->cUU[ci] = c->cUV[ci] = c->cVV[ci] = 0;
    for (int j = 1; j < 6; j++) { c->cUU[ci] += U[j] * U[j]; c->cUV[ci] += U[j] * V[j]; c->cVV[ci] += V[j] * V[j]; }
    c->cDm[ci] = d->efc_D[i] / (mu * mu * (1 + mu * mu));
  }
}
/* PrimalEval */
static void primal_eval(Primal *c, LsPnt *p) {
  double a = p->alpha, mu = c->mu;
  double cost = a * a * c->quadGauss[2] + a * c->quadGauss[1] + c->quadGauss[0];
  double d0 = 2 * a * c->quadGauss[2] + c->quadGauss[1], d1 = 2 * c->quadGauss[2];
  for (int i = 0; i < c->d->nefc; i += 6) {
    int ci = i / 6;
    const double *q = c->quad[i];
    double N = c->cU0[ci] + a * c->cV0[ci];
    double Tsqr = c->cUU[ci] + a * (2 * c->cUV[ci] + a * c->cVV[ci]);
    int bottom = 0;
    if (Tsqr <= 0) { if (N < 0) bottom = 1; }
    else {
      double T = sqrt(Tsqr);
      if (N >= mu * T) { /* top: nothing */ }
      else if (mu * N + T <= 0) bottom = 1;
      else {
        double N1 = c->cV0[ci], T1 = (c->cUV[ci] + a * c->cVV[ci]) / T;
        double T2 = c->cVV[ci] / T - (c->cUV[ci] + a * c->cVV[ci]) * T1 / (T * T);
        double NT = N - mu * T, Dm = c->cDm[ci];
        cost += 0.5 * Dm * NT * NT;
        d0 += Dm * NT * (N1 - mu * T1);
        d1 += Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NT * (-mu * T2));
      }
    }
    if (bottom) { cost += a * a * q[2] + a * q[1] + q[0]; d0 += 2 * a * q[2] + q[1]; d1 += 2 * q[2]; }
  }
  if (d1 <= 0) d1 = MINVAL;
  p->cost = cost; p->deriv[0] = d0; p->deriv[1] = d1;
  c->ls_evals++;
}
static int update_bracket(Primal *c, LsPnt *p, const LsPnt cand[3], LsPnt *pnext) {
  int flag = 0;
  for (int i = 0; i < 3; i++) {
    if (p->deriv[0] < 0 && cand[i].deriv[0] < 0 && p->deriv[0] < cand[i].deriv[0]) { *p = cand[i]; flag = 1; }
    else if (p->deriv[0] > 0 && cand[i].deriv[0] > 0 && p->deriv[0] > cand[i].deriv[0]) { *p = cand[i]; flag = 2; }
  }
  if (flag) { pnext->alpha = p->alpha - p->deriv[0] / p->deriv[1]; primal_eval(c, pnext); }
  return flag;
}
/* PrimalSearch: exact 1-D minimisation along `search`; returns alpha */
static double primal_search(Primal *c) {
  const TsgModel *m = c->m; TsgoData *d = c->d;
  LsPnt p0, p1, p2, pmid, p1next, p2next;
  double snorm = 0;
  for (int k = 0; k < NV; k++) snorm += c->search[k] * c->search[k];
  snorm = sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double scale = 1 / (m->meaninertia * NV);
  double gtol = m->tolerance * m->ls_tolerance * snorm / scale;
  for (int k = 0; k < NV; k++) c->Mv[k] = Mdiag(m, k) * c->search[k];
  for (int i = 0; i < d->nefc; i++) { double v = 0; for (int k = 0; k < NV; k++) v += d->efc_J[i][k] * c->search[k]; c->Jv[i] = v; }
  primal_prepare(c);
  int evals0 = c->ls_evals;
#define LSITER (c->ls_evals - evals0)
  p0.alpha = 0; primal_eval(c, &p0);
  p1.alpha = p0.alpha - p0.deriv[0] / p0.deriv[1]; primal_eval(c, &p1);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  int dir = p1.deriv[0] < 0 ? 1 : -1, p2update = 0;
  p2 = p1;
  while (p1.deriv[0] * dir <= -gtol && LSITER < m->ls_iterations) {
    p2 = p1; p2update = 1;
    p1.alpha -= p1.deriv[0] / p1.deriv[1]; primal_eval(c, &p1);
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  }
  if (LSITER >= m->ls_iterations) return p1.alpha;
  if (!p2update) return p1.alpha;
  p2next = p1;
  p1next.alpha = p1.alpha - p1.deriv[0] / p1.deriv[1]; primal_eval(c, &p1next);
  while (LSITER < m->ls_iterations) {
    pmid.alpha = 0.5 * (p1.alpha + p2.alpha); primal_eval(c, &pmid);
    LsPnt cand[3] = {p1next, p2next, pmid};
    int best = -1; double bestcost = 0;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i].deriv[0]) < gtol && (best == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; best = i; }
    if (best >= 0) return cand[best].alpha;
    int b1 = update_bracket(c, &p1, cand, &p1next);
    int b2 = update_bracket(c, &p2, cand, &p2next);
    if (!b1 && !b2) return pmid.alpha;
  }
#undef LSITER
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}
static void jar_of(const TsgoData *d, const double *qacc, double *jar) {
  for (int i = 0; i < d->nefc; i++) {
    double v = 0;
    for (int k = 0; k < NV; k++) v += d->efc_J[i][k] * qacc[k];
    jar[i] = v - d->efc_aref[i];
  }
}
static void sol_newton(const TsgModel *m, TsgoData *d) {
  static _Thread_local Primal ctx;
  Primal *c = &ctx;
  c->m = m; c->d = d; c->ls_evals = 0;
  double scale = 1 / (m->meaninertia * NV);
  for (int k = 0; k < NV; k++) c->Ma[k] = Mdiag(m, k) * d->qacc[k];
  jar_of(d, d->qacc, c->Jaref);
  primal_update_constraint(c);
  primal_update_gradient(c);
  for (int k = 0; k < NV; k++) c->search[k] = -c->Mgrad[k];
  int iter = 0;
  while (iter < m->iterations) {
    double alpha = primal_search(c);
    if (alpha == 0) break;
    for (int k = 0; k < NV; k++) { d->qacc[k] += alpha * c->search[k]; c->Ma[k] += alpha * c->Mv[k]; }
    for (int i = 0; i < d->nefc; i++) c->Jaref[i] += alpha * c->Jv[i];
    double oldcost = c->cost;
    primal_update_constraint(c);
    primal_update_gradient(c);
    for (int k = 0; k < NV; k++) c->search[k] = -c->Mgrad[k];
    double improvement = scale * (oldcost - c->cost), gn = 0;
    for (int k = 0; k < NV; k++) gn += c->grad[k] * c->grad[k];
    double gradient = scale * sqrt(gn);
    iter++;
    if (improvement < m->tolerance || gradient < m->tolerance) break;
  }
  d->solver_iter = iter; d->ls_evals = c->ls_evals; d->solver_cost = c->cost;
}
double tsgo_primal_cost(const TsgModel *m, TsgoData *d, const double *qacc, double *grad) {
  static _Thread_local double jar[TSGO_MAXEFC];
  double force_save[TSGO_MAXEFC]; int state_save[TSGO_MAXEFC];
  memcpy(force_save, d->efc_force, sizeof(force_save)); memcpy(state_save, d->efc_state, sizeof(state_save));
  jar_of(d, qacc, jar);
  double s = constraint_update(m, d, jar, 0, 0);
  for (int k = 0; k < NV; k++) {
    double Ma = Mdiag(m, k) * qacc[k];
    s += 0.5 * (Ma - d->qfrc_smooth[k]) * (qacc[k] - d->qacc_smooth[k]);
    if (grad) {
      double f = 0;
      for (int i = 0; i < d->nefc; i++) f += d->efc_J[i][k] * d->efc_force[i];
      grad[k] = Ma - d->qfrc_smooth[k] - f;
    }
  }
  memcpy(d->efc_force, force_save, sizeof(force_save)); memcpy(d->efc_state, state_save, sizeof(state_save));
  return s;
}
/* mj_fwdConstraint: warm start choice + Newton */
static void fwd_constraint(const TsgModel *m, TsgoData *d) {
  if (d->nefc == 0) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(d->qacc));
    memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(d->qacc));
    memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
    d->solver_iter = 0; d->ls_evals = 0;
    return;
  }
  double cost_ws = tsgo_primal_cost(m, d, d->qacc_warmstart, 0);
  double cost_sm = tsgo_primal_cost(m, d, d->qacc_smooth, 0);
  memcpy(d->qacc, cost_ws > cost_sm ? d->qacc_smooth : d->qacc_warmstart, sizeof(d->qacc));
  sol_newton(m, d);
  memcpy(d->qacc_warmstart, d->qacc, sizeof(d->qacc));
}

/* ------------------------------------------------------------------ mj_forward */
void tsgo_forward(const TsgModel *m, TsgoData *d) {
  tsgo_kinematics(m, d);
  tsgo_tendon(m, d);
  tsgo_collision(m, d);
  tsgo_make_constraint(m, d); /* uses qvel for efc_vel / aref (mj_referenceConstraint) */
  fwd_velocity_actuation(m, d);
  fwd_constraint(m, d);
}

/* ------------------------------------------------------------------ implicitfast + mj_advance */
static void implicit_advance(const TsgModel *m, TsgoData *d) {
  double h = m->timestep, qacc[NV];
  /* qDeriv = d(passive + actuator)/d qvel = sum_t B_t J_t' J_t, kept on the per-bar
   * 6x6 blocks only (D sparsity = dofs of one kinematic tree), unless CROSSBAR flag */
  double Bt[TSG_NTEN];
  for (int t = 0; t < TSG_NTEN; t++) Bt[t] = -m->ten_damping[t];
  for (int i = 0; i < TSG_NACT; i++) {
    double bv = m->act_bias[2];
    if (bv == 0) continue;
    if (m->forcelimited && !(m->flags & TSG_FLAG_ACTVEL_WHEN_CLAMPED)) {
      double f = d->actuator_force[i];
      if (f <= m->forcerange[0] || f >= m->forcerange[1]) continue;
    }
    Bt[m->act_tendon[i]] += bv;
  }
  double A[NV][NV];
  memset(A, 0, sizeof(A));
  for (int a = 0; a < NV; a++) for (int b = 0; b < NV; b++) {
    if (a / 6 != b / 6 && !(m->flags & TSG_FLAG_CROSSBAR_DERIV)) continue;
    double v = 0;
    for (int t = 0; t < TSG_NTEN; t++) v += Bt[t] * d->ten_J[t][a] * d->ten_J[t][b];
    A[a][b] = -h * v;
  }
  for (int k = 0; k < NV; k++) { A[k][k] += Mdiag(m, k); qacc[k] = d->qfrc_smooth[k] + d->qfrc_constraint[k]; }
  /* SPD solve (dense Cholesky; block-diagonal unless CROSSBAR) */
  for (int j = 0; j < NV; j++) {
    double s = A[j][j];
    for (int k = 0; k < j; k++) s -= A[j][k] * A[j][k];
    double piv = sqrt(s);
    A[j][j] = piv;
    for (int i = j + 1; i < NV; i++) { double t = A[i][j]; for (int k = 0; k < j; k++) t -= A[i][k] * A[j][k]; A[i][j] = t / piv; }
  }
  for (int i = 0; i < NV; i++) { double t = qacc[i]; for (int k = 0; k < i; k++) t -= A[i][k] * qacc[k]; qacc[i] = t / A[i][i]; }
  for (int i = NV - 1; i >= 0; i--) { double t = qacc[i]; for (int k = i + 1; k < NV; k++) t -= A[k][i] * qacc[k]; qacc[i] = t / A[i][i]; }
  /* mj_advance */
  if (m->act_dyntype == TSG_DYN_FILTER) for (int i = 0; i < TSG_NACT; i++) d->act[i] += h * d->act_dot[i];
  for (int k = 0; k < NV; k++) d->qvel[k] += h * qacc[k];
  for (int b = 0; b < TSG_NBAR; b++) { /* mj_integratePos, free joint */
    double *q = d->qpos + 7 * b, *v = d->qvel + 6 * b;
    for (int k = 0; k < 3; k++) q[k] += h * v[k];
    double ax[3] = {v[3], v[4], v[5]}, qr[4], qn[4];
    double ang = h * normalize3(ax);
    if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
    else { double s = sin(ang * 0.5); qr[0] = cos(ang * 0.5); qr[1] = ax[0] * s; qr[2] = ax[1] * s; qr[3] = ax[2] * s; }
    normalize4(q + 3);
    mulQuat(qn, q + 3, qr);
    memcpy(q + 3, qn, sizeof(qn));
  }
  d->time += h;
}

/* ------------------------------------------------------------------ mj_step */
static int check_bad(const double *x, int n) { for (int i = 0; i < n; i++) if (is_bad(x[i])) return 1; return 0; }
void tsgo_step(const TsgModel *m, TsgoData *d, int nstep) {
  for (int s = 0; s < nstep; s++) {
    int w = 0;
    if (check_bad(d->qpos, TSG_NQ)) w |= 1;   /* mj_checkPos */
    if (check_bad(d->qvel, TSG_NV)) w |= 2;   /* mj_checkVel */
    if (w) { int keep = d->warning | w; tsgo_reset_data(m, d); d->warning = keep; }
    tsgo_forward(m, d);
    if (check_bad(d->qacc, TSG_NV)) {         /* mj_checkAcc */
      int keep = d->warning | 4; tsgo_reset_data(m, d); d->warning = keep;
      tsgo_forward(m, d);
    }
    implicit_advance(m, d);
  }
}

/* ------------------------------------------------------------------ mj_contactForce / mj_rnePostConstraint */
void tsgo_contact_force(const TsgModel *m, const TsgoData *d, int id, double out[6]) {
  (void)m;
  for (int k = 0; k < 6; k++) out[k] = 0;
  if (id < 0 || id >= d->ncon || d->contact[id].efc_address < 0) return;
  for (int k = 0; k < 6; k++) out[k] = d->efc_force[d->contact[id].efc_address + k];
}
void tsgo_rne_post_constraint(const TsgModel *m, TsgoData *d) {
  memset(d->cfrc_ext, 0, sizeof(d->cfrc_ext));
  for (int i = 0; i < d->ncon; i++) {
    const TsgoContact *c = &d->contact[i];
    if (c->efc_address < 0) continue;
    double lf[6], F[3], T[3];
    tsgo_contact_force(m, d, i, lf);
    mulMTV(F, c->frame, lf); mulMTV(T, c->frame, lf + 3);
    for (int side = 0; side < 2; side++) {
      int body = side ? c->body2 : c->body1;
      const double *com = body ? d->xpos[body - 1] : d->com_world;
      double r[3], tq[3], s = side ? 1.0 : -1.0;
      sub3(r, c->pos, com); cross3(tq, r, F); add3(tq, tq, T);
      for (int k = 0; k < 3; k++) { d->cfrc_ext[body][k] += s * tq[k]; d->cfrc_ext[body][3 + k] += s * F[k]; }
    }
  }
}

/* ------------------------------------------------------------------ CPU baseline driver */
static unsigned long long splitmix(unsigned long long *s) {
  unsigned long long z = (*s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
int tsgo_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}
/* persistent batch of independent envs stepped by a pthread pool (the timed CPU arm of bench.py) */
typedef struct TsgoBatch {
  const TsgModel *m; int n_envs; TsgoData *d; unsigned long long *rng;
  int frame_skip; double lo, hi; int n_steps; int next; pthread_mutex_t mu;
} TsgoBatch;
TsgoBatch *tsgo_batch_create(const TsgModel *m, int n_envs, unsigned long long seed) {
  TsgoBatch *b = (TsgoBatch *)calloc(1, sizeof(TsgoBatch));
  b->m = m; b->n_envs = n_envs;
  b->d = (TsgoData *)malloc(sizeof(TsgoData) * (size_t)n_envs);
  b->rng = (unsigned long long *)malloc(sizeof(unsigned long long) * (size_t)n_envs);
  for (int e = 0; e < n_envs; e++) { tsgo_reset_data(m, &b->d[e]); b->rng[e] = seed * 1000003ULL + (unsigned long long)e; }
  pthread_mutex_init(&b->mu, 0);
  return b;
}
void tsgo_batch_destroy(TsgoBatch *b) { if (b) { free(b->d); free(b->rng); free(b); } }
static void *batch_worker(void *arg) {
  TsgoBatch *b = (TsgoBatch *)arg;
  for (;;) {
    pthread_mutex_lock(&b->mu);
    int e = b->next++;
    pthread_mutex_unlock(&b->mu);
    if (e >= b->n_envs) break;
    TsgoData *d = &b->d[e];
    for (int st = 0; st < b->n_steps; st++) {
      for (int i = 0; i < TSG_NACT; i++) {
        double u = (double)(splitmix(&b->rng[e]) >> 11) * (1.0 / 9007199254740992.0);
        d->ctrl[i] = b->lo + (b->hi - b->lo) * u;
      }
      tsgo_step(b->m, d, b->frame_skip);
      tsgo_rne_post_constraint(b->m, d);
    }
  }
  return 0;
}
/* n_steps env-steps (frame_skip substeps + mj_rnePostConstraint each) for every env, uniform random ctrl */
long tsgo_batch_step(TsgoBatch *b, int n_steps, int frame_skip, double lo, double hi, int n_threads) {
  if (n_threads <= 0) n_threads = tsgo_max_threads();
  if (n_threads > 256) n_threads = 256;
  b->n_steps = n_steps; b->frame_skip = frame_skip; b->lo = lo; b->hi = hi; b->next = 0;
  pthread_t th[256];
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], 0, batch_worker, b);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
  return (long)b->n_envs * n_steps;
}
void tsgo_batch_get(const TsgoBatch *b, int e, double *qpos, double *qvel) {
  memcpy(qpos, b->d[e].qpos, sizeof(double) * TSG_NQ); memcpy(qvel, b->d[e].qvel, sizeof(double) * TSG_NV);
}

typedef struct {
  const TsgModel *m; int n_envs, n_steps, frame_skip, warm_steps; double lo, hi; unsigned long long seed;
  int next; pthread_mutex_t mu; double checksum; long total;
} BenchJob;
static void *bench_worker(void *arg) {
  BenchJob *j = (BenchJob *)arg;
  TsgoData *d = (TsgoData *)malloc(sizeof(TsgoData));
  double checksum = 0; long total = 0;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int e = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (e >= j->n_envs) break;
    unsigned long long s = j->seed * 1000003ULL + (unsigned long long)e;
    tsgo_reset_data(j->m, d);
    for (int st = 0; st < j->warm_steps + j->n_steps; st++) {
      for (int i = 0; i < TSG_NACT; i++) {
        double u = (double)(splitmix(&s) >> 11) * (1.0 / 9007199254740992.0);
        d->ctrl[i] = j->lo + (j->hi - j->lo) * u;
      }
      tsgo_step(j->m, d, j->frame_skip);
      tsgo_rne_post_constraint(j->m, d);
      if (st >= j->warm_steps) total++;
    }
    for (int k = 0; k < TSG_NQ; k++) checksum += d->qpos[k];
  }
  free(d);
  pthread_mutex_lock(&j->mu);
  j->checksum += checksum; j->total += total;
  pthread_mutex_unlock(&j->mu);
  return 0;
}
long tsgo_bench(const TsgModel *m, int n_envs, int n_steps, int frame_skip, int warm_steps,
                double ctrl_lo, double ctrl_hi, unsigned long long seed, int n_threads, double *out_checksum) {
  BenchJob j = {m, n_envs, n_steps, frame_skip, warm_steps, ctrl_lo, ctrl_hi, seed, 0, PTHREAD_MUTEX_INITIALIZER, 0.0, 0};
  if (n_threads <= 0) n_threads = tsgo_max_threads();
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], 0, bench_worker, &j);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
  if (out_checksum) *out_checksum = j.checksum;
  return j.total;
}

/* n independent states stepped `nstep` times each from the given (qpos, qvel, act, warm start, ctrl) by a pthread
 * pool: the checker of the full-batch parity tests.  ncon_minmax[e] = {min, max} active contacts over the substeps. */
typedef struct {
  const TsgModel *m; int n, nstep;
  const double *qpos, *qvel, *act, *warm, *ctrl; double *oq, *ov, *ot; int *mm;
  int next; pthread_mutex_t mu;
} StatesJob;
static void *states_worker(void *arg) {
  StatesJob *j = (StatesJob *)arg;
  TsgoData *d = (TsgoData *)malloc(sizeof(TsgoData));
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int e0 = j->next; j->next += 16;
    pthread_mutex_unlock(&j->mu);
    if (e0 >= j->n) break;
    for (int e = e0; e < e0 + 16 && e < j->n; e++) {
      tsgo_reset_data(j->m, d);
      memcpy(d->qpos, j->qpos + (size_t)e * TSG_NQ, sizeof(double) * TSG_NQ);
      memcpy(d->qvel, j->qvel + (size_t)e * TSG_NV, sizeof(double) * TSG_NV);
      memcpy(d->act, j->act + (size_t)e * TSG_NACT, sizeof(double) * TSG_NACT);
      memcpy(d->qacc_warmstart, j->warm + (size_t)e * TSG_NV, sizeof(double) * TSG_NV);
      memcpy(d->ctrl, j->ctrl + (size_t)e * TSG_NACT, sizeof(double) * TSG_NACT);
      int lo = 1 << 30, hi = -1;
      for (int s = 0; s < j->nstep; s++) {
        tsgo_step(j->m, d, 1);
        int nc = d->nefc / 6;
        if (nc < lo) lo = nc;
        if (nc > hi) hi = nc;
      }
      memcpy(j->oq + (size_t)e * TSG_NQ, d->qpos, sizeof(double) * TSG_NQ);
      memcpy(j->ov + (size_t)e * TSG_NV, d->qvel, sizeof(double) * TSG_NV);
      memcpy(j->ot + (size_t)e * TSG_NTEN, d->ten_length, sizeof(double) * TSG_NTEN);
      if (j->mm) { j->mm[2 * e] = lo; j->mm[2 * e + 1] = hi; }
    }
  }
  free(d);
  return 0;
}
void tsgo_step_states(const TsgModel *m, int n, int nstep, const double *qpos, const double *qvel, const double *act,
                      const double *warm, const double *ctrl, double *out_qpos, double *out_qvel, double *out_ten,
                      int *ncon_minmax, int n_threads) {
  StatesJob j = {m, n, nstep, qpos, qvel, act, warm, ctrl, out_qpos, out_qvel, out_ten, ncon_minmax, 0, PTHREAD_MUTEX_INITIALIZER};
  if (n_threads <= 0) n_threads = tsgo_max_threads();
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  for (int t = 0; t < n_threads; t++) pthread_create(&th[t], 0, states_worker, &j);
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
}
