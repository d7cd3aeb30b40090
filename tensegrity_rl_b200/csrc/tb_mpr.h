// tb_mpr.h -- convex narrow phase with libccd's MPR semantics (what MuJoCo 2.3.7 runs for cylinder-cylinder,
// sphere-cylinder and height-field-prism pairs; libccd is a third-party dependency of MuJoCo, absent from the
// reference tree, restated here from its published algorithm).  One pair per lane, run to completion by that lane.
#pragma once
#include "tb_math.h"
#include "tb_model.h"

namespace tb {

// type GEOM_SPHERE / GEOM_CYL: centre pos, unit axis (the bar axis in the world frame), size = (radius, half length).
// type 100 = height-field prism: its three (x, y) columns with the top heights and the common base height; vertex
// i < 3 is (px[i], py[i], pbase), vertex 3 + i is (px[i], py[i], pz[i]).
template <typename real>
struct CObj { int type; real pos[3]; real axis[3]; real size[2]; real px[3], py[3], pz[3], pbase; };
template <typename real> struct Supp { real v[3], v1[3]; };  // v = v1 - v2 ; v2 recovered as v1 - v

template <typename real> TB_FN bool ccd_is_zero(real x) { return tabs(x) < Lim<real>::EPS; }
template <typename real> TB_FN bool ccd_eq(real a_, real b_) {
  real ab = tabs(a_ - b_);
  if (ab < Lim<real>::EPS) return true;
  real a = tabs(a_), b = tabs(b_);
  return (b > a) ? (ab < Lim<real>::EPS * b) : (ab < Lim<real>::EPS * a);
}
template <typename real> TB_FN bool ccd_vec_is_origin(const real* a) { return ccd_eq(a[0], real(0)) && ccd_eq(a[1], real(0)) && ccd_eq(a[2], real(0)); }
template <typename real> TB_FN void ccd_normalize(real* v) { real s = trcp(tsqrt(dot3(v, v))); v[0] *= s; v[1] *= s; v[2] *= s; }

template <typename real> TB_FN void obj_support(const CObj<real>& o, const real* dir, real* out) {
  if (o.type == 100) {
    // first maximum of vertex . dir over the vertices in order (bottom 0..2, top 3..5), as libccd's loop finds it; the
    // winner's coordinates are carried along (no dynamic index: the object stays in registers)
    real t[3], bd = 0, bx = 0, by = 0, bz = 0;
    TB_UNROLL
    for (int j = 0; j < 3; j++) t[j] = o.px[j] * dir[0] + o.py[j] * dir[1];
    TB_UNROLL
    for (int i = 0; i < 6; i++) {
      const real vz = i < 3 ? o.pbase : o.pz[i % 3];
      const real dd = t[i % 3] + vz * dir[2];
      if (i == 0 || dd > bd) { bd = dd; bx = o.px[i % 3]; by = o.py[i % 3]; bz = vz; }
    }
    out[0] = bx; out[1] = by; out[2] = bz;
    return;
  }
  // spheres and cylinders are bodies of revolution about the bar axis a (third column of R): the support point is
  // pos + r perp / |perp| + sign(dir . a) h a with perp = dir - (dir . a) a  (sphere: pos + r dir) -- the same point as
  // R * support_local(R^T dir), without the two rotations
  if (o.type == GEOM_SPHERE) { for (int k = 0; k < 3; k++) out[k] = o.pos[k] + o.size[0] * dir[k]; return; }
  const real a[3] = {o.axis[0], o.axis[1], o.axis[2]};
  const real da = dot3(dir, a);
  real perp[3] = {dir[0] - da * a[0], dir[1] - da * a[1], dir[2] - da * a[2]};
  const real tmp = tsqrt(dot3(perp, perp));
  const real s = tmp > Lim<real>::MINVAL ? tdiv(o.size[0], tmp) : real(0);
  const real hz = (da > 0 ? real(1) : (da < 0 ? real(-1) : real(0))) * o.size[1];
  for (int k = 0; k < 3; k++) out[k] = o.pos[k] + s * perp[k] + hz * a[k];
}
template <typename real> TB_FN void obj_center(const CObj<real>& o, real* c) {
  if (o.type == 100) {
    c[0] = c[1] = c[2] = 0;
    TB_UNROLL
    for (int i = 0; i < 6; i++) { c[0] += o.px[i % 3]; c[1] += o.py[i % 3]; c[2] += i < 3 ? o.pbase : o.pz[i % 3]; }
    c[0] /= 6; c[1] /= 6; c[2] /= 6;
  } else copy3(c, o.pos);
}
template <typename real> TB_FN void mink_support(const CObj<real>& o1, const CObj<real>& o2, const real* dir, Supp<real>& s) {
  real nd[3] = {-dir[0], -dir[1], -dir[2]}, v2[3];
  obj_support(o1, dir, s.v1);
  obj_support(o2, nd, v2);
  sub3(s.v, s.v1, v2);
}
template <typename real> TB_FN void portal_dir(const Supp<real>& p1, const Supp<real>& p2, const Supp<real>& p3, real* dir) {
  real a[3], b[3];
  sub3(a, p2.v, p1.v); sub3(b, p3.v, p1.v);
  cross3(dir, a, b); ccd_normalize(dir);
}
template <typename real>
TB_FN bool portal_reach_tol(const Supp<real>& p1, const Supp<real>& p2, const Supp<real>& p3, const Supp<real>& v4, const real* dir, real tol) {
  real dv4 = dot3(v4.v, dir);
  real d1 = dv4 - dot3(p1.v, dir), d2 = dv4 - dot3(p2.v, dir), d3 = dv4 - dot3(p3.v, dir);
  d1 = tmin(d1, d2); d1 = tmin(d1, d3);
  return ccd_eq(d1, tol) || d1 < tol;
}
template <typename real>
TB_FN void expand_portal(const Supp<real>& p0, Supp<real>& p1, Supp<real>& p2, Supp<real>& p3, const Supp<real>& v4) {
  real v4v0[3];
  cross3(v4v0, v4.v, p0.v);
  if (dot3(p1.v, v4v0) > 0) { if (dot3(p2.v, v4v0) > 0) p1 = v4; else p3 = v4; }
  else { if (dot3(p3.v, v4v0) > 0) p2 = v4; else p1 = v4; }
}
template <typename real> TB_FN real seg_dist2_origin(const real* x0, const real* b, real* wit) {
  real d[3];
  sub3(d, b, x0);
  real t = real(-1) * dot3(x0, d); t /= dot3(d, d);
  if (t < 0 || ccd_is_zero(t)) copy3(wit, x0);
  else if (t > 1 || ccd_eq(t, real(1))) copy3(wit, b);
  else { scl3(wit, d, t); add3(wit, wit, x0); }
  return dot3(wit, wit);
}
template <typename real> TB_FN real tri_dist2_origin(const real* x0, const real* B, const real* C, real* wit) {
  real d1[3], d2[3];
  sub3(d1, B, x0); sub3(d2, C, x0);
  real v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  real s, t, dist, dd = w * v - r * r;
  if (ccd_is_zero(dd)) s = t = -1;
  else { s = (q * r - w * p) / dd; t = (-s * r - q) / w; }
  if ((ccd_is_zero(s) || s > 0) && (ccd_eq(s, real(1)) || s < 1) && (ccd_is_zero(t) || t > 0) &&
      (ccd_eq(t, real(1)) || t < 1) && (ccd_eq(t + s, real(1)) || t + s < 1)) {
    scl3(d1, d1, s); scl3(d2, d2, t);
    copy3(wit, x0); add3(wit, wit, d1); add3(wit, wit, d2);
    dist = dot3(wit, wit);
  } else {
    real w2[3], dist2;
    dist = seg_dist2_origin(x0, B, wit);
    dist2 = seg_dist2_origin(x0, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
    dist2 = seg_dist2_origin(B, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
  }
  return dist;
}

// MPR for the lanes of a warp that have a pair (`has`), in lock step: the algorithm is run as a per-lane state
// machine in which every tick is ONE Minkowski support query plus the bookkeeping of the phase the lane is in
// (portal discovery / refinement / penetration search), so lanes whose searches take different paths still share the
// instruction stream, and the portal lives in registers.  Warp-collective (every lane of the warp calls it).
// Returns true on penetration; dir points from obj1 to obj2.
template <typename real>
TB_NOINL bool mpr_penetration(const CObj<real>& o1, const CObj<real>& o2, real tol, int max_iter, bool has,
                              real* depth, real* dir_out, real* pos_out) {
  enum { S1 = 0, S2, S3, S4, S5, DONE };
  Supp<real> p0, p1, p2, p3, v4;
  real dir[3] = {1, 0, 0}, c2[3] = {0, 0, 0};
  int st = DONE, it = 0;
  bool hit = false;
  p0 = Supp<real>(); p1 = p0; p2 = p0; p3 = p0; v4 = p0;
  if (has) {
    obj_center(o1, p0.v1); obj_center(o2, c2);
    sub3(p0.v, p0.v1, c2);
    if (ccd_vec_is_origin(p0.v)) p0.v[0] += Lim<real>::EPS * real(10);
    scl3(dir, p0.v, real(-1)); ccd_normalize(dir);
    st = S1;
  }
  // after the portal is closed: refinement needs a support query only while the origin ray misses the portal
  auto refine_check = [&]() {
    portal_dir(p1, p2, p3, dir);
    real dot = dot3(dir, p1.v);
    if (ccd_is_zero(dot) || dot > 0) { it = 0; st = S5; }   // findPenetr starts with the same portal direction
    else st = S4;
  };
  TB_UNROLL1
  while (any(st != DONE)) {
    Supp<real> s;
    s = Supp<real>();
    if (st != DONE) mink_support(o1, o2, dir, s);
    if (st == S1) {
      p1 = s;
      real dot = dot3(p1.v, dir);
      if (ccd_is_zero(dot) || dot < 0) st = DONE;
      else {
        cross3(dir, p0.v, p1.v);
        if (ccd_is_zero(dot3(dir, dir))) {
          // origin on v1 (touching: depth 0, no direction -> MuJoCo drops it) or on the v0-v1 segment
          if (!ccd_vec_is_origin(p1.v)) {
            real v2[3];
            sub3(v2, p1.v1, p1.v);
            add3(pos_out, p1.v1, v2); scl3(pos_out, pos_out, real(0.5));
            copy3(dir_out, p1.v); *depth = tsqrt(dot3(dir_out, dir_out)); ccd_normalize(dir_out);
            hit = true;
          }
          st = DONE;
        } else { ccd_normalize(dir); st = S2; }
      }
    } else if (st == S2) {
      p2 = s;
      real dot = dot3(p2.v, dir);
      if (ccd_is_zero(dot) || dot < 0) st = DONE;
      else {
        real va[3], vb[3];
        sub3(va, p1.v, p0.v); sub3(vb, p2.v, p0.v);
        cross3(dir, va, vb); ccd_normalize(dir);
        if (dot3(dir, p0.v) > 0) { Supp<real> t = p1; p1 = p2; p2 = t; scl3(dir, dir, real(-1)); }
        st = S3;
      }
    } else if (st == S3) {
      p3 = s;
      real dot = dot3(p3.v, dir);
      if (ccd_is_zero(dot) || dot < 0) st = DONE;
      else {
        bool cont = false;
        real va[3], vb[3];
        cross3(va, p1.v, p3.v); dot = dot3(va, p0.v);
        if (dot < 0 && !ccd_is_zero(dot)) { p2 = p3; cont = true; }
        if (!cont) {
          cross3(va, p3.v, p2.v); dot = dot3(va, p0.v);
          if (dot < 0 && !ccd_is_zero(dot)) { p1 = p3; cont = true; }
        }
        if (cont) {
          sub3(va, p1.v, p0.v); sub3(vb, p2.v, p0.v);
          cross3(dir, va, vb); ccd_normalize(dir);
        } else refine_check();
      }
    } else if (st == S4) {
      v4 = s;
      real dot = dot3(v4.v, dir);
      if (!(ccd_is_zero(dot) || dot > 0) || portal_reach_tol(p1, p2, p3, v4, dir, tol)) st = DONE;
      else { expand_portal(p0, p1, p2, p3, v4); refine_check(); }
    } else if (st == S5) {
      v4 = s;
      if (portal_reach_tol(p1, p2, p3, v4, dir, tol) || it > max_iter) {
        real pdir[3];
        *depth = tsqrt(tri_dist2_origin(p1.v, p2.v, p3.v, pdir));
        if (ccd_is_zero(pdir[0]) && ccd_is_zero(pdir[1]) && ccd_is_zero(pdir[2])) copy3(pdir, dir);
        ccd_normalize(pdir);
        copy3(dir_out, pdir);
        // findPos: barycentric blend of the witness points
        real vec[3], bb[4], sum;
        portal_dir(p1, p2, p3, dir);
        cross3(vec, p1.v, p2.v); bb[0] = dot3(vec, p3.v);
        cross3(vec, p3.v, p2.v); bb[1] = dot3(vec, p0.v);
        cross3(vec, p0.v, p1.v); bb[2] = dot3(vec, p3.v);
        cross3(vec, p2.v, p1.v); bb[3] = dot3(vec, p0.v);
        sum = bb[0] + bb[1] + bb[2] + bb[3];
        if (ccd_is_zero(sum) || sum < 0) {
          bb[0] = 0;
          cross3(vec, p2.v, p3.v); bb[1] = dot3(vec, dir);
          cross3(vec, p3.v, p1.v); bb[2] = dot3(vec, dir);
          cross3(vec, p1.v, p2.v); bb[3] = dot3(vec, dir);
          sum = bb[1] + bb[2] + bb[3];
        }
        real inv = trcp(sum), q1[3], q2[3], w2[3];
        // v0's second witness is obj2's centre
        scl3(q1, p0.v1, bb[0]); scl3(q2, c2, bb[0]);
        addscl3(q1, p1.v1, bb[1]); sub3(w2, p1.v1, p1.v); addscl3(q2, w2, bb[1]);
        addscl3(q1, p2.v1, bb[2]); sub3(w2, p2.v1, p2.v); addscl3(q2, w2, bb[2]);
        addscl3(q1, p3.v1, bb[3]); sub3(w2, p3.v1, p3.v); addscl3(q2, w2, bb[3]);
        scl3(q1, q1, inv); scl3(q2, q2, inv);
        add3(pos_out, q1, q2); scl3(pos_out, pos_out, real(0.5));
        hit = true;
        st = DONE;
      } else {
        expand_portal(p0, p1, p2, p3, v4);
        it++;
        portal_dir(p1, p2, p3, dir);
      }
    }
  }
  return hit;
}

}  // namespace tb
