// tb_math.h -- scalar / 3-vector helpers of the bar-lane kernel, templated on the arithmetic type (double: the
// reference's precision; float: the optional fp32 mode).
#pragma once
#include "tb_simt.h"

namespace tb {

template <typename real> struct Lim;
template <> struct Lim<double> {
  static constexpr double MINVAL = 1e-15, MAXVAL = 1e10, EPS = 2.220446049250313e-16, TINY = 2.2250738585072014e-308;
};
template <> struct Lim<float> {
  static constexpr float MINVAL = 1e-15f, MAXVAL = 1e10f, EPS = 1.1920929e-07f, TINY = 1.17549435e-38f;
};

#if TB_DEV
// fp64 sqrt: MUFU.RSQ64H seed + two coupled Goldschmidt steps + a residual correction (<= 1 ulp from IEEE; 0 for
// x == 0 and subnormal x, which every caller treats as zero).  12 instructions instead of the 38-instruction library
// routine; the kernel executes a few hundred square roots per env per substep (norms, cone radii, MPR portals).
TB_FN double rsqrt_seed(double x) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
TB_FN double tsqrt(double x) {
  double y = rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  g = fma(fma(-g, g, x), h, g);
  return x >= 2.2250738585072014e-308 ? g : (x < 0 ? x * __longlong_as_double(0x7ff8000000000000ll) : 0.0);
}
TB_FN float tsqrt(float x) { return sqrtf(x); }
TB_FN double trcp(double x) { return __drcp_rn(x); }
TB_FN float trcp(float x) { return __frcp_rn(x); }
// a / b: correctly rounded reciprocal + one residual step (<= 1 ulp from IEEE; callers guarantee finite non-zero b)
TB_FN double tdiv(double a, double b) { double r = __drcp_rn(b), q = a * r; return fma(fma(-b, q, a), r, q); }
TB_FN float tdiv(float a, float b) { return __fdividef(a, b); }
#else
TB_FN double tsqrt(double x) { return sqrt(x); }
TB_FN float tsqrt(float x) { return sqrtf(x); }
TB_FN double trcp(double x) { return 1.0 / x; }
TB_FN float trcp(float x) { return 1.0f / x; }
TB_FN double tdiv(double a, double b) { return a / b; }
TB_FN float tdiv(float a, float b) { return a / b; }
#endif
TB_FN double tabs(double x) { return fabs(x); }
TB_FN float tabs(float x) { return fabsf(x); }
TB_FN double tmin(double a, double b) { return fmin(a, b); }
TB_FN float tmin(float a, float b) { return fminf(a, b); }
TB_FN double tmax(double a, double b) { return fmax(a, b); }
TB_FN float tmax(float a, float b) { return fmaxf(a, b); }
TB_FN double tfloor(double a) { return floor(a); }
TB_FN float tfloor(float a) { return floorf(a); }
TB_FN double tceil(double a) { return ceil(a); }
TB_FN float tceil(float a) { return ceilf(a); }
TB_FN void tsincos(double a, double* s, double* c) { sincos(a, s, c); }
TB_FN void tsincos(float a, float* s, float* c) { sincosf(a, s, c); }
template <typename real> TB_FN real clampr(real x, real lo, real hi) { return tmin(hi, tmax(lo, x)); }
template <typename real> TB_FN bool is_bad(real x) { return !(x <= Lim<real>::MAXVAL && x >= -Lim<real>::MAXVAL); }

template <typename real> TB_FN real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename real> TB_FN void cross3(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename real> TB_FN void sub3(real* r, const real* a, const real* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
template <typename real> TB_FN void add3(real* r, const real* a, const real* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
template <typename real> TB_FN void copy3(real* r, const real* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
template <typename real> TB_FN void scl3(real* r, const real* a, real s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
template <typename real> TB_FN void addscl3(real* r, const real* a, real s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
template <typename real> TB_FN real normalize3(real* a) {
  real n = tsqrt(dot3(a, a));
  if (n < Lim<real>::MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { real s = trcp(n); a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
// r = R v, r = R^T v for a row-major 3x3
template <typename real> TB_FN void mulMV(real* r, const real* R, const real* v) {
  real x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  real y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  real z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename real> TB_FN void mulMTV(real* r, const real* R, const real* v) {
  real x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  real y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  real z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename real> TB_FN void quat2mat(real* R, const real* q) {
  real q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  real q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  real q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2 * (q12 - q03); R[2] = 2 * (q13 + q02);
  R[3] = 2 * (q12 + q03); R[5] = 2 * (q23 - q01);
  R[6] = 2 * (q13 - q02); R[7] = 2 * (q23 + q01);
}
template <typename real> TB_FN void normalize4(real* q) {
  real n = tsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < Lim<real>::MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (tabs(n - 1) > Lim<real>::MINVAL) { real s = trcp(n); q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}

}  // namespace tb
