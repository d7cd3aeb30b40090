// tsg_core.cuh -- one-warp-per-env tensegrity step for sm_100a.
//
// Each warp owns ONE environment for a whole env step: the 18-DoF state, tendon
// Jacobians, contact rows and the 18x18 Newton Hessian live in that warp's slice
// of shared memory for all frame_skip substeps; HBM is touched only at the step
// boundary (one contiguous state record per env, ctrl in, obs/reward/done out).
// Work inside a substep is spread over the 32 lanes as "items" (tendons, dofs,
// contact rows, Hessian entries, collision candidates) separated by __syncwarp().
//
// What is computed is MuJoCo 2.3.7's mj_step for this model:
// kinematics -> tendons -> collision (plane / height field / bar-bar with MPR)
// -> elliptic condim-6 contact rows -> smooth forces -> Newton with exact line
// search -> implicitfast -> advance; then the tr_env / tensegrity_env epilogue
// (tr_env.py:327-527, tensegrity_env.py:291-430).
//
// The same source compiles as plain C++ with TSG_HOST_EMUL (a serial "one warp"
// emulator used ONLY by tests/emul to debug the lane logic without a GPU; the
// product library never builds or calls it).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__) && !defined(TSG_HOST_EMUL)
#define TSG_DEVICE 1
#define TSG_FN __device__ __forceinline__
#define TSG_FN_NOINLINE __device__ __noinline__
#define TSG_UNROLL1 _Pragma("unroll 1")
// Virtual warp: TSG_VW lanes (32, 16 or 8) cooperate on one env, so 32 / TSG_VW envs share every instruction a
// physical warp fetches while their control flow coincides (the kernel is instruction-fetch bound); where it does
// not, SIMT divergence serialises them.  `lane` is the lane inside the virtual warp everywhere below.
#ifndef TSG_VW
#define TSG_VW 32
#endif
// default alignment scope: the CTA (see align_any below); -DTSG_NO_ALIGN builds the free-running variant
#if !defined(TSG_ALIGN) && !defined(TSG_ALIGN_WARP) && !defined(TSG_NO_ALIGN)
#define TSG_ALIGN 1
#endif
#define TSG_VMASK() ((TSG_VW == 32) ? 0xffffffffu : ((((1u << (TSG_VW & 31)) - 1u)) << ((threadIdx.x & 31) / TSG_VW * TSG_VW)))
#define LANE_FOR(i, n) for (int i = lane; i < (n); i += TSG_VW)
#define LANE_FOR_ALL(i, n) for (int i##_b = 0, i = lane; i##_b < (n); i##_b += TSG_VW, i += TSG_VW)
#define WSYNC() __syncwarp(TSG_VMASK())
#else
#define TSG_DEVICE 0
#define TSG_FN static inline
#define TSG_FN_NOINLINE static
#define TSG_UNROLL1
#ifdef TSG_EMUL_REVERSE  // second emulator build: items of a phase in reverse order, to expose intra-phase hazards
#define LANE_FOR(i, n) for (int i = (n) - 1; i >= 0; --i)
#else
#define LANE_FOR(i, n) for (int i = 0; i < (n); ++i)
#endif
#define LANE_FOR_ALL(i, n) for (int i = 0; i < (n); ++i)
#define WSYNC() ((void)0)
#endif

namespace tsg {

constexpr int NBAR = 3, NGEOM = 15, NTEN = 9, NEND = 18, NACT = 6, NQ = 21, NV = 18;
#ifndef TSG_MAXC_S
#define TSG_MAXC_S 3  // 3 slots in 6288 B of scratch per env: the 3 resident CTAs stay within the 164 KB shared-memory
                      // configuration (92 KB of L1 left for the stack and the spill area; 196 KB costs 5 %)
#endif
constexpr int MAXC_S = TSG_MAXC_S;       // contact slots per env in shared memory
constexpr int MAXC = 32;        // total contact slots per env (slots >= MAXC_S spill to a per-warp global area)
constexpr int MAXCAND = 64;     // narrow-phase candidates per collision pass
constexpr int NTRI = NV * (NV + 1) / 2;  // packed lower triangle of the Hessian
constexpr double MINVAL = 1e-15, MAXVAL = 1e10, MINIMP = 0.0001, MAXIMP = 0.9999;
constexpr double CCD_EPS = 2.220446049250313e-16;
constexpr int GEOM_SPHERE = 2, GEOM_CYL = 5;
constexpr int ZONE_TOP = 0, ZONE_BOTTOM = 1, ZONE_MIDDLE = 2;
constexpr int STATE_STRIDE = 96;  // doubles per env record in HBM (768 B, 128 B aligned)
constexpr int INFO_DIM = 32;
constexpr int HEADING_SLOTS = 32;
constexpr int NDRAW = 10;

// offsets into the per-env state record
enum StateOff {
  SO_QPOS = 0, SO_QVEL = 21, SO_WARM = 39, SO_CTRL = 57, SO_ACT = 63,
  SO_XY_PREV = 69, SO_PSI_PREV = 71, SO_RESET_PSI = 72, SO_WAYPT = 73, SO_ORI = 75,
  SO_STEP_NUM = 77, SO_EP_RET = 78, SO_EP_LEN = 79, SO_XVEL = 80, SO_YVEL = 81,
  SO_HEAD_N = 82, SO_HEAD_POS = 83, SO_FLAGS = 84, SO_NRESET = 85, SO_USED = 86
};
// info row (per env, per step)
enum InfoOff {
  IO_REW_FWD = 0, IO_REW_CTRL, IO_REW_SURVIVE, IO_X, IO_Y, IO_PSI, IO_XVEL, IO_YVEL,
  IO_TEN = 8 /* 9 */, IO_TERMINATED = 17, IO_TRUNCATED, IO_NCON, IO_NITER, IO_NLS, IO_BARFORCE, IO_MAXCFRC,
  IO_WAYPT = 24 /* 2 */, IO_ORI = 26 /* 2 */, IO_OVERFLOW = 28, IO_BAD = 29, IO_NMPR = 30, IO_RESET_PSI = 31
};

struct DevModel {
  double h, grav[3], tol, ls_tol, mpr_tol, meaninertia;
  int iterations, ls_iterations, mpr_iterations;
  unsigned flags;
  double M[NV], invM[NV];
  double inertia[NBAR][3];
  double invw_tran[NBAR];
  int gtype[NGEOM];
  int pad0;
  double gsize[NGEOM][2];
  double gpos[NGEOM][3];
  // tendons
  int tbody[NEND];
  double tsite[NEND][3];
  double tk[NTEN], tdamp[NTEN], tls[NTEN][2];
  int ten_act[NTEN];
  int nends[NBAR];
  int ends[NBAR][8];
  int act_tendon[NACT];
  int dyntype, ctrllimited, forcelimited, pad1;
  double dynprm0, gain, bias[3], ctrlrange[2], forcerange[2];
  // contact
  double K, B, solimp[5], mu, fr[5], dscale[6], fscale[6];
  double wtab[2][6];
  double solscale;    // 1 / (meaninertia * nv): scale of the solver convergence tests
  double inv_mu2;     // 1 / (mu^2 (1 + mu^2))  // Hessian row weights per zone: bottom = dscale, middle = (0, fr^2)
  // floor
  int floor_type, nrow, ncol, pad2;
  double fpos[3], fnormal[3], hsize[4];
  double hdx, hdy;   // height-field cell size: 2 * hsize[0] / (ncol - 1), 2 * hsize[1] / (nrow - 1)
  const float* hdata;
  double qpos0[NQ];
  unsigned char tri_i[NTRI + 5], tri_j[NTRI + 5];  // packed-triangle index -> (row, col)
};

struct EnvCfg {
  int env_kind, task, frame_skip, obs_dim, use_cap_velocity, terminate_when_unhealthy, is_test;
  int reward_delay_steps, max_episode_steps, warmup_steps, npose, use_obs_noise;
  double desired_direction, ctrl_cost_weight, healthy_reward, yaw_reward_weight;
  double min_reset_heading, max_reset_heading;
  double tendon_reset_mean, tendon_reset_stdev, tendon_min_length, tendon_max_length;
  double waypt_range[2], waypt_angle_range[2];
  double ditch_reward_max, ditch_reward_stdev, waypt_reward_amplitude, waypt_reward_stdev, kill_force, dt;
  double obs_noise_tendon_stdev, obs_noise_cap_pos_stdev;
  double reset_pose[6][NQ];
};

struct Con {
  // Jacobian of the 6 contact rows, per side (side 0 = body b1, side 1 = body b2), sign applied.
  // translational row a: linear part = sgn * frame[a], angular part = Jt[side][a]; rotational row 3+a: angular Jr[side][a]
  double Jt[2][3][3];
  double Jr[2][3][3];
  double frame[9];
  double pos[3];
  double aref[6];  // after the initial residual is built this holds jv = J * search
  double jar[6], force[6];
  double bvec[2][6];
  double su[6];
  double dist, D0, wcoef, ca, cb;
  double U0, V0, UU, UV, VV, q0, q1, q2;
  int b1, b2;  // bar index 0..2, or -1 for the world
  int zone, active;
};

struct Scratch {
  // env state carried across substeps
  double qpos[NQ], qvel[NV], warm[NV], ctrl[NACT], act[NACT];
  // position stage (kept for the epilogue: stale body positions, end-cap centres)
  double xstale[9];
  double xmat[27];
  double sph[18];
  double tlen[NTEN];
  // velocity / force stage (the tendon directions / Jacobian arms / forces live in the union below: u.ten)
  double actdot[NACT];
  double fsm[NV], asmooth[NV], fcon[NV];
  // solver vectors
  double qacc[NV], grad[NV], search[NV], rhs[NV], dinv[NV];
  double qG[3], lsr[3], gauss, cost;
  // time-multiplexed region: collision temporaries -> Hessian -> line-search sums -> post-solve data
  union {
    struct { int cand[MAXCAND]; int hf_cell[NGEOM][4]; double hf_zmin[NGEOM]; } col;
    struct { double H[NTRI]; } hes;
    struct { double acc[MAXC][3]; } ls;
    struct { double Dblk[NBAR][21]; double cfrc[4][6]; double obs[64]; } post;
    // tendon directions, Jacobian arms, forces, damping coefficients: alive from stage_tendon to stage_smooth and,
    // recomputed after the solve, during stage_damping_blocks -- next to post.Dblk, over cfrc / obs (dead inside a substep)
    struct { double pad_[NBAR * 21]; double tdir[NTEN * 3], tJw[NEND * 3], tfrc[NTEN], tB[NTEN]; } ten;
  } u;
  Con con[MAXC_S];
  Con* spill;  // global memory, MAXC - MAXC_S slots owned by this warp
  unsigned char order[MAXC];   // active slot list (MAXC <= 255)
  int nact, nslot, overflow, bad;
  int niter_total, nls_total, nmpr_total, ls_evals;
  int align, pad2;
};

struct EnvScratch : Scratch {
  double action[NACT];
  double draws[NDRAW + 2];
};

// ---- shared-memory context.  Heavy stages are compiled ONCE as __noinline__ functions; on the device
// they re-derive their context from the dynamic shared-memory base so that the compiler keeps LDS/STS
// addressing (a reference passed through a call would degrade to generic loads).
constexpr size_t align16(size_t x) { return (x + 15) & ~size_t(15); }
constexpr size_t SMEM_MODEL = align16(sizeof(DevModel));
constexpr size_t SMEM_CFG = align16(sizeof(EnvCfg));
constexpr size_t SMEM_SCRATCH = align16(sizeof(EnvScratch));
#if TSG_DEVICE
#define CTX_PARAMS int lane
#define CTX_ARGS lane
#define CTX_BIND                                                                                         \
  extern __shared__ __align__(16) unsigned char tsg_smem[];                                              \
  EnvScratch& S = *reinterpret_cast<EnvScratch*>(tsg_smem + SMEM_MODEL + SMEM_CFG + (threadIdx.x / TSG_VW) * SMEM_SCRATCH); \
  const DevModel& m = *reinterpret_cast<const DevModel*>(tsg_smem);                                      \
  const EnvCfg& c = *reinterpret_cast<const EnvCfg*>(tsg_smem + SMEM_MODEL);                             \
  (void)c; (void)m;
#else
#define CTX_PARAMS EnvScratch &S, const DevModel &m, const EnvCfg &c, int lane
#define CTX_ARGS S, m, c, lane
#define CTX_BIND (void)c;
#endif

TSG_FN Con& con_at(Scratch& S, int slot) { return slot < MAXC_S ? S.con[slot] : S.spill[slot - MAXC_S]; }
TSG_FN const Con& con_at(const Scratch& S, int slot) { return slot < MAXC_S ? S.con[slot] : S.spill[slot - MAXC_S]; }
// J element (row r, dof k of the side's bar) and row . 6-vector products on the compressed storage
TSG_FN double Jel(const Con& c, int side, int r, int k) {
  if (r < 3) return k < 3 ? (side ? c.frame[3 * r + k] : -c.frame[3 * r + k]) : c.Jt[side][r][k - 3];
  return k < 3 ? 0.0 : c.Jr[side][r - 3][k - 3];
}
TSG_FN double Jrow_dot(const Con& c, int side, int r, const double* v6) {
  if (r < 3) {
    const double* f = c.frame + 3 * r; const double* a = c.Jt[side][r];
    double lin = f[0] * v6[0] + f[1] * v6[1] + f[2] * v6[2];
    return (side ? lin : -lin) + a[0] * v6[3] + a[1] * v6[4] + a[2] * v6[5];
  }
  const double* a = c.Jr[side][r - 3];
  return a[0] * v6[3] + a[1] * v6[4] + a[2] * v6[5];
}
TSG_FN double Jrow_dot_all(const Con& c, int r, const double* v18) {
  double v = Jrow_dot(c, 1, r, v18 + 6 * c.b2);
  if (c.b1 >= 0) v += Jrow_dot(c, 0, r, v18 + 6 * c.b1);
  return v;
}

// ------------------------------------------------------------------ small math
// The kernel is instruction-fetch bound (ncu: stall_no_instruction dominates), so the long IEEE fp64 sqrt /
// divide sequences exist ONCE as out-of-line functions instead of being expanded at every use.
TSG_FN_NOINLINE double tsg_div(double a, double b) { return a / b; }   // IEEE; cold call sites (collision geometry)
#if TSG_DEVICE
TSG_FN double tsg_rcp(double x) { return __drcp_rn(x); }  // correctly rounded, == 1.0 / x
// sqrt for the hot call sites: MUFU.RSQ64H seed + two coupled Goldschmidt steps + a final residual correction
// (<= 1 ulp from the IEEE result; 0 for x == 0 and for subnormal x, which every caller treats as zero anyway).
// The library sqrt() is a 38-instruction routine whose out-of-line copy also spills callee-saved registers to local
// memory (ncu: 11 % of the kernel's long-scoreboard stalls); this one is 12 instructions in 6 registers.
TSG_FN double tsg_sqrt_inl(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  r = fma(-h, g, 0.5);
  g = fma(g, r, g); h = fma(h, r, h);
  g = fma(fma(-g, g, x), h, g);
  return x >= 2.2250738585072014e-308 ? g : (x < 0 ? x * __longlong_as_double(0x7ff8000000000000ll) : 0.0);
}
TSG_FN_NOINLINE double tsg_sqrt(double x) { return tsg_sqrt_inl(x); }   // out-of-line copy for the colder call sites
// a / b for the hot call sites (line-search steps, impedance): correctly rounded reciprocal + one residual step
// (<= 1 ulp from IEEE; callers guarantee a finite non-zero b)
TSG_FN_NOINLINE double tsg_fdiv(double a, double b) {
  double r = __drcp_rn(b), q = a * r;
  return fma(fma(-b, q, a), r, q);
}
#else
TSG_FN double tsg_rcp(double x) { return 1.0 / x; }
TSG_FN double tsg_sqrt(double x) { return sqrt(x); }
TSG_FN double tsg_sqrt_inl(double x) { return sqrt(x); }
TSG_FN double tsg_fdiv(double a, double b) { return a / b; }
#endif
TSG_FN_NOINLINE double tsg_inv(double x) { return tsg_rcp(x); }  // out-of-line 1 / x for the scattered call sites
// Sum / any over the lanes of per-lane partial results.  Device: xor butterfly, so every lane ends with the same
// bits (control flow branches on these values).  Host emulator: the item loop already ran over all items.
#if TSG_DEVICE
TSG_FN_NOINLINE double warp_sum(double v) {
  const unsigned vm = TSG_VMASK();
#pragma unroll
  for (int o = TSG_VW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(vm, v, o, TSG_VW);
  return v;
}
TSG_FN void warp_sum4(double& a, double& b, double& c, double& d) {   // one call site (line_search): inline, the arguments stay in registers
  const unsigned vm = TSG_VMASK();
#pragma unroll
  for (int o = TSG_VW / 2; o > 0; o >>= 1) {
    a += __shfl_xor_sync(vm, a, o, TSG_VW); b += __shfl_xor_sync(vm, b, o, TSG_VW);
    c += __shfl_xor_sync(vm, c, o, TSG_VW); d += __shfl_xor_sync(vm, d, o, TSG_VW);
  }
}
TSG_FN bool warp_any(bool p) { return __any_sync(TSG_VMASK(), p) != 0; }
#else
TSG_FN double warp_sum(double v) { return v; }
TSG_FN void warp_sum4(double&, double&, double&, double&) {}
TSG_FN bool warp_any(bool p) { return p; }
#endif
// reductions over dofs: lane-parallel partial sums + butterfly (default) or the same loop run by every lane
#ifdef TSG_NO_WSUM
#define RED_FOR(i, n) TSG_UNROLL1 for (int i = 0; i < (n); ++i)
#define RED_SUM(v) (v)
#define RED_SUM4(a, b, c, d) ((void)0)
#define RED_ANY(p) (p)
#else
#define RED_FOR(i, n) LANE_FOR(i, n)
#define RED_SUM(v) warp_sum(v)
#define RED_SUM4(a, b, c, d) warp_sum4(a, b, c, d)
#define RED_ANY(p) warp_any(p)
#endif
TSG_FN double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
TSG_FN void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
TSG_FN void sub3(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
TSG_FN void add3(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
TSG_FN void copy3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
TSG_FN void scl3(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
TSG_FN void addscl3(double* r, const double* a, double s) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
TSG_FN double normalize3(double* a) {
  double n = tsg_sqrt(dot3(a, a));
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; }
  else { double s = tsg_inv(n); a[0] *= s; a[1] *= s; a[2] *= s; }
  return n;
}
TSG_FN void mulMV(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  double y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  double z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
TSG_FN void mulMTV(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
  double y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
  double z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
TSG_FN void quat2mat(double* R, const double* q) {
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2 * (q12 - q03); R[2] = 2 * (q13 + q02);
  R[3] = 2 * (q12 + q03); R[5] = 2 * (q23 - q01);
  R[6] = 2 * (q13 - q02); R[7] = 2 * (q23 + q01);
}
TSG_FN void normalize4(double* q) {
  double n = tsg_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > MINVAL) { double s = tsg_inv(n); q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
TSG_FN bool is_bad(double x) { return !(x <= MAXVAL && x >= -MAXVAL); }
TSG_FN double clampd(double x, double lo, double hi) { return fmin(hi, fmax(lo, x)); }

// exclusive prefix sum of `v` over the items of one LANE_FOR_ALL pass; `total` is a warp-uniform
// running count.  Device: shuffle scan.  Host emulator: items arrive serially in order.
TSG_FN int scan_slot(int v, int& total, int lane) {
#if TSG_DEVICE
  int incl = v;
  const unsigned vm = TSG_VMASK();
#pragma unroll
  for (int o = 1; o < TSG_VW; o <<= 1) {
    int t = __shfl_up_sync(vm, incl, o, TSG_VW);
    if (lane >= o) incl += t;
  }
  int slot = total + incl - v;
  total += __shfl_sync(vm, incl, TSG_VW - 1, TSG_VW);
  return slot;
#else
  (void)lane;
  int slot = total;
  total += v;
  return slot;
#endif
}

// ------------------------------------------------------------------ kinematics + tendons
TSG_FN void geom_center(const Scratch& S, const DevModel& m, int g, double* out) {
  int b = g / 5;
  double v[3];
  mulMV(v, S.xmat + 9 * b, m.gpos[g]);
  add3(out, S.qpos + 7 * b, v);
}
TSG_FN void stage_position(Scratch& S, const DevModel& m, int lane) {
  LANE_FOR(b, NBAR) {
    double* q = S.qpos + 7 * b;
    normalize4(q + 3);
    quat2mat(S.xmat + 9 * b, q + 3);
    copy3(S.xstale + 3 * b, q);
  }
  WSYNC();
  LANE_FOR(i, 6) geom_center(S, m, 5 * (i / 2) + 1 + (i % 2), S.sph + 3 * i);  // end-cap centres s0..s5
}

TSG_FN void stage_tendon(Scratch& S, const DevModel& m, int lane) {
  LANE_FOR(t, NTEN) {
    double dir[3], w[2][3], p[2][3];
    for (int e = 0; e < 2; e++) {
      int end = 2 * t + e, b = m.tbody[end];
      double v[3];
      mulMV(v, S.xmat + 9 * b, m.tsite[end]);
      add3(p[e], S.qpos + 7 * b, v);
    }
    sub3(dir, p[1], p[0]);
    double len = normalize3(dir);
    double vel = 0;
    for (int e = 0; e < 2; e++) {
      int end = 2 * t + e, b = m.tbody[end];
      double r[3], c[3];
      sub3(r, p[e], S.qpos + 7 * b);
      cross3(c, r, dir);
      mulMTV(w[e], S.xmat + 9 * b, c);
      copy3(S.u.ten.tJw + 3 * end, w[e]);
      double s = e ? 1.0 : -1.0;
      vel += s * (dot3(dir, S.qvel + 6 * b) + dot3(w[e], S.qvel + 6 * b + 3));
    }
    copy3(S.u.ten.tdir + 3 * t, dir);
    S.tlen[t] = len;
    double frc = 0, Bt = -m.tdamp[t];
    if (m.tk[t] > 0) {
      if (len > m.tls[t][1]) frc = m.tk[t] * (m.tls[t][1] - len);
      else if (len < m.tls[t][0]) frc = m.tk[t] * (m.tls[t][0] - len);
    }
    frc -= m.tdamp[t] * vel;
    int a = m.ten_act[t];
    if (a >= 0) {
      double ctrl = S.ctrl[a], input;
      if (m.ctrllimited) ctrl = clampd(ctrl, m.ctrlrange[0], m.ctrlrange[1]);
      if (m.dyntype) { S.actdot[a] = tsg_div(ctrl - S.act[a], fmax(MINVAL, m.dynprm0)); input = S.act[a]; }
      else { S.actdot[a] = 0; input = ctrl; }
      double f = m.gain * input + m.bias[0] + m.bias[1] * len + m.bias[2] * vel;
      bool clamped = false;
      if (m.forcelimited) {
        clamped = (f <= m.forcerange[0] || f >= m.forcerange[1]);
        f = clampd(f, m.forcerange[0], m.forcerange[1]);
      }
      frc += f;
      if (m.bias[2] != 0 && ((m.flags & 1u) || !clamped)) Bt += m.bias[2];
    }
    S.u.ten.tfrc[t] = frc; S.u.ten.tB[t] = Bt;
  }
  WSYNC();
}

// qfrc_smooth / qacc_smooth
TSG_FN void stage_smooth(Scratch& S, const DevModel& m, int lane) {
  LANE_FOR(i, NV) {
    int b = i / 6, j = i % 6;
    double f = 0;
    for (int n = 0; n < m.nends[b]; n++) {
      int end = m.ends[b][n], t = end >> 1;
      double s = (end & 1) ? 1.0 : -1.0;
      double Jv = j < 3 ? S.u.ten.tdir[3 * t + j] : S.u.ten.tJw[3 * end + j - 3];
      f += s * S.u.ten.tfrc[t] * Jv;
    }
    double bias;
    if (j < 3) bias = -m.M[i] * m.grav[j];
    else {
      const double* w = S.qvel + 6 * b + 3; const double* I = m.inertia[b];
      int k = j - 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      bias = w[k1] * (I[k2] * w[k2]) - w[k2] * (I[k1] * w[k1]);
    }
    S.fsm[i] = f - bias;
    S.asmooth[i] = (f - bias) * m.invM[i];
  }
  WSYNC();
}
// per-bar blocks of d(passive + actuator force)/d(qvel) (implicitfast), computed after the solver
TSG_FN void stage_damping_blocks(Scratch& S, const DevModel& m, int lane) {
  LANE_FOR(idx, 63) {
    int b = idx / 21, tri = idx % 21;
    int r = m.tri_i[tri], c = m.tri_j[tri];
    double v = 0;
    for (int n = 0; n < m.nends[b]; n++) {
      int end = m.ends[b][n], t = end >> 1;
      double Jr = r < 3 ? S.u.ten.tdir[3 * t + r] : S.u.ten.tJw[3 * end + r - 3];
      double Jc = c < 3 ? S.u.ten.tdir[3 * t + c] : S.u.ten.tJw[3 * end + c - 3];
      v += S.u.ten.tB[t] * Jr * Jc;
    }
    S.u.post.Dblk[b][tri] = v;
  }
  WSYNC();
}

// ------------------------------------------------------------------ MPR (libccd semantics)
// type 100 = height-field prism, stored as its three (x, y) columns with the top heights and the common base height:
// vertex i < 3 is (px[i], py[i], pbase), vertex 3 + i is (px[i], py[i], pz[i])  (10 values instead of an 18-double
// vertex array behind a pointer: the support function is the hot spot of the height-field path)
struct CObj { int type, bar; double pos[3]; double size[2]; double px[3], py[3], pz[3], pbase; };   // bar: whose rotation matrix
// The narrow phase is compiled once, out of line; its objects name their bar instead of carrying a pointer to the
// bar's rotation matrix, which the support function then reads straight from the env's shared-memory scratch
// (a pointer stored in the object would make every matrix access a generic load).
#if TSG_DEVICE
#define XM_PARAMS int
#define XM_ARGS 0
#define XM_PASS 0
#define XM_BIND                                                                                              \
  extern __shared__ __align__(16) unsigned char tsg_smem[];                                                  \
  const double* xmat_base = reinterpret_cast<const EnvScratch*>(tsg_smem + SMEM_MODEL + SMEM_CFG + (threadIdx.x / TSG_VW) * SMEM_SCRATCH)->xmat;
#else
#define XM_PARAMS const double* xmat_base
#define XM_ARGS S.xmat
#define XM_PASS xmat_base
#define XM_BIND
#endif
struct Supp { double v[3], v1[3]; };  // v = v1 - v2 ; v2 recovered as v1 - v

TSG_FN bool ccd_is_zero(double x) { return fabs(x) < CCD_EPS; }
TSG_FN bool ccd_eq(double a_, double b_) {
  double ab = fabs(a_ - b_);
  if (ab < CCD_EPS) return true;
  double a = fabs(a_), b = fabs(b_);
  return (b > a) ? (ab < CCD_EPS * b) : (ab < CCD_EPS * a);
}
TSG_FN bool ccd_vec_is_origin(const double* a) { return ccd_eq(a[0], 0) && ccd_eq(a[1], 0) && ccd_eq(a[2], 0); }
TSG_FN void ccd_normalize(double* v) { double s = tsg_inv(tsg_sqrt(dot3(v, v))); v[0] *= s; v[1] *= s; v[2] *= s; }

TSG_FN void obj_support(const CObj& o, const double* dir, double* out, const double* xmat_base) {
  if (o.type == 100) {
    // first maximum of vertex . dir over the vertices in order (bottom 0..2, top 3..5), as libccd's loop finds it
    double t[3], bd = 0; int best = 0;
    for (int j = 0; j < 3; j++) t[j] = o.px[j] * dir[0] + o.py[j] * dir[1];
    for (int i = 0; i < 6; i++) {
      double dd = t[i % 3] + (i < 3 ? o.pbase : o.pz[i - 3]) * dir[2];
      if (i == 0 || dd > bd) { bd = dd; best = i; }
    }
    out[0] = o.px[best % 3]; out[1] = o.py[best % 3]; out[2] = best < 3 ? o.pbase : o.pz[best - 3];
    return;
  }
  double ld[3], res[3];
  const double* R = xmat_base + 9 * o.bar;
  mulMTV(ld, R, dir);
  if (o.type == GEOM_SPHERE) scl3(res, ld, o.size[0]);
  else {
    double tmp = tsg_sqrt(ld[0] * ld[0] + ld[1] * ld[1]);
    if (tmp > MINVAL) { double it = tsg_div(o.size[0], tmp); res[0] = ld[0] * it; res[1] = ld[1] * it; }
    else res[0] = res[1] = 0;
    res[2] = (ld[2] > 0 ? 1.0 : (ld[2] < 0 ? -1.0 : 0.0)) * o.size[1];
  }
  mulMV(out, R, res);
  add3(out, out, o.pos);
}
TSG_FN void obj_center(const CObj& o, double* c) {
  if (o.type == 100) {
    c[0] = c[1] = c[2] = 0;
    for (int i = 0; i < 6; i++) { c[0] += o.px[i % 3]; c[1] += o.py[i % 3]; c[2] += i < 3 ? o.pbase : o.pz[i - 3]; }
    c[0] /= 6; c[1] /= 6; c[2] /= 6;
  } else copy3(c, o.pos);
}
TSG_FN_NOINLINE void mink_support(const CObj& o1, const CObj& o2, const double* dir, Supp& s, XM_PARAMS) {
  XM_BIND
  double nd[3] = {-dir[0], -dir[1], -dir[2]}, v2[3];
  obj_support(o1, dir, s.v1, xmat_base);
  obj_support(o2, nd, v2, xmat_base);
  sub3(s.v, s.v1, v2);
}
TSG_FN void portal_dir(const Supp* p, double* dir) {
  double a[3], b[3];
  sub3(a, p[2].v, p[1].v); sub3(b, p[3].v, p[1].v);
  cross3(dir, a, b); ccd_normalize(dir);
}
TSG_FN bool portal_reach_tol(const Supp* p, const Supp& v4, const double* dir, double tol) {
  double dv4 = dot3(v4.v, dir);
  double d1 = dv4 - dot3(p[1].v, dir), d2 = dv4 - dot3(p[2].v, dir), d3 = dv4 - dot3(p[3].v, dir);
  d1 = fmin(d1, d2); d1 = fmin(d1, d3);
  return ccd_eq(d1, tol) || d1 < tol;
}
TSG_FN void expand_portal(Supp* p, const Supp& v4) {
  double v4v0[3];
  cross3(v4v0, v4.v, p[0].v);
  if (dot3(p[1].v, v4v0) > 0) { if (dot3(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4; }
  else { if (dot3(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4; }
}
TSG_FN double seg_dist2_origin(const double* x0, const double* b, double* wit) {
  double d[3];
  sub3(d, b, x0);
  double t = -1.0 * dot3(x0, d); t /= dot3(d, d);
  if (t < 0 || ccd_is_zero(t)) copy3(wit, x0);
  else if (t > 1 || ccd_eq(t, 1.0)) copy3(wit, b);
  else { scl3(wit, d, t); add3(wit, wit, x0); }
  return dot3(wit, wit);
}
TSG_FN double tri_dist2_origin(const double* x0, const double* B, const double* C, double* wit) {
  double d1[3], d2[3];
  sub3(d1, B, x0); sub3(d2, C, x0);
  double v = dot3(d1, d1), w = dot3(d2, d2), p = dot3(x0, d1), q = dot3(x0, d2), r = dot3(d1, d2);
  double s, t, dist, dd = w * v - r * r;
  if (ccd_is_zero(dd)) s = t = -1;
  else { s = tsg_div(q * r - w * p, dd); t = tsg_div(-s * r - q, w); }
  if ((ccd_is_zero(s) || s > 0) && (ccd_eq(s, 1.0) || s < 1) && (ccd_is_zero(t) || t > 0) &&
      (ccd_eq(t, 1.0) || t < 1) && (ccd_eq(t + s, 1.0) || t + s < 1)) {
    scl3(d1, d1, s); scl3(d2, d2, t);
    copy3(wit, x0); add3(wit, wit, d1); add3(wit, wit, d2);
    dist = dot3(wit, wit);
  } else {
    double w2[3], dist2;
    dist = seg_dist2_origin(x0, B, wit);
    dist2 = seg_dist2_origin(x0, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
    dist2 = seg_dist2_origin(B, C, w2);
    if (dist2 < dist) { dist = dist2; copy3(wit, w2); }
  }
  return dist;
}
// returns true on penetration; dir points from obj1 to obj2
TSG_FN_NOINLINE bool mpr_penetration(const CObj& o1, const CObj& o2, double tol, int max_iter,
                                     double* depth, double* dir_out, double* pos_out, XM_PARAMS) {
  Supp p[4], v4;
  double dir[3], va[3], vb[3], dot, c2[3];
  // ---- discoverPortal
  obj_center(o1, p[0].v1); obj_center(o2, c2);
  sub3(p[0].v, p[0].v1, c2);
  if (ccd_vec_is_origin(p[0].v)) p[0].v[0] += CCD_EPS * 10.0;
  scl3(dir, p[0].v, -1.0); ccd_normalize(dir);
  mink_support(o1, o2, dir, p[1], XM_PASS);
  dot = dot3(p[1].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return false;
  cross3(dir, p[0].v, p[1].v);
  if (ccd_is_zero(dot3(dir, dir))) {
    // origin on v1 (touching: depth 0, no direction -> MuJoCo drops it) or on the v0-v1 segment
    if (ccd_vec_is_origin(p[1].v)) return false;
    double v2[3];
    sub3(v2, p[1].v1, p[1].v);
    add3(pos_out, p[1].v1, v2); scl3(pos_out, pos_out, 0.5);
    copy3(dir_out, p[1].v); *depth = tsg_sqrt(dot3(dir_out, dir_out)); ccd_normalize(dir_out);
    return true;
  }
  ccd_normalize(dir);
  mink_support(o1, o2, dir, p[2], XM_PASS);
  dot = dot3(p[2].v, dir);
  if (ccd_is_zero(dot) || dot < 0) return false;
  sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
  cross3(dir, va, vb); ccd_normalize(dir);
  if (dot3(dir, p[0].v) > 0) { Supp t = p[1]; p[1] = p[2]; p[2] = t; scl3(dir, dir, -1.0); }
  for (;;) {
    bool cont = false;
    mink_support(o1, o2, dir, p[3], XM_PASS);
    dot = dot3(p[3].v, dir);
    if (ccd_is_zero(dot) || dot < 0) return false;
    cross3(va, p[1].v, p[3].v); dot = dot3(va, p[0].v);
    if (dot < 0 && !ccd_is_zero(dot)) { p[2] = p[3]; cont = true; }
    if (!cont) {
      cross3(va, p[3].v, p[2].v); dot = dot3(va, p[0].v);
      if (dot < 0 && !ccd_is_zero(dot)) { p[1] = p[3]; cont = true; }
    }
    if (!cont) break;
    sub3(va, p[1].v, p[0].v); sub3(vb, p[2].v, p[0].v);
    cross3(dir, va, vb); ccd_normalize(dir);
  }
  // ---- refinePortal
  for (;;) {
    portal_dir(p, dir);
    dot = dot3(dir, p[1].v);
    if (ccd_is_zero(dot) || dot > 0) break;
    mink_support(o1, o2, dir, v4, XM_PASS);
    dot = dot3(v4.v, dir);
    if (!(ccd_is_zero(dot) || dot > 0) || portal_reach_tol(p, v4, dir, tol)) return false;
    expand_portal(p, v4);
  }
  // ---- findPenetr
  for (int it = 0;; it++) {
    portal_dir(p, dir);
    mink_support(o1, o2, dir, v4, XM_PASS);
    if (portal_reach_tol(p, v4, dir, tol) || it > max_iter) {
      double pdir[3];
      *depth = tsg_sqrt(tri_dist2_origin(p[1].v, p[2].v, p[3].v, pdir));
      if (ccd_is_zero(pdir[0]) && ccd_is_zero(pdir[1]) && ccd_is_zero(pdir[2])) copy3(pdir, dir);
      ccd_normalize(pdir);
      copy3(dir_out, pdir);
      // findPos: barycentric blend of the witness points
      double vec[3], b[4], sum;
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
      double inv = tsg_inv(sum), p1[3] = {0, 0, 0}, p2[3] = {0, 0, 0};
      // v0's second witness is obj2's centre
      for (int i = 0; i < 4; i++) {
        double v2[3];
        if (i == 0) copy3(v2, c2); else sub3(v2, p[i].v1, p[i].v);
        addscl3(p1, p[i].v1, b[i]); addscl3(p2, v2, b[i]);
      }
      scl3(p1, p1, inv); scl3(p2, p2, inv);
      add3(pos_out, p1, p2); scl3(pos_out, pos_out, 0.5);
      return true;
    }
    expand_portal(p, v4);
  }
}

// ------------------------------------------------------------------ collision
TSG_FN void make_frame(double* f) {
  normalize3(f);
  f[3] = f[4] = f[5] = 0;
  if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1;
  double t = dot3(f, f + 3);
  addscl3(f + 3, f, -t);
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}
TSG_FN void set_contact(Con& c, int b1, int b2, double dist, const double* pos, const double* normal) {
  c.b1 = b1; c.b2 = b2; c.dist = dist;
  copy3(c.pos, pos);
  copy3(c.frame, normal);
  make_frame(c.frame);
  c.active = dist < 0 ? 1 : 0;  // includemargin 0: dist >= 0 gives no rows
}
// squared distance between segments p1 +- a1, p2 +- a2 (a = half-axis vectors)
TSG_FN double segseg_dist2(const double* p1, const double* a1, const double* p2, const double* a2) {
  double r[3]; sub3(r, p1, p2);
  double A = dot3(a1, a1), E = dot3(a2, a2), Bq = dot3(a1, a2), C = dot3(a1, r), F = dot3(a2, r);
  double den = A * E - Bq * Bq, s = 0, t;
  if (den > 1e-30) s = clampd(tsg_fdiv(Bq * F - C * E, den), -1.0, 1.0);
  t = tsg_fdiv(Bq * s + F, E);
  if (t < -1.0) { t = -1.0; s = clampd(tsg_fdiv(-Bq - C, A), -1.0, 1.0); }
  else if (t > 1.0) { t = 1.0; s = clampd(tsg_fdiv(Bq - C, A), -1.0, 1.0); }
  double d[3] = {r[0] + s * a1[0] - t * a2[0], r[1] + s * a1[1] - t * a2[1], r[2] + s * a1[2] - t * a2[2]};
  return dot3(d, d);
}
TSG_FN double ptseg_dist2(const double* c, const double* p, const double* a) {
  double r[3]; sub3(r, c, p);
  double t = clampd(tsg_fdiv(dot3(r, a), dot3(a, a)), -1.0, 1.0);
  addscl3(r, a, -t);
  return dot3(r, r);
}

TSG_FN_NOINLINE void plane_cylinder_points(const DevModel& m, const double* pos2, const double* axis_in, double radius, double half,
                                  const double* xaxis, int& cnt, double dist[4], double pts[4][3]) {
  const double* normal = m.fnormal;
  double axis[3], vec[3];
  copy3(axis, axis_in);
  double prjaxis = dot3(normal, axis);
  if (prjaxis > 0) { scl3(axis, axis, -1); prjaxis = -prjaxis; }
  sub3(vec, pos2, m.fpos);
  double dist0 = dot3(vec, normal);
  scl3(vec, axis, prjaxis); sub3(vec, vec, normal);
  double len_sqr = dot3(vec, vec);
  if (len_sqr >= MINVAL * MINVAL) scl3(vec, vec, tsg_div(radius, tsg_sqrt(len_sqr)));
  else scl3(vec, xaxis, radius);
  double prjvec = dot3(vec, normal);
  scl3(axis, axis, half); prjaxis *= half;
  cnt = 0;
  if (dist0 + prjaxis + prjvec <= 0) {
    dist[cnt] = dist0 + prjaxis + prjvec;
    add3(pts[cnt], pos2, vec); add3(pts[cnt], pts[cnt], axis); addscl3(pts[cnt], normal, -dist[cnt] * 0.5);
    cnt++;
  } else return;
  if (dist0 - prjaxis + prjvec <= 0) {
    dist[cnt] = dist0 - prjaxis + prjvec;
    add3(pts[cnt], pos2, vec); sub3(pts[cnt], pts[cnt], axis); addscl3(pts[cnt], normal, -dist[cnt] * 0.5);
    cnt++;
  }
  double prjvec1 = -prjvec * 0.5;
  if (dist0 + prjaxis + prjvec1 <= 0) {
    double vec1[3];
    cross3(vec1, vec, axis); normalize3(vec1); scl3(vec1, vec1, radius * tsg_sqrt(3.0) * 0.5);
    for (int s = 0; s < 2; s++) {
      dist[cnt] = dist0 + prjaxis + prjvec1;
      add3(pts[cnt], pos2, axis); addscl3(pts[cnt], vec1, s ? -1.0 : 1.0); addscl3(pts[cnt], vec, -0.5);
      addscl3(pts[cnt], normal, -dist[cnt] * 0.5);
      cnt++;
    }
  }
}

// floor = plane: analytic plane-sphere / plane-cylinder
TSG_FN void collide_plane(Scratch& S, const DevModel& m, int lane, int& nslot) {
  LANE_FOR_ALL(g, NGEOM) {
    int need = 0, cnt = 0;
    double dist[4], pts[4][3];
    if (g < NGEOM) {
      int b = g / 5;
      double c[3], tmp[3];
      geom_center(S, m, g, c);
      sub3(tmp, c, m.fpos);
      double cdist = dot3(tmp, m.fnormal);
      if (m.gtype[g] == GEOM_SPHERE) {
        double r = m.gsize[g][0];
        if (cdist <= r) { cnt = 1; dist[0] = cdist - r; copy3(pts[0], c); addscl3(pts[0], m.fnormal, -dist[0] / 2 - r); }
      } else if (cdist <= tsg_sqrt(m.gsize[g][0] * m.gsize[g][0] + m.gsize[g][1] * m.gsize[g][1])) {
        const double* R = S.xmat + 9 * b;
        double axis[3] = {R[2], R[5], R[8]}, xaxis[3] = {R[0], R[3], R[6]};
        plane_cylinder_points(m, c, axis, m.gsize[g][0], m.gsize[g][1], xaxis, cnt, dist, pts);
      }
      need = cnt;
    }
    int slot = scan_slot(need, nslot, lane);
    if (need) {
      int b = g / 5;
      for (int k = 0; k < cnt; k++) {
        if (slot + k < MAXC) set_contact(con_at(S, slot + k), -1, b, dist[k], pts[k], m.fnormal);
        else S.overflow = 1;
      }
    }
  }
}

// floor = height field (hfield frame axis-aligned at fpos): geom AABBs -> prism candidates -> MPR
TSG_FN void hf_prism(const DevModel& m, int r, int cmin, int k, CObj& o) {
  // prism k of row r: vertices n = k, k+1, k+2 of the strip (c = cmin + n/2, i = n%2 -> row r+1 / r)
  double dx = m.hdx, dy = m.hdy;
  for (int j = 0; j < 3; j++) {
    int n = k + j, c = cmin + n / 2, rr = r + ((n & 1) ? 0 : 1);
    o.px[j] = dx * c - m.hsize[0]; o.py[j] = dy * rr - m.hsize[1];
    o.pz[j] = (double)m.hdata[rr * m.ncol + c] * m.hsize[2];
  }
  o.pbase = -m.hsize[3];
}
// Conservative cull before MPR: a prism lies entirely on or below the plane through its three top vertices, so a geom
// whose support point towards that plane is still above it (by more than a rounding slack) cannot touch the prism --
// MPR would report "no contact" for it (MuJoCo runs MPR on every prism that passes the height test; dropping pairs
// that cannot collide does not change the contact list).  Sphere: n.c - r|n|; cylinder: n.c - (hl |n.a| + r |n x a|).
TSG_FN bool hf_above_top_plane(const Scratch& S, const DevModel& m, int g, int r, int cmin, int k) {
  double top[9], gc[3], pos[3], e1[3], e2[3], n[3];
  for (int j = 0; j < 3; j++) {   // the three top vertices of hf_prism(m, r, cmin, k)
    int q = k + j, c = cmin + q / 2, rr = r + ((q & 1) ? 0 : 1);
    top[3 * j] = m.hdx * c - m.hsize[0]; top[3 * j + 1] = m.hdy * rr - m.hsize[1];
    top[3 * j + 2] = (double)m.hdata[rr * m.ncol + c] * m.hsize[2];
  }
  sub3(e1, top + 3, top); sub3(e2, top + 6, top);
  cross3(n, e1, e2);
  if (n[2] < 0) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; }
  double nn = dot3(n, n);
  if (!(nn > 0)) return false;
  geom_center(S, m, g, gc);
  sub3(pos, gc, m.fpos);
  double rad = m.gsize[g][0], reach;
  if (m.gtype[g] == GEOM_SPHERE) reach = rad * tsg_sqrt(nn);
  else {
    const double* R = S.xmat + 9 * (g / 5);
    double na = n[0] * R[2] + n[1] * R[5] + n[2] * R[8];
    reach = m.gsize[g][1] * fabs(na) + rad * tsg_sqrt(fmax(0.0, nn - na * na));
  }
  double sep = dot3(n, pos) - reach - dot3(n, top);
  return sep > 1e-9 * tsg_sqrt(nn);
}
// flags the (geom, prism) items [base, base + span) whose prism top reaches the geom's AABB bottom and lists them, in
// order, in S.u.col.cand (at most MAXCAND entries are stored); returns how many were flagged
TSG_FN int hf_flag_items(Scratch& S, const DevModel& m, int lane, int base, int span) {
  constexpr int PMAX = 24;
  int ncand = 0;
  LANE_FOR_ALL(ii, span) {   // every lane takes part in the scan
    int i = base + ii, g = i / PMAX, p = i % PMAX, flag = 0;
    if (ii < span && g < NGEOM) {
      int per_row = S.u.col.hf_cell[g][1], nrows = S.u.col.hf_cell[g][3];
      if (per_row > 0 && p < per_row * nrows) {
        int r = S.u.col.hf_cell[g][2] + p / per_row, k = p % per_row, cmin = S.u.col.hf_cell[g][0];
        double zmin = S.u.col.hf_zmin[g];
        for (int j = 0; j < 3; j++) {
          int n = k + j, c = cmin + n / 2, rr = r + ((n & 1) ? 0 : 1);
          if ((double)m.hdata[rr * m.ncol + c] * m.hsize[2] >= zmin) flag = 1;
        }
        if (flag && hf_above_top_plane(S, m, g, r, cmin, k)) flag = 0;
      }
      if (per_row > 0 && p == PMAX - 1 && per_row * nrows > PMAX) S.overflow = 1;
    }
    int slot = scan_slot(flag, ncand, lane);
    if (flag && slot < MAXCAND) S.u.col.cand[slot] = i;
  }
  WSYNC();
  return ncand;
}
TSG_FN void collide_hfield(Scratch& S, const DevModel& m, int lane, int& nslot) {
  LANE_FOR(g, NGEOM) {
    int b = g / 5;
    double pos[3], gc[3];
    geom_center(S, m, g, gc);
    sub3(pos, gc, m.fpos);
    double r = m.gsize[g][0], hl = m.gsize[g][1];
    double rb = m.gtype[g] == GEOM_SPHERE ? r : tsg_sqrt(r * r + hl * hl);
    bool ok = true;
    for (int i = 0; i < 2; i++) if (m.hsize[i] < pos[i] - rb || -m.hsize[i] > pos[i] + rb) ok = false;
    if (m.hsize[2] < pos[2] - rb || -m.hsize[3] > pos[2] + rb) ok = false;
    double ext[3];
    if (m.gtype[g] == GEOM_SPHERE) ext[0] = ext[1] = ext[2] = r;
    else { // AABB half extents of a cylinder = what the +-axis support queries return
      const double* R = S.xmat + 9 * b;
      for (int i = 0; i < 3; i++) {
        double az = R[3 * i + 2];
        ext[i] = r * tsg_sqrt(fmax(0.0, 1.0 - az * az)) + hl * fabs(az);
      }
    }
    double xmin = pos[0] - ext[0], xmax = pos[0] + ext[0], ymin = pos[1] - ext[1], ymax = pos[1] + ext[1];
    double zmin = pos[2] - ext[2], zmax = pos[2] + ext[2];
    if (xmin > m.hsize[0] || xmax < -m.hsize[0] || ymin > m.hsize[1] || ymax < -m.hsize[1] || zmin > m.hsize[2] || zmax < -m.hsize[3]) ok = false;
    int cmin = (int)floor(tsg_div(xmin + m.hsize[0], 2 * m.hsize[0]) * (m.ncol - 1));
    int cmax = (int)ceil(tsg_div(xmax + m.hsize[0], 2 * m.hsize[0]) * (m.ncol - 1));
    int rmin = (int)floor(tsg_div(ymin + m.hsize[1], 2 * m.hsize[1]) * (m.nrow - 1));
    int rmax = (int)ceil(tsg_div(ymax + m.hsize[1], 2 * m.hsize[1]) * (m.nrow - 1));
    if (cmin < 0) cmin = 0;
    if (cmax > m.ncol - 1) cmax = m.ncol - 1;
    if (rmin < 0) rmin = 0;
    if (rmax > m.nrow - 1) rmax = m.nrow - 1;
    int per_row = 2 * (cmax - cmin + 1) - 2;
    if (!ok || per_row <= 0 || rmax <= rmin) { per_row = 0; rmax = rmin; }
    S.u.col.hf_cell[g][0] = cmin; S.u.col.hf_cell[g][1] = per_row; S.u.col.hf_cell[g][2] = rmin; S.u.col.hf_cell[g][3] = rmax - rmin;
    S.u.col.hf_zmin[g] = zmin;
  }
  WSYNC();
  // candidates (geom, prism) in MuJoCo's order.  All NGEOM * PMAX items are flagged first; when the whole list fits
  // the MAXCAND slots (the usual case: a dozen candidates) it is handed to MPR in ONE go -- one or two lane-parallel
  // passes per substep instead of one per 64-item block that has a candidate -- else the blocks are processed one at
  // a time, so the list can never overflow.
  constexpr int PMAX = 24, NITEM = NGEOM * PMAX;
  int ntot = hf_flag_items(S, m, lane, 0, NITEM);
  const bool one = ntot <= MAXCAND;
  TSG_UNROLL1
  for (int base = 0; base < NITEM; base += (one ? NITEM : MAXCAND)) {
    int ncand = one ? ntot : hf_flag_items(S, m, lane, base, MAXCAND);
    LANE_FOR_ALL(n, ncand) {
      bool hit = false;
      double depth = 0, dir[3] = {0, 0, 1}, pos[3] = {0, 0, 0};
      int b = 0;
      if (n < ncand) {
        int i = S.u.col.cand[n], g = i / PMAX, p = i % PMAX;
        b = g / 5;
        int per_row = S.u.col.hf_cell[g][1];
        double gc[3];
        CObj o1, o2;
        hf_prism(m, S.u.col.hf_cell[g][2] + p / per_row, S.u.col.hf_cell[g][0], p % per_row, o1);
        o1.type = 100; o1.bar = 0;
        o2.type = m.gtype[g]; o2.bar = b;
        geom_center(S, m, g, gc);
        sub3(o2.pos, gc, m.fpos);
        o2.size[0] = m.gsize[g][0]; o2.size[1] = m.gsize[g][1];
        hit = mpr_penetration(o1, o2, m.mpr_tol, m.mpr_iterations, &depth, dir, pos, XM_ARGS);
        if (hit && ccd_vec_is_origin(dir)) hit = false;
        if (hit) {
          add3(pos, pos, m.fpos);
          if ((m.flags & 4u) && o2.type == GEOM_SPHERE) {
            double nn[3]; sub3(nn, gc, pos);
            if (tsg_sqrt(dot3(nn, nn)) > MINVAL) { normalize3(nn); copy3(dir, nn); }
          }
        }
      }
      int slot = scan_slot(hit ? 1 : 0, nslot, lane);
      if (hit) { if (slot < MAXC) set_contact(con_at(S, slot), -1, b, -depth, pos, dir); else S.overflow = 1; }
    }
    if (lane == 0) S.nmpr_total += ncand;
    if (nslot > MAXC) nslot = MAXC;
    WSYNC();
  }
}

// bar-bar: 75 geom pairs, analytic capsule pre-filter (conservative), then sphere-sphere / MPR
TSG_FN void collide_bars(Scratch& S, const DevModel& m, int lane, int& nslot) {
  TSG_UNROLL1
  for (int base = 0; base < 75; base += MAXCAND) {
    int ncand = 0;
    const int span = 75 - base < MAXCAND ? 75 - base : MAXCAND;   // the last block holds 11 pairs: one lane pass
    LANE_FOR_ALL(ii, span) {
      int flag = 0, i = base + ii;
      if (ii < span) {
        int pr = i / 25, b1 = pr == 2 ? 1 : 0, b2 = pr == 0 ? 1 : 2;
        int g1 = 5 * b1 + (i % 25) / 5, g2 = 5 * b2 + i % 5;
        int t1 = m.gtype[g1], t2 = m.gtype[g2];
        double c1[3], c2[3];
        geom_center(S, m, g1, c1); geom_center(S, m, g2, c2);
        double rs = m.gsize[g1][0] + m.gsize[g2][0] + 1e-6;
        const double *R1 = S.xmat + 9 * b1, *R2 = S.xmat + 9 * b2;
        double a1[3] = {R1[2] * m.gsize[g1][1], R1[5] * m.gsize[g1][1], R1[8] * m.gsize[g1][1]};
        double a2[3] = {R2[2] * m.gsize[g2][1], R2[5] * m.gsize[g2][1], R2[8] * m.gsize[g2][1]};
        double d2;
        if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) { double d[3]; sub3(d, c1, c2); d2 = dot3(d, d); }
        else if (t1 == GEOM_SPHERE) d2 = ptseg_dist2(c1, c2, a2);
        else if (t2 == GEOM_SPHERE) d2 = ptseg_dist2(c2, c1, a1);
        else d2 = segseg_dist2(c1, a1, c2, a2);
        flag = d2 <= rs * rs;
      }
      int slot = scan_slot(flag, ncand, lane);
      if (flag) S.u.col.cand[slot] = i;
    }
    WSYNC();
    LANE_FOR_ALL(n, ncand) {
      bool hit = false;
      double dist = 1, pos[3] = {0, 0, 0}, nrm[3] = {1, 0, 0};
      int b1 = 0, b2 = 1;
      if (n < ncand) {
        int i = S.u.col.cand[n];
        int pr = i / 25;
        b1 = pr == 2 ? 1 : 0; b2 = pr == 0 ? 1 : 2;
        int g1 = 5 * b1 + (i % 25) / 5, g2 = 5 * b2 + i % 5;
        if (m.gtype[g1] > m.gtype[g2]) { int t = g1; g1 = g2; g2 = t; t = b1; b1 = b2; b2 = t; }  // lower type first
        double c1[3], c2[3];
        geom_center(S, m, g1, c1); geom_center(S, m, g2, c2);
        if (m.gtype[g2] == GEOM_SPHERE) {
          sub3(nrm, c2, c1);
          double len = normalize3(nrm), r1 = m.gsize[g1][0];
          dist = len - r1 - m.gsize[g2][0];
          hit = dist <= 0;
          copy3(pos, c1); addscl3(pos, nrm, r1 + dist / 2);
        } else {
          CObj o1, o2;
          o1.type = m.gtype[g1]; o1.bar = b1; copy3(o1.pos, c1);
          o1.size[0] = m.gsize[g1][0]; o1.size[1] = m.gsize[g1][1];
          o2.type = m.gtype[g2]; o2.bar = b2; copy3(o2.pos, c2);
          o2.size[0] = m.gsize[g2][0]; o2.size[1] = m.gsize[g2][1];
          double depth;
          hit = mpr_penetration(o1, o2, m.mpr_tol, m.mpr_iterations, &depth, nrm, pos, XM_ARGS);
          if (hit && ccd_vec_is_origin(nrm)) hit = false;
          dist = -depth;
          if (hit && (m.flags & 4u) && o1.type == GEOM_SPHERE) {
            double nn[3]; sub3(nn, pos, c1);
            if (tsg_sqrt(dot3(nn, nn)) > MINVAL) { normalize3(nn); copy3(nrm, nn); }
          }
        }
      }
      int slot = scan_slot(hit ? 1 : 0, nslot, lane);
      if (hit) { if (slot < MAXC) set_contact(con_at(S, slot), b1, b2, dist, pos, nrm); else S.overflow = 1; }
    }
    if (lane == 0) S.nmpr_total += ncand;
    if (nslot > MAXC) nslot = MAXC;
    WSYNC();
  }
}

TSG_FN double impedance(const DevModel& m, double pos) {
  double d0 = clampd(m.solimp[0], MINIMP, MAXIMP), dw = clampd(m.solimp[1], MINIMP, MAXIMP);
  double width = fmax(MINVAL, m.solimp[2]), mid = clampd(m.solimp[3], MINIMP, MAXIMP), power = fmax(1.0, m.solimp[4]);
  if (d0 == dw || width <= MINVAL) return 0.5 * (d0 + dw);
  double x = tsg_fdiv(fabs(pos), width), y;
  if (x >= 1) return dw;
  if (x == 0) return d0;
  if (power == 1) y = x;
  else if (power == 2) y = (x <= mid) ? tsg_inv(mid) * (x * x) : 1 - tsg_inv(1 - mid) * ((1 - x) * (1 - x));
  else if (x <= mid) y = (1 / pow(mid, power - 1)) * pow(x, power);
  else y = 1 - (1 / pow(1 - mid, power - 1)) * pow(1 - x, power);
  return d0 + y * (dw - d0);
}

// collision + constraint rows; leaves S.nact active contacts listed in S.order
TSG_FN void stage_constraint(Scratch& S, const DevModel& m, int lane) {
  int nslot = 0;
  if (m.floor_type == 0) { collide_plane(S, m, lane, nslot); if (nslot > MAXC) nslot = MAXC; WSYNC(); }
  else collide_hfield(S, m, lane, nslot);
  collide_bars(S, m, lane, nslot);
  // compact the active slots
  int nact = 0;
  LANE_FOR_ALL(s, MAXC) {
    int flag = (s < nslot) && con_at(S, s < nslot ? s : 0).active;
    int k = scan_slot(flag, nact, lane);
    if (flag) S.order[k] = s;
  }
  if (lane == 0) { S.nact = nact; S.nslot = nslot; }
  WSYNC();
  // Jacobian blocks: item = (contact, side, axis)
  LANE_FOR(i, nact * 6) {
    Con& c = con_at(S, S.order[i / 6]);
    int side = (i % 6) / 3, ax = i % 3, b = side ? c.b2 : c.b1;
    double* Jt = c.Jt[side][ax];
    double* Jr = c.Jr[side][ax];
    if (b < 0) { for (int k = 0; k < 3; k++) { Jt[k] = 0; Jr[k] = 0; } }
    else {
      double s = side ? 1.0 : -1.0, rr[3], t[3], w[3];
      const double* a = c.frame + 3 * ax;
      sub3(rr, c.pos, S.qpos + 7 * b);
      cross3(t, rr, a);
      mulMTV(w, S.xmat + 9 * b, t);
      Jt[0] = s * w[0]; Jt[1] = s * w[1]; Jt[2] = s * w[2];
      mulMTV(w, S.xmat + 9 * b, a);
      Jr[0] = s * w[0]; Jr[1] = s * w[1]; Jr[2] = s * w[2];
    }
  }
  WSYNC();
  // rows: velocity, impedance, reference acceleration; item = (contact, row)
  LANE_FOR(i, nact * 6) {
    Con& c = con_at(S, S.order[i / 6]);
    int r = i % 6;
    double vel = Jrow_dot_all(c, r, S.qvel);
    double imp = impedance(m, c.dist);
    if (r == 0) {
      double tran = (c.b1 >= 0 ? m.invw_tran[c.b1] : 0.0) + m.invw_tran[c.b2];
      c.D0 = tsg_inv(fmax(MINVAL, tsg_fdiv(1 - imp, imp) * tran));
    }
    c.aref[r] = -m.B * vel - (r ? 0.0 : m.K * imp * c.dist);
  }
  WSYNC();
}

// ------------------------------------------------------------------ Newton solver
// (every routine that is called from more than one place exists once, out of line: see tsg_sqrt above)
enum { VEC_SMOOTH = 0, VEC_WARM = 1, VEC_QACC = 2 };
// jar = J a - aref for a = qacc_smooth / warm start / qacc
TSG_FN_NOINLINE void compute_jar(int which, CTX_PARAMS) {
  CTX_BIND
  const double* a = which == VEC_SMOOTH ? S.asmooth : (which == VEC_WARM ? S.warm : S.qacc);
  LANE_FOR(i, S.nact * 6) {
    Con& k = con_at(S, S.order[i / 6]);
    int r = i % 6;
    k.jar[r] = Jrow_dot_all(k, r, a) - k.aref[r];
  }
  WSYNC();
}
// mj_constraintUpdate for one elliptic contact; returns its cost
TSG_FN double con_update(Con& c, const DevModel& m, bool full) {
  double U[6], T = 0, mu = m.mu;
  U[0] = c.jar[0] * mu;
  for (int j = 1; j < 6; j++) { U[j] = c.jar[j] * m.fr[j - 1]; T += U[j] * U[j]; }
  double N = U[0];
  T = tsg_sqrt_inl(T);
  if (N >= mu * T || (T <= 0 && N >= 0)) {
    if (full) { for (int j = 0; j < 6; j++) c.force[j] = 0; c.zone = ZONE_TOP; c.wcoef = c.ca = c.cb = 0; }
    return 0;
  }
  if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
    double s = 0;
    for (int j = 0; j < 6; j++) {
      double D = c.D0 * m.dscale[j];
      s += 0.5 * D * c.jar[j] * c.jar[j];
      if (full) c.force[j] = -D * c.jar[j];
    }
    if (full) { c.zone = ZONE_BOTTOM; c.wcoef = c.D0; c.ca = c.cb = 0; }
    return s;
  }
  double Dm = c.D0 * m.inv_mu2, NT = N - mu * T;
  if (full) {
    double invT = tsg_inv(T);
    c.force[0] = -Dm * NT * mu;
    double kap = mu * mu - mu * N * invT;
    c.su[0] = 0;
    for (int j = 1; j < 6; j++) {
      c.force[j] = -c.force[0] * invT * U[j] * m.fr[j - 1];
      c.su[j] = m.fr[j - 1] * U[j] * invT;
    }
    c.ca = Dm; c.cb = Dm * kap; c.wcoef = c.cb; c.zone = ZONE_MIDDLE;
  }
  return 0.5 * Dm * NT * NT;
}
// total cost at `which` (constraint part summed over contacts through the shared accumulator + Gauss part);
// full: also forces / zones / Hessian weights.  The Gauss part alone is left in S.gauss.
TSG_FN_NOINLINE double total_cost(int which, int full, CTX_PARAMS) {
  CTX_BIND
  LANE_FOR(n, S.nact) S.u.ls.acc[n][0] = con_update(con_at(S, S.order[n]), m, full != 0);
  WSYNC();
  double s = 0;
  TSG_UNROLL1
  for (int n = 0; n < S.nact; n++) s += S.u.ls.acc[n][0];
  double g = 0;
  if (which != VEC_SMOOTH) {
    const double* a = which == VEC_WARM ? S.warm : S.qacc;
    RED_FOR(k, NV) { double d = a[k] - S.asmooth[k]; g += 0.5 * m.M[k] * d * d; }
    g = RED_SUM(g);
  }
  WSYNC();
  S.gauss = g;
  S.cost = s + g;
  return s + g;
}

// H = L D L^T of the packed Hessian in shared memory (lane i owns row i; no square roots; structurally zero
// columns -- bars not coupled by a contact -- are skipped).  The right-hand side rides along as an extra column, so
// the forward substitution costs no extra phases; columns keep their raw entries t_ik = L_ik d_k (scaled on use).
TSG_FN void factor_solve(Scratch& S, int lane) {
  double* H = S.u.hes.H;
  LANE_FOR(i, NV) S.rhs[i] = S.grad[i];
  WSYNC();
  TSG_UNROLL1
  for (int k = 0; k < NV - 1; k++) {
    int kk = k * (k + 1) / 2;
    double dk = fmax(H[kk + k], MINVAL), inv = tsg_rcp(dk);
    double yk = S.rhs[k];   // final: every earlier column has been eliminated
    LANE_FOR(i, NV) {
      if (i > k) {
        double* Hi = H + i * (i + 1) / 2;
        double t = Hi[k];
        if (t != 0.0) {
          double lik = t * inv;
          int o = (k + 1) * (k + 2) / 2 + k;   // H[j][k], j = k + 1, ...
          TSG_UNROLL1
          for (int j = k + 1; j < i; j++) { Hi[j] -= lik * H[o]; o += j + 1; }
          Hi[i] -= lik * t;
          S.rhs[i] -= lik * yk;
        }
      } else if (i == k) S.dinv[k] = inv;
    }
    WSYNC();
  }
  if (lane == 0) S.dinv[NV - 1] = tsg_rcp(fmax(H[NTRI - 1], MINVAL));
  WSYNC();
  LANE_FOR(i, NV) S.rhs[i] *= S.dinv[i];   // z = D^-1 y
  WSYNC();
  TSG_UNROLL1
  for (int k = NV - 1; k > 0; k--) {   // L^T x = z with L_ki = t_ki / d_i
    double xk = S.rhs[k];   // final: rows > k were substituted in earlier steps, one barrier per step is enough
    LANE_FOR(i, NV) if (i < k) { double t = H[k * (k + 1) / 2 + i]; if (t != 0.0) S.rhs[i] -= t * S.dinv[i] * xk; }
    WSYNC();
  }
  LANE_FOR(i, NV) S.search[i] = -S.rhs[i];
  WSYNC();
}

// column k (0..5) of the side's 6x6 Jacobian block
TSG_FN void Jcol(const Con& c, int side, int k, double* col) {
  if (k < 3) {
    double s = side ? 1.0 : -1.0;
    col[0] = s * c.frame[k]; col[1] = s * c.frame[3 + k]; col[2] = s * c.frame[6 + k];
    col[3] = col[4] = col[5] = 0;
  } else {
    col[0] = c.Jt[side][0][k - 3]; col[1] = c.Jt[side][1][k - 3]; col[2] = c.Jt[side][2][k - 3];
    col[3] = c.Jr[side][0][k - 3]; col[4] = c.Jr[side][1][k - 3]; col[5] = c.Jr[side][2][k - 3];
  }
}

// gradient (+ qfrc_constraint), then -- unless `grad_only` -- Hessian, factorisation, search = -H^-1 grad.
// Needs total_cost(.., full) done.  Returns |grad|^2.  The converged last iteration only needs the gradient.
TSG_FN_NOINLINE double newton_direction(int grad_only, double oldcost, double cost, CTX_PARAMS) {
  CTX_BIND
  int nact = S.nact;
  // cone vectors b = sum_j su_j J_j  (item = contact, side, dof)
  LANE_FOR(i, nact * 12) {
    Con& k = con_at(S, S.order[i / 12]);
    if (k.zone == ZONE_MIDDLE) {
      int side = (i % 12) / 6, d = i % 6;
      double col[6], v = 0;
      Jcol(k, side, d, col);
      for (int j = 1; j < 6; j++) v += k.su[j] * col[j];
      k.bvec[side][d] = v;
    }
  }
  // gradient + qfrc_constraint (item = dof)
  double gn = 0;
  LANE_FOR(i, NV) {
    int b = i / 6, d = i % 6;
    double f = 0;
    TSG_UNROLL1
    for (int n = 0; n < nact; n++) {
      const Con& k = con_at(S, S.order[n]);
      int side = k.b2 == b ? 1 : (k.b1 == b ? 0 : -1);
      if (side < 0) continue;
      double col[6];
      Jcol(k, side, d, col);
      for (int r = 0; r < 6; r++) f += col[r] * k.force[r];
    }
    S.fcon[i] = f;
    double gi = m.M[i] * (S.qacc[i] - S.asmooth[i]) - f;
    S.grad[i] = gi;
    gn += gi * gi;
  }
  WSYNC();
#ifdef TSG_NO_WSUM
  gn = 0;
  RED_FOR(k, NV) gn += S.grad[k] * S.grad[k];
#else
  gn = warp_sum(gn);
#endif
  if (grad_only) {  // convergence test of mj_solPrimal: skip the factorisation nobody will use
    double scale = m.solscale, tg = m.tol * m.meaninertia * NV;
    if (scale * (oldcost - cost) < m.tol || gn < tg * tg) return -1.0;
  }
  // Hessian, packed lower triangle: mass matrix, then one contact at a time adds its own block entries
  // (21 for a floor contact, 78 for a bar-bar contact) -- no entry is touched by two lanes in a phase
  LANE_FOR(e, NTRI) S.u.hes.H[e] = (m.tri_i[e] == m.tri_j[e]) ? m.M[m.tri_i[e]] : 0.0;
  WSYNC();
  TSG_UNROLL1
  for (int n = 0; n < nact; n++) {
    const Con& k = con_at(S, S.order[n]);
    if (k.zone == ZONE_TOP) continue;
    const double* wt = m.wtab[k.zone == ZONE_MIDDLE ? 1 : 0];
    int two = k.b1 >= 0;
    LANE_FOR(e, two ? 78 : 21) {
      // local 12x12 (or 6x6) lower triangle: index 0..5 -> side 1 (b2), 6..11 -> side 0 (b1)
      int li = m.tri_i[e], lj = m.tri_j[e];
      int si = li < 6 ? 1 : 0, sj = lj < 6 ? 1 : 0, ki = li % 6, kj = lj % 6;
      int gi = 6 * (si ? k.b2 : k.b1) + ki, gj = 6 * (sj ? k.b2 : k.b1) + kj;
      double ci[6], cj[6], acc = 0;
      Jcol(k, si, ki, ci); Jcol(k, sj, kj, cj);
      for (int r = 0; r < 6; r++) acc += wt[r] * ci[r] * cj[r];
      double v = k.wcoef * acc;
      if (k.zone == ZONE_MIDDLE) {
        double bi_ = k.bvec[si][ki], bj_ = k.bvec[sj][kj];
        double ai = m.mu * (ci[0] - bi_), aj = m.mu * (cj[0] - bj_);
        v += k.ca * ai * aj - k.cb * bi_ * bj_;
      }
      int hi = gi > gj ? gi : gj, lo = gi > gj ? gj : gi;
      S.u.hes.H[hi * (hi + 1) / 2 + lo] += v;
    }
    WSYNC();
  }
  factor_solve(S, lane);
  return gn;
}

struct LsPnt { double alpha, cost, d0, d1; };

// cost and its first two derivatives along the search direction at step alpha -> S.lsr[0..2]
TSG_FN_NOINLINE void ls_eval(double a, CTX_PARAMS) {
  CTX_BIND
  double mu = m.mu;
  LANE_FOR(n, S.nact) {
    const Con& k = con_at(S, S.order[n]);
    double cost = 0, d0 = 0, d1 = 0;
    double N = k.U0 + a * k.V0, Tsqr = k.UU + a * (2 * k.UV + a * k.VV);
    bool bottom = false;
    if (Tsqr <= 0) { if (N < 0) bottom = true; }
    else {
      double T = tsg_sqrt_inl(Tsqr);
      if (N >= mu * T) {}
      else if (mu * N + T <= 0) bottom = true;
      else {
        double invT = tsg_rcp(T);
        double N1 = k.V0, T1 = (k.UV + a * k.VV) * invT;
        double T2 = k.VV * invT - (k.UV + a * k.VV) * T1 * (invT * invT);
        double NT = N - mu * T, Dm = k.D0 * m.inv_mu2;
        cost = 0.5 * Dm * NT * NT;
        d0 = Dm * NT * (N1 - mu * T1);
        d1 = Dm * ((N1 - mu * T1) * (N1 - mu * T1) + NT * (-mu * T2));
      }
    }
    if (bottom) { cost = a * a * k.q2 + a * k.q1 + k.q0; d0 = 2 * a * k.q2 + k.q1; d1 = 2 * k.q2; }
    S.u.ls.acc[n][0] = cost; S.u.ls.acc[n][1] = d0; S.u.ls.acc[n][2] = d1;
  }
  WSYNC();
  double cost = a * a * S.qG[2] + a * S.qG[1] + S.qG[0], d0 = 2 * a * S.qG[2] + S.qG[1], d1 = 2 * S.qG[2];
  TSG_UNROLL1
  for (int n = 0; n < S.nact; n++) { cost += S.u.ls.acc[n][0]; d0 += S.u.ls.acc[n][1]; d1 += S.u.ls.acc[n][2]; }
  WSYNC();
  if (d1 <= 0) d1 = MINVAL;
  S.lsr[0] = cost; S.lsr[1] = d0; S.lsr[2] = d1;  // every lane writes the same values
  if (lane == 0) S.ls_evals++;
  WSYNC();
}
TSG_FN void ls_point(LsPnt& p, double alpha, EnvScratch& S, const DevModel& m, const EnvCfg& c, int lane) {
  p.alpha = alpha;
  ls_eval(alpha, CTX_ARGS);
  p.cost = S.lsr[0]; p.d0 = S.lsr[1]; p.d1 = S.lsr[2];
}
TSG_FN int ls_update_bracket(LsPnt& p, const LsPnt* cand, LsPnt& pnext, EnvScratch& S, const DevModel& m, const EnvCfg& c, int lane) {
  int flag = 0;
  for (int i = 0; i < 3; i++) {
    if (p.d0 < 0 && cand[i].d0 < 0 && p.d0 < cand[i].d0) { p = cand[i]; flag = 1; }
    else if (p.d0 > 0 && cand[i].d0 > 0 && p.d0 > cand[i].d0) { p = cand[i]; flag = 2; }
  }
  if (flag) ls_point(pnext, p.alpha - tsg_fdiv(p.d0, p.d1), S, m, c, lane);
  return flag;
}
// exact line search along S.search from S.qacc (jar current, S.gauss = current Gauss cost); returns alpha
TSG_FN_NOINLINE double line_search(CTX_PARAMS) {
  CTX_BIND
  double snorm = 0, qG1 = 0, qG2 = 0, gs = 0;
  RED_FOR(k, NV) {
    double sk = S.search[k];
    snorm += sk * sk;
    gs += S.grad[k] * sk;
    qG1 += sk * (m.M[k] * S.qacc[k]) - S.fsm[k] * sk;
    qG2 += 0.5 * sk * (m.M[k] * sk);
  }
  RED_SUM4(snorm, gs, qG1, qG2);
  snorm = tsg_sqrt(snorm);
  if (snorm < MINVAL) return 0;
  double gtol = m.tol * m.ls_tol * snorm * (m.meaninertia * NV);
  WSYNC();
  S.qG[0] = S.gauss; S.qG[1] = qG1; S.qG[2] = qG2; S.ls_evals = 0;
  // jv = J search (stored over aref, which is dead by now)
  LANE_FOR(i, S.nact * 6) {
    Con& k = con_at(S, S.order[i / 6]);
    int r = i % 6;
    k.aref[r] = Jrow_dot_all(k, r, S.search);
  }
  WSYNC();
  LANE_FOR(n, S.nact) {
    Con& k = con_at(S, S.order[n]);
    double q0 = 0, q1 = 0, q2 = 0, UU = 0, UV = 0, VV = 0;
    for (int j = 0; j < 6; j++) {
      double D = k.D0 * m.dscale[j], ja = k.jar[j], jv = k.aref[j];
      q0 += 0.5 * D * ja * ja; q1 += D * ja * jv; q2 += 0.5 * D * jv * jv;
      if (j > 0) { double U = ja * m.fr[j - 1], V = jv * m.fr[j - 1]; UU += U * U; UV += U * V; VV += V * V; }
    }
    k.q0 = q0; k.q1 = q1; k.q2 = q2;
    k.U0 = k.jar[0] * m.mu; k.V0 = k.aref[0] * m.mu; k.UU = UU; k.UV = UV; k.VV = VV;
  }
  WSYNC();
  LsPnt p0, p1, p2, pmid, p1next, p2next;
  // alpha = 0 needs no evaluation pass: cost is the current cost, the slope is grad . search and, search being the
  // Newton direction (H search = -grad), the curvature search' H search = -slope
  p0.alpha = 0; p0.cost = S.cost; p0.d0 = gs; p0.d1 = -gs > 0 ? -gs : MINVAL;
  if (lane == 0) S.ls_evals = 1;
  WSYNC();
  ls_point(p1, p0.alpha - tsg_fdiv(p0.d0, p0.d1), S, m, c, lane);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.d0) < gtol) return p1.alpha;
  int dir = p1.d0 < 0 ? 1 : -1, p2update = 0;
  p2 = p1;
  TSG_UNROLL1
  while (p1.d0 * dir <= -gtol && S.ls_evals < m.ls_iterations) {
    p2 = p1; p2update = 1;
    ls_point(p1, p1.alpha - tsg_fdiv(p1.d0, p1.d1), S, m, c, lane);
    if (fabs(p1.d0) < gtol) return p1.alpha;
  }
  if (S.ls_evals >= m.ls_iterations || !p2update) return p1.alpha;
  p2next = p1;
  ls_point(p1next, p1.alpha - tsg_fdiv(p1.d0, p1.d1), S, m, c, lane);
  TSG_UNROLL1
  while (S.ls_evals < m.ls_iterations) {
    ls_point(pmid, 0.5 * (p1.alpha + p2.alpha), S, m, c, lane);
    LsPnt cand[3] = {p1next, p2next, pmid};
    int best = -1; double bestcost = 0;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i].d0) < gtol && (best == -1 || cand[i].cost < bestcost)) { bestcost = cand[i].cost; best = i; }
    if (best >= 0) return cand[best].alpha;
    int b1 = ls_update_bracket(p1, cand, p1next, S, m, c, lane);
    int b2 = ls_update_bracket(p2, cand, p2next, S, m, c, lane);
    if (!b1 && !b2) return pmid.alpha;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}

// Alignment scope: the envs that share a physical warp (TSG_ALIGN_WARP, with TSG_VW < 32) or the whole CTA
// (TSG_ALIGN) walk through the same sequence of phases, idle where they have nothing to do, so that they share
// fetched instructions.  cta_any = barrier + "is any env of the scope still active"; align_sync = plain barrier.
#if TSG_DEVICE && defined(TSG_ALIGN_WARP)
#define TSG_ALIGNED 1
TSG_FN bool align_any(bool go) { return __any_sync(0xffffffffu, go) != 0; }
TSG_FN void align_sync() { __syncwarp(0xffffffffu); }
#elif TSG_DEVICE && defined(TSG_ALIGN)
#define TSG_ALIGNED 1
TSG_FN bool align_any(bool go) { return __syncthreads_or(go ? 1 : 0) != 0; }
TSG_FN void align_sync() { __syncthreads(); }
#else
#define TSG_ALIGNED 0
#endif
TSG_FN bool cta_any(const Scratch& S, bool go) {
#if TSG_ALIGNED && !defined(TSG_ALIGN_SUBSTEP_ONLY)
  if (S.align) return align_any(go);
#endif
  return go;
}
// mj_fwdConstraint: warm-start choice + Newton iterations.  Leaves S.qacc, S.fcon, S.warm.
// In aligned mode every warp of the CTA walks through the same sequence of phases (idle where it has nothing to
// do), which keeps the instruction stream of the SM coherent: the kernel is instruction-fetch bound.
TSG_FN void stage_solve(EnvScratch& S, const DevModel& m, const EnvCfg& c, int lane) {
  bool active = S.nact > 0;
  if (!active) {
    LANE_FOR(i, NV) { S.qacc[i] = S.asmooth[i]; S.warm[i] = S.asmooth[i]; S.fcon[i] = 0; }
    WSYNC();
    if (!S.align) return;
  }
  double cost = 0;
  cta_any(S, true);
  if (active) {
    // cost at qacc_smooth (Gauss term 0), then at the warm start
    compute_jar(VEC_SMOOTH, CTX_ARGS);
    double cost_sm = total_cost(VEC_SMOOTH, 0, CTX_ARGS);
    compute_jar(VEC_WARM, CTX_ARGS);
    double cost_ws = total_cost(VEC_WARM, 1, CTX_ARGS);   // full: it is the starting point in the common case
    bool use_smooth = cost_ws > cost_sm;
    LANE_FOR(i, NV) S.qacc[i] = use_smooth ? S.asmooth[i] : S.warm[i];
    WSYNC();
    cost = cost_ws;
    if (use_smooth) { compute_jar(VEC_QACC, CTX_ARGS); cost = total_cost(VEC_QACC, 1, CTX_ARGS); }
  }
  cta_any(S, true);
  if (active) newton_direction(0, 0.0, 0.0, CTX_ARGS);
  int iter = 0, nls = 0;
  TSG_UNROLL1
  for (;;) {
    bool go = active && iter < m.iterations;
    if (!cta_any(S, go)) break;
    double alpha = 0;
    if (go) {
      alpha = line_search(CTX_ARGS);
      nls += S.ls_evals;
      if (alpha == 0) active = false;
    }
    bool upd = go && active;
#ifndef TSG_ONE_BARRIER
    cta_any(S, true);
#endif
    if (upd) {
      WSYNC();
      LANE_FOR(i, NV + S.nact * 6) {
        if (i < NV) S.qacc[i] += alpha * S.search[i];
        else { int n = (i - NV) / 6, r = (i - NV) % 6; Con& k = con_at(S, S.order[n]); k.jar[r] += alpha * k.aref[r]; }
      }
      WSYNC();
      double oldcost = cost;
      cost = total_cost(VEC_QACC, 1, CTX_ARGS);
      iter++;
      if (newton_direction(1, oldcost, cost, CTX_ARGS) < 0) active = false;   // converged
    }
  }
  if (S.nact > 0) {
    WSYNC();
    LANE_FOR(i, NV) S.warm[i] = S.qacc[i];
    if (lane == 0) { S.niter_total += iter; S.nls_total += nls; }
    WSYNC();
  }
}

// ------------------------------------------------------------------ implicitfast + advance
TSG_FN void stage_integrate(Scratch& S, const DevModel& m, int lane) {
  double h = m.h;
  stage_tendon(S, m, lane);   // the tendon Jacobians shared their storage with the solver: recompute (same inputs, same values)
  stage_damping_blocks(S, m, lane);
  LANE_FOR(b, NBAR) {
    // (M - h D) x = qfrc_smooth + qfrc_constraint on the bar's 6x6 block, in place in shared memory
    double* A = S.u.post.Dblk[b];
    double* x = S.rhs + 6 * b;
    TSG_UNROLL1
    for (int t = 0; t < 21; t++) A[t] = -h * A[t];
    TSG_UNROLL1
    for (int r = 0; r < 6; r++) { A[r * (r + 1) / 2 + r] += m.M[6 * b + r]; x[r] = S.fsm[6 * b + r] + S.fcon[6 * b + r]; }
    TSG_UNROLL1
    for (int j = 0; j < 6; j++) {
      double* Aj = A + j * (j + 1) / 2;
      double s = Aj[j];
      TSG_UNROLL1
      for (int k = 0; k < j; k++) s -= Aj[k] * Aj[k];
      double inv = tsg_inv(tsg_sqrt(s));
      Aj[j] = inv;  // store 1 / L_jj
      TSG_UNROLL1
      for (int i = j + 1; i < 6; i++) {
        double* Ai = A + i * (i + 1) / 2;
        double t = Ai[j];
        TSG_UNROLL1
        for (int k = 0; k < j; k++) t -= Ai[k] * Aj[k];
        Ai[j] = t * inv;
      }
    }
    TSG_UNROLL1
    for (int i = 0; i < 6; i++) {
      double* Ai = A + i * (i + 1) / 2;
      double t = x[i];
      TSG_UNROLL1
      for (int k = 0; k < i; k++) t -= Ai[k] * x[k];
      x[i] = t * Ai[i];
    }
    TSG_UNROLL1
    for (int i = 5; i >= 0; i--) {
      double t = x[i];
      TSG_UNROLL1
      for (int k = i + 1; k < 6; k++) t -= A[k * (k + 1) / 2 + i] * x[k];
      x[i] = t * A[i * (i + 1) / 2 + i];
    }
    double* v = S.qvel + 6 * b; double* q = S.qpos + 7 * b;
    for (int k = 0; k < 6; k++) v[k] += h * x[k];
    for (int k = 0; k < 3; k++) q[k] += h * v[k];
    double ax[3] = {v[3], v[4], v[5]}, qr[4], qn[4];
    double ang = h * normalize3(ax);
    if (ang == 0) { qr[0] = 1; qr[1] = qr[2] = qr[3] = 0; }
    else { double sn, cs; sincos(ang * 0.5, &sn, &cs); qr[0] = cs; qr[1] = ax[0] * sn; qr[2] = ax[1] * sn; qr[3] = ax[2] * sn; }
    normalize4(q + 3);
    const double* a = q + 3;
    qn[0] = a[0] * qr[0] - a[1] * qr[1] - a[2] * qr[2] - a[3] * qr[3];
    qn[1] = a[0] * qr[1] + a[1] * qr[0] + a[2] * qr[3] - a[3] * qr[2];
    qn[2] = a[0] * qr[2] - a[1] * qr[3] + a[2] * qr[0] + a[3] * qr[1];
    qn[3] = a[0] * qr[3] + a[1] * qr[2] - a[2] * qr[1] + a[3] * qr[0];
    q[3] = qn[0]; q[4] = qn[1]; q[5] = qn[2]; q[6] = qn[3];
  }
  if (m.dyntype) { LANE_FOR(i, NACT) S.act[i] += h * S.actdot[i]; }
  WSYNC();
}

TSG_FN void reset_data(Scratch& S, const DevModel& m, int lane) {  // mj_resetData
  LANE_FOR(i, NQ) S.qpos[i] = m.qpos0[i];
  LANE_FOR(i, NV) { S.qvel[i] = 0; S.warm[i] = 0; }
  LANE_FOR(i, NACT) { S.ctrl[i] = 0; S.act[i] = 0; }
  WSYNC();
}
// mj_forward
TSG_FN_NOINLINE void forward(CTX_PARAMS) {
  CTX_BIND
  stage_position(S, m, lane);
  stage_tendon(S, m, lane);
  stage_smooth(S, m, lane);
  cta_any(S, true);
  stage_constraint(S, m, lane);
  stage_solve(S, m, c, lane);
}
// one mj_step
TSG_FN_NOINLINE void substep(CTX_PARAMS) {
  CTX_BIND
  // mj_checkPos / mj_checkVel
  bool bad = false;
  RED_FOR(i, NQ + NV) bad |= is_bad(i < NQ ? S.qpos[i] : S.qvel[i - NQ]);
  bad = RED_ANY(bad);
  WSYNC();
  if (bad) { if (lane == 0) S.bad |= 1; reset_data(S, m, lane); }
  forward(CTX_ARGS);
  bad = false;
  RED_FOR(i, NV) bad |= is_bad(S.qacc[i]);  // mj_checkAcc
  bad = RED_ANY(bad);
  WSYNC();
  if (bad) {  // the repeated forward pass runs unaligned: its CTA barriers have no partners
    int al = S.align;
    if (lane == 0) { S.bad |= 4; S.align = 0; }
    reset_data(S, m, lane); forward(CTX_ARGS);
    if (lane == 0) S.align = al;
    WSYNC();
  }
  stage_integrate(S, m, lane);
}

#if TSG_ALIGNED
// barrier protocol of one aligned substep for a (virtual) warp that has no env (tail of the batch): must mirror
// simulate() -> forward() -> stage_solve() exactly
// (the SAME barrier primitive as the working warps at every point: a plain and a reducing barrier must not meet)
TSG_FN void aligned_idle_substep() {
  align_sync();            // simulate: substep start
#ifdef TSG_ALIGN_SUBSTEP_ONLY
  return;
#endif
  align_any(true);         // forward: before stage_constraint
  align_any(true);         // stage_solve: before the warm-start costs
  align_any(true);         // stage_solve: before the first Newton direction
  for (;;) {
    if (!align_any(false)) break;
#ifndef TSG_ONE_BARRIER
    align_any(true);
#endif
  }
}
#endif

// mj_rnePostConstraint: cfrc_ext rows [torque; force] for world + 3 bars, from the last forward pass
TSG_FN void stage_cfrc(Scratch& S, const DevModel& m, int lane) {
  LANE_FOR(i, 24) {
    int body = i / 6, comp = i % 6;  // body 0 = world
    double com[3];
    if (body == 0) {
      double mt = 0; com[0] = com[1] = com[2] = 0;
      for (int b = 0; b < NBAR; b++) { addscl3(com, S.xstale + 3 * b, m.M[6 * b]); mt += m.M[6 * b]; }
      scl3(com, com, 1 / mt);
    } else copy3(com, S.xstale + 3 * (body - 1));
    double acc = 0;
    for (int n = 0; n < S.nact; n++) {
      const Con& c = con_at(S, S.order[n]);
      double s;
      if (c.b2 == body - 1) s = 1; else if (c.b1 == body - 1) s = -1; else continue;
      double F[3], T[3];
      mulMTV(F, c.frame, c.force); mulMTV(T, c.frame, c.force + 3);
      if (comp >= 3) acc += s * F[comp - 3];
      else { double r[3], tq[3]; sub3(r, c.pos, com); cross3(tq, r, F); acc += s * (tq[comp] + T[comp]); }
    }
    S.u.post.cfrc[body][comp] = acc;
  }
  WSYNC();
}

}  // namespace tsg
