// tsg_env.cuh -- tr_env / tensegrity_env semantics around the physics core, per warp.
//   step    : tr_env.py:327-527 ; tensegrity_env.py:291-410
//   obs     : tr_env.py:529-646 ; tensegrity_env.py:412-430
//   reset   : tr_env.py:709-872 ; tensegrity_env.py:433-512  (+ gym MujocoEnv.reset / set_state)
#pragma once
#include "tsg_core.cuh"

namespace tsg {

constexpr double PI = 3.14159265358979323846;
enum { ENV_TR = 0, ENV_LEGACY = 1 };
enum { TASK_STRAIGHT = 0, TASK_TURN = 1, TASK_AIMING = 2, TASK_TRACKING = 3, TASK_VEL_TRACK = 4 };

struct Aux {  // warp-uniform env bookkeeping kept in registers
  double xy_prev[2], psi_prev, reset_psi, waypt[2], ori[2];
  double step_num, ep_ret, ep_len, xvel, yvel;
  int head_n, head_pos;
  double* heading;  // [HEADING_SLOTS] in global memory
};
struct StepOut {  // warp-uniform results of one env step
  double reward, fwd, ctrl_cost, healthy, psi, xy[2];
  int terminated;
  double maxcfrc, barforce;
};

TSG_FN double angle_normalize(double t) {  // tr_env.py:648-654
  while (t > PI) t -= 2 * PI;
  while (t <= -PI) t += 2 * PI;
  return t;
}

struct Pose { double xy[2], left[3], right[3], psi; };
// COM / left-right end-cap centroids from the (stale) kinematics of the last forward pass
TSG_FN void read_pose(const Scratch& S, Pose& P) {
  P.xy[0] = (S.xstale[0] + S.xstale[3] + S.xstale[6]) / 3;
  P.xy[1] = (S.xstale[1] + S.xstale[4] + S.xstale[7]) / 3;
  for (int k = 0; k < 3; k++) {
    P.left[k] = (S.sph[3 * 0 + k] + S.sph[3 * 2 + k] + S.sph[3 * 4 + k]) / 3;   // s0, s2, s4
    P.right[k] = (S.sph[3 * 1 + k] + S.sph[3 * 3 + k] + S.sph[3 * 5 + k]) / 3;  // s1, s3, s5
  }
  P.psi = atan2(-(P.left[0] - P.right[0]), P.left[1] - P.right[1]);
}

TSG_FN double ditch_reward(const EnvCfg& c, const Aux& A, const double* xy) {  // tr_env.py:656-667
  double pv[2] = {A.waypt[0] - A.ori[0], A.waypt[1] - A.ori[1]};
  double dp = tsg_sqrt(pv[0] * pv[0] + pv[1] * pv[1]);
  double pn[2] = {pv[0] / dp, pv[1] / dp};
  double tv[2] = {A.waypt[0] - xy[0], A.waypt[1] - xy[1]};
  double along = tv[0] * pn[0] + tv[1] * pn[1];
  double bx = tv[0] - along * pn[0], by = tv[1] - along * pn[1];
  double bias = tsg_sqrt(bx * bx + by * by);
  double ditch = c.ditch_reward_max * (1.0 - fabs(along) / dp) * exp(-(bias * bias) / (2 * c.ditch_reward_stdev * c.ditch_reward_stdev));
  double dx = xy[0] - A.waypt[0], dy = xy[1] - A.waypt[1];
  double dn = tsg_sqrt(dx * dx + dy * dy);
  double wp = c.waypt_reward_amplitude * exp(-(dn * dn) / (2 * c.waypt_reward_stdev * c.waypt_reward_stdev));
  return ditch + wp;
}

// scipy Rotation.from_matrix(M).as_quat() -> (x, y, z, w)
TSG_FN void mat2quat_scipy(const double* M, double* q) {
  double tr = M[0] + M[4] + M[8];
  double dec[4] = {M[0], M[4], M[8], tr};
  int ch = 0;
  for (int i = 1; i < 4; i++) if (dec[i] > dec[ch]) ch = i;
  if (ch != 3) {
    int i = ch, j = (i + 1) % 3, k = (j + 1) % 3;
    q[i] = 1 - tr + 2 * M[4 * i];
    q[j] = M[3 * j + i] + M[3 * i + j];
    q[k] = M[3 * k + i] + M[3 * i + k];
    q[3] = M[3 * k + j] - M[3 * j + k];
  } else {
    q[0] = M[7] - M[5]; q[1] = M[2] - M[6]; q[2] = M[3] - M[1]; q[3] = 1 + tr;
  }
  double n = tsg_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; i++) q[i] /= n;
}

// Philox4x32-10, counter = (env id, reset count), key = seed
TSG_FN void philox(unsigned long long seed, unsigned long long ctr_lo, unsigned long long ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = (uint32_t)ctr_hi, c3 = (uint32_t)(ctr_hi >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
TSG_FN double u01(uint32_t a, uint32_t b) {  // 53-bit uniform in [0,1)
  unsigned long long x = (((unsigned long long)a << 32) | b) >> 11;
  return (double)x * (1.0 / 9007199254740992.0);
}
// Observation noise (tr_env.py:552-644).  S.u.post.obs holds the true observation on entry and the noisy one on
// return: every component of the cap positions / cap velocities gets N(0, obs_noise_cap_pos_stdev), every tendon
// length N(0, obs_noise_tendon_stdev); the tracking vector loses the mean cap-position noise and the target yaw is
// re-derived from it (:626-639); the vel_track command is passed through (:641-644).  Normals: Philox keyed by
// (seed, stream id, reset count, episode step, pair index) + Box-Muller, one pair per lane; they are staged in the
// implicit-damping scratch, which is dead at this point.
// the pr-th pair of standard normals of one observation (Philox + Box-Muller)
TSG_FN void noise_pair(unsigned long long seed, unsigned long long stream, unsigned long long nreset,
                       unsigned long long step, int pr, double& z0, double& z1) {
  uint32_t r[4];
  philox(seed ^ 0x6f62736e6f697365ull, stream, (((nreset << 24) | (step & 0xffffffull)) << 8) | (unsigned long long)pr, r);
  double u1 = 1.0 - u01(r[0], r[1]), u2 = u01(r[2], r[3]);
  double rad = sqrt(-2.0 * log(u1));
  z0 = rad * cos(2 * PI * u2); z1 = rad * sin(2 * PI * u2);
}
TSG_FN_NOINLINE void obs_noise(unsigned long long seed, unsigned long long stream, unsigned long long nreset,
                               unsigned long long step, CTX_PARAMS) {
  CTX_BIND
  const int nvel = c.use_cap_velocity ? 18 : 0, n = 27 + nvel;
  double* z = &S.u.post.Dblk[0][0];
  LANE_FOR(pr, (n + 1) / 2) {
    double z0, z1;
    noise_pair(seed, stream, nreset, step, pr, z0, z1);
    z[2 * pr] = z0;
    if (2 * pr + 1 < n) z[2 * pr + 1] = z1;
  }
  WSYNC();
  const double sp = c.obs_noise_cap_pos_stdev, st = c.obs_noise_tendon_stdev;
  double cn0 = 0, cn1 = 0;   // mean of the noisy centroid-relative cap positions (x, y)
  for (int cap = 0; cap < 6; cap++) {
    cn0 += sp * z[3 * cap] + S.u.post.obs[3 * cap];
    cn1 += sp * z[3 * cap + 1] + S.u.post.obs[3 * cap + 1];
  }
  cn0 /= 6; cn1 /= 6;
  WSYNC();
  LANE_FOR(i, n) S.u.post.obs[i] = (i < 18 + nvel ? sp : st) * z[i] + S.u.post.obs[i];
  if ((c.task == TASK_TRACKING || c.task == TASK_AIMING) && lane == 0) {
    double tx = S.u.post.obs[n] - cn0, ty = S.u.post.obs[n + 1] - cn1, nn = sqrt(tx * tx + ty * ty);
    S.u.post.obs[n] = tx; S.u.post.obs[n + 1] = ty; S.u.post.obs[n + 2] = atan2(ty / nn, tx / nn);
  }
  WSYNC();
}
// true observation -> io.real_obs row (if registered), then the noise
#define TSG_OBS_NOISE(row, stream, nreset, step)                                                                     \
  if (c.use_obs_noise) {                                                                                              \
    if (io.real_obs && (row) >= 0) { LANE_FOR(i_, c.obs_dim) io.real_obs[(size_t)(row) * c.obs_dim + i_] = S.u.post.obs[i_]; } \
    WSYNC();                                                                                                          \
    obs_noise(io.seed, (unsigned long long)(stream), (unsigned long long)(nreset), (unsigned long long)(step), CTX_ARGS); \
  }

// observation into S.obs (stale positions / tendon lengths, fresh qvel)
TSG_FN void compute_obs(EnvScratch& S, const DevModel& m, const EnvCfg& c, const Aux& A, int lane) {
  if (c.env_kind == ENV_LEGACY) {
    LANE_FOR(i, 39) {
      if (i < 12) {
        if (i % 4 == 0) {  // geom rXY: body frame with x, y columns negated (geom quat 0 0 0 1)
          int b = i / 4; const double* R = S.xmat + 9 * b;
          double G[9] = {-R[0], -R[1], R[2], -R[3], -R[4], R[5], -R[6], -R[7], R[8]}, q[4];
          mat2quat_scipy(G, q);
          for (int k = 0; k < 4; k++) S.u.post.obs[i + k] = q[k];
        }
      } else if (i < 30) S.u.post.obs[i] = S.qvel[i - 12];
      else S.u.post.obs[i] = S.tlen[i - 30];
    }
    WSYNC();
    return;
  }
  double cen[3] = {0, 0, 0};
  for (int b = 0; b < NBAR; b++)
    for (int k = 0; k < 3; k++) cen[k] += S.sph[3 * (2 * b) + k] + S.sph[3 * (2 * b + 1) + k];
  for (int k = 0; k < 3; k++) cen[k] /= 6;
  int nvel = c.use_cap_velocity ? 18 : 0;
  LANE_FOR(i, 18 + nvel + 9) {
    if (i < 18) { int cap = i / 3, k = i % 3; S.u.post.obs[i] = S.sph[3 * cap + k] - cen[k]; }
    else if (i < 18 + nvel) {
      int j = i - 18, cap = j / 3, k = j % 3, b = cap / 2;
      double r[3], w[3] = {S.qvel[6 * b + 3], S.qvel[6 * b + 4], S.qvel[6 * b + 5]}, cr[3];
      sub3(r, S.sph + 3 * cap, S.xstale + 3 * b);
      cross3(cr, w, r);  // local-frame angular velocity used as if world-frame (tr_env.py:599-604)
      S.u.post.obs[i] = S.qvel[6 * b + k] + cr[k];
    } else S.u.post.obs[i] = S.tlen[i - 18 - nvel];
  }
  int base = 27 + nvel;
  if (c.task == TASK_TRACKING || c.task == TASK_AIMING) {
    double tx = A.waypt[0] - cen[0], ty = A.waypt[1] - cen[1], n = tsg_sqrt(tx * tx + ty * ty);
    if (lane == 0) { S.u.post.obs[base] = tx; S.u.post.obs[base + 1] = ty; S.u.post.obs[base + 2] = atan2(ty / n, tx / n); }
  } else if (c.task == TASK_VEL_TRACK) {
    if (lane == 0) { S.u.post.obs[base] = 0.5 * cos(A.reset_psi); S.u.post.obs[base + 1] = 0.5 * sin(A.reset_psi); S.u.post.obs[base + 2] = 0.0; }
  }
  WSYNC();
}

// do_simulation(ctrl, frame_skip): ctrl already in S.ctrl
TSG_FN_NOINLINE void simulate(CTX_PARAMS) {
  CTX_BIND
  for (int s = 0; s < c.frame_skip; s++) {
#if TSG_ALIGNED
    if (S.align) align_sync();
#endif
    substep(CTX_ARGS);
  }
  stage_cfrc(S, m, lane);
}

// heading ring buffer (the reference's deque, never cleared across episodes) lives in global memory
TSG_FN void heading_push(Aux& A, double v, int lane) {
  if (lane == 0) A.heading[(A.head_pos + A.head_n) % HEADING_SLOTS] = v;
  A.head_n++;
  WSYNC();
}
TSG_FN double heading_pop(Aux& A) {
  double v = A.heading[A.head_pos];
  A.head_pos = (A.head_pos + 1) % HEADING_SLOTS;
  A.head_n--;
  return v;
}

// one env.step(action) with action in S.action; updates S, A; obs NOT computed here
TSG_FN_NOINLINE void env_step(Aux& A, StepOut& O, CTX_PARAMS) {
  CTX_BIND
  double dt = c.dt;
  double xy_before[2] = {A.xy_prev[0], A.xy_prev[1]}, psi_before = A.psi_prev;
  if (c.env_kind == ENV_TR) {  // _action_filter, k_FILTER = 1 (tr_env.py:680-683)
    LANE_FOR(i, NACT) S.ctrl[i] = S.ctrl[i] + 1.0 * (S.action[i] - S.ctrl[i]) * dt;
  } else {
    LANE_FOR(i, NACT) S.ctrl[i] = S.action[i];
  }
  WSYNC();
  simulate(CTX_ARGS);
  Pose P; read_pose(S, P);
  double xvel = (P.xy[0] - xy_before[0]) / dt, yvel = (P.xy[1] - xy_before[1]) / dt;
  A.xvel = xvel; A.yvel = yvel;
  double psi_after = P.psi;
  if (c.env_kind == ENV_LEGACY && c.task == TASK_TURN)  // tensegrity_env.py:320-322
    psi_after = atan2(P.right[1] - P.left[1], P.right[0] - P.left[0]);
  double psi_info = psi_after;
  // control cost
  double cc = 0;
  for (int i = 0; i < NACT; i++) {
    double a = S.action[i];
    double v = (c.env_kind == ENV_TR) ? (a + 0.5 - S.tlen[i]) : a;
    cc += v * v;
  }
  cc *= c.ctrl_cost_weight;
  double fwd = 0, ctrl_cost = cc;
  double healthy = c.terminate_when_unhealthy ? c.healthy_reward : 0.0;
  int delay = c.reward_delay_steps;
  bool finite = true;
  for (int i = 0; i < NQ; i++) finite &= isfinite(S.qpos[i]);
  for (int i = 0; i < NV; i++) finite &= isfinite(S.qvel[i]);
  bool moving_any = false;
  for (int i = 0; i < NV; i++) moving_any |= fabs(S.qvel[i]) > 0.1;
  bool healthy_turn = finite && moving_any;
  bool healthy_lin = finite && ((xvel > 1e-4 || xvel < -1e-4) || (yvel > 1e-4 || yvel < -1e-4));
  bool is_healthy = healthy_lin;
  bool extra_term = false;
  if (c.task == TASK_TURN) {
    is_healthy = healthy_turn;
    heading_push(A, psi_after, lane);
    if (A.head_n > delay) {
      double old_psi = heading_pop(A), pa = psi_after;
      if (pa < -PI / 2 && old_psi > PI / 2) pa = 2 * PI + pa;
      else if (pa > PI / 2 && old_psi < -PI / 2) pa = -2 * PI + pa;
      if (c.env_kind == ENV_TR) psi_info = pa;  // tr_env rebinds psi_after, legacy too
      else psi_info = pa;
      fwd = (pa - old_psi) / (dt * delay) * c.desired_direction;
    } else { fwd = 0; ctrl_cost = 0; }
  } else if (c.task == TASK_STRAIGHT) {
    double dx = P.xy[0] - xy_before[0], dy = P.xy[1] - xy_before[1];
    double psi_diff = fabs(atan2(dy, dx) - A.reset_psi);
    fwd = c.desired_direction * (tsg_sqrt(dx * dx + dy * dy) * cos(psi_diff) / dt);
  } else if (c.task == TASK_AIMING) {
    is_healthy = healthy_turn;
    double tx = A.waypt[0] - xy_before[0], ty = A.waypt[1] - xy_before[1], n = tsg_sqrt(tx * tx + ty * ty);
    double target_psi = atan2(ty / n, tx / n);
    double newp = angle_normalize(target_psi - psi_after);
    heading_push(A, newp, lane);
    if (A.head_n > delay) {
      double oldp = heading_pop(A);
      fwd = -(fabs(newp) - fabs(oldp)) / (dt * delay) * c.yaw_reward_weight;
    }
    healthy = 0;
    extra_term = A.step_num > 1000;
  } else if (c.task == TASK_TRACKING) {
    fwd = ditch_reward(c, A, P.xy) - ditch_reward(c, A, xy_before);
    healthy = 0;
    extra_term = A.step_num > 1000;
  } else {  // vel_track, tr_env.py:461-474, 669-678
    double ang = angle_normalize(psi_after - psi_before) / dt;
    double cx = 0.5 * cos(A.reset_psi), cy = 0.5 * sin(A.reset_psi);
    double le = tsg_sqrt((xvel - cx) * (xvel - cx) + (yvel - cy) * (yvel - cy)), ae = ang - 0.0;
    fwd = 1.0 * exp(-5.0 * le * le) + 0.5 * exp(-7.0 * ae * ae);
  }
  bool terminated = c.terminate_when_unhealthy ? !is_healthy : false;
  if (extra_term) terminated = true;
  double maxc = 0;
  for (int i = 0; i < 24; i++) maxc = fmax(maxc, fabs(S.u.post.cfrc[i / 6][i % 6]));
  if (maxc > c.kill_force) terminated = true;  // tr_env.py:480-481
  double barf = 0;  // run.py:155-161 total bar-bar contact force magnitude
  for (int n = 0; n < S.nact; n++) {
    const Con& k = con_at(S, S.order[n]);
    if (k.b1 >= 0) barf += tsg_sqrt(k.force[0] * k.force[0] + k.force[1] * k.force[1] + k.force[2] * k.force[2]);
  }
  O.reward = fwd + healthy - ctrl_cost;
  O.fwd = fwd; O.ctrl_cost = ctrl_cost; O.healthy = healthy; O.psi = psi_info;
  O.xy[0] = P.xy[0]; O.xy[1] = P.xy[1];
  O.terminated = terminated ? 1 : 0; O.maxcfrc = maxc; O.barforce = barf;
  A.step_num += 1;
  A.xy_prev[0] = P.xy[0]; A.xy_prev[1] = P.xy[1]; A.psi_prev = P.psi;
}

// refresh the "stale" pose bookkeeping after a forward pass (set_state)
TSG_FN void aux_from_forward(const Scratch& S, Aux& A) {
  Pose P; read_pose(S, P);
  A.xy_prev[0] = P.xy[0]; A.xy_prev[1] = P.xy[1]; A.psi_prev = P.psi;
}

// env.reset() = MujocoEnv.reset (mj_resetData) + reset_model, split in three pieces so that it can run either in
// one go (tsg_reset / fallback) or one warm-up step per launch on a background pool slot.  Random draws in S.draws.
TSG_FN void reset_setpoints(EnvScratch& S, const EnvCfg& c, int lane) {
  const double* u = S.draws;
  LANE_FOR(i, NACT) {
    double t = u[2 + i] * c.tendon_reset_stdev + c.tendon_reset_mean;
    if (t > c.tendon_max_length) t = c.tendon_max_length; else if (t < c.tendon_min_length) t = c.tendon_min_length;
    S.action[i] = t;
  }
  WSYNC();
}
TSG_FN void reset_begin(EnvScratch& S, const DevModel& m, const EnvCfg& c, Aux& A, int lane) {
  const double* u = S.draws;
  reset_data(S, m, lane);
  int idx = (int)floor(u[0] * c.npose);
  if (idx > c.npose - 1) idx = c.npose - 1;
  if (idx < 0) idx = 0;
  bool extra_set_state = (c.env_kind == ENV_TR) ? (c.task == TASK_TURN || c.task == TASK_TRACKING || c.task == TASK_AIMING)
                                                 : (c.task == TASK_TURN);
  int nfwd = (c.env_kind == ENV_TR ? 1 : 0) + (extra_set_state ? 1 : 0);
  LANE_FOR(i, NQ) S.qpos[i] = c.reset_pose[idx][i];
  WSYNC();
  for (int k = 0; k < nfwd; k++) forward(CTX_ARGS);  // set_state -> mj_forward
  // rotate the whole robot about world z by theta (positions and orientations)
  double theta = c.min_reset_heading + u[1] * (c.max_reset_heading - c.min_reset_heading);
  double ct = cos(theta), st = sin(theta), ch = cos(0.5 * theta), sh = sin(0.5 * theta);
  LANE_FOR(b, NBAR) {
    // start again from the table pose: mj_kinematics normalised qpos in place, the reference re-uses its own copy
    double p[7];
    for (int k = 0; k < 7; k++) p[k] = c.reset_pose[idx][7 * b + k];
    double* q = S.qpos + 7 * b;
    q[0] = ct * p[0] - st * p[1]; q[1] = st * p[0] + ct * p[1]; q[2] = p[2];
    double n = tsg_sqrt(p[3] * p[3] + p[4] * p[4] + p[5] * p[5] + p[6] * p[6]);
    double w = p[3] / n, x = p[4] / n, y = p[5] / n, z = p[6] / n;
    q[3] = ch * w - sh * z; q[4] = ch * x - sh * y; q[5] = ch * y + sh * x; q[6] = ch * z + sh * w;  // q_z(theta) * q
  }
  WSYNC();
  forward(CTX_ARGS);
  aux_from_forward(S, A);
  reset_setpoints(S, c, lane);
  if (c.env_kind == ENV_TR) { LANE_FOR(i, NACT) S.ctrl[i] = S.action[i]; WSYNC(); }
}
// one of the warmup_steps settling steps at the set-points in S.action
TSG_FN void reset_warm_step(EnvScratch& S, const DevModel& m, const EnvCfg& c, Aux& A, int lane) {
  if (c.env_kind == ENV_TR) { simulate(CTX_ARGS); aux_from_forward(S, A); }  // do_simulation, no filter
  else { StepOut O; env_step(A, O, CTX_ARGS); }                              // full self.step
}
TSG_FN void reset_finish(EnvScratch& S, const DevModel& m, const EnvCfg& c, Aux& A, int lane) {
  const double* u = S.draws;
  Pose P; read_pose(S, P);
  A.reset_psi = P.psi;
  double lo = c.waypt_range[0], hi = c.waypt_range[1];
  if (c.env_kind == ENV_TR && c.task == TASK_TRACKING) {
    A.ori[0] = (P.left[0] + P.right[0]) / 2; A.ori[1] = (P.left[1] + P.right[1]) / 2;
    double len = lo + u[8] * (hi - lo);
    double yaw = c.waypt_angle_range[0] + u[9] * (c.waypt_angle_range[1] - c.waypt_angle_range[0]) + A.reset_psi;
    if (c.is_test) { len = 0.5 * hi + 0.5 * lo; yaw = (0.5 * c.waypt_angle_range[1] + 0.5 * c.waypt_angle_range[0]) + A.reset_psi; }
    A.waypt[0] = A.ori[0] + len * cos(yaw); A.waypt[1] = A.ori[1] + len * sin(yaw);
  } else if (c.env_kind == ENV_TR && c.task == TASK_AIMING) {
    A.ori[0] = P.left[0] + P.right[0] / 2;  // operator-precedence quirk kept (tr_env.py:843)
    A.ori[1] = (P.left[1] + P.right[1]) / 2;
    double len = lo + u[8] * (hi - lo);
    double yaw = -PI + u[9] * (2 * PI) + A.reset_psi;
    if (c.is_test) { len = 0.5 * hi + 0.5 * lo; yaw = (0.75 * PI + 0.25 * (-PI)) + A.reset_psi; }
    A.waypt[0] = A.ori[0] + len * cos(yaw); A.waypt[1] = A.ori[1] + len * sin(yaw);
    if (c.is_test) { A.waypt[0] = 0; A.waypt[1] = 0; }
  }
  A.step_num = 0;
  if (c.env_kind == ENV_TR && (c.task == TASK_TURN || c.task == TASK_AIMING)) {
    StepOut O;
    for (int k = 0; k < c.reward_delay_steps; k++) env_step(A, O, CTX_ARGS);
  }
  A.ep_ret = 0; A.ep_len = 0;
}
TSG_FN void env_reset(EnvScratch& S, const DevModel& m, const EnvCfg& c, Aux& A, int lane) {
  reset_begin(S, m, c, A, lane);
  for (int k = 0; k < c.warmup_steps; k++) reset_warm_step(S, m, c, A, lane);
  reset_finish(S, m, c, A, lane);
}

// ---- state record <-> scratch
TSG_FN void load_env(EnvScratch& S, Aux& A, const double* rec, const double* head, bool need_head, int lane) {
  LANE_FOR(i, SO_XY_PREV) {
    double v = rec[i];
    if (i < SO_QVEL) S.qpos[i] = v;
    else if (i < SO_WARM) S.qvel[i - SO_QVEL] = v;
    else if (i < SO_CTRL) S.warm[i - SO_WARM] = v;
    else if (i < SO_ACT) S.ctrl[i - SO_CTRL] = v;
    else S.act[i - SO_ACT] = v;
  }
  A.heading = const_cast<double*>(head);
  A.xy_prev[0] = rec[SO_XY_PREV]; A.xy_prev[1] = rec[SO_XY_PREV + 1]; A.psi_prev = rec[SO_PSI_PREV];
  A.reset_psi = rec[SO_RESET_PSI]; A.waypt[0] = rec[SO_WAYPT]; A.waypt[1] = rec[SO_WAYPT + 1];
  A.ori[0] = rec[SO_ORI]; A.ori[1] = rec[SO_ORI + 1];
  A.step_num = rec[SO_STEP_NUM]; A.ep_ret = rec[SO_EP_RET]; A.ep_len = rec[SO_EP_LEN];
  A.xvel = rec[SO_XVEL]; A.yvel = rec[SO_YVEL];
  A.head_n = (int)rec[SO_HEAD_N]; A.head_pos = (int)rec[SO_HEAD_POS];
  if (lane == 0) { S.overflow = 0; S.bad = 0; S.niter_total = 0; S.nls_total = 0; S.nmpr_total = 0; S.nact = 0; }
  WSYNC();
}
TSG_FN void store_env(const EnvScratch& S, const Aux& A, double* rec, double* head, bool need_head, int lane) {
  LANE_FOR(i, SO_XY_PREV) {
    double v;
    if (i < SO_QVEL) v = S.qpos[i];
    else if (i < SO_WARM) v = S.qvel[i - SO_QVEL];
    else if (i < SO_CTRL) v = S.warm[i - SO_WARM];
    else if (i < SO_ACT) v = S.ctrl[i - SO_CTRL];
    else v = S.act[i - SO_ACT];
    rec[i] = v;
  }
  if (lane == 0) {
    rec[SO_XY_PREV] = A.xy_prev[0]; rec[SO_XY_PREV + 1] = A.xy_prev[1]; rec[SO_PSI_PREV] = A.psi_prev;
    rec[SO_RESET_PSI] = A.reset_psi; rec[SO_WAYPT] = A.waypt[0]; rec[SO_WAYPT + 1] = A.waypt[1];
    rec[SO_ORI] = A.ori[0]; rec[SO_ORI + 1] = A.ori[1];
    rec[SO_STEP_NUM] = A.step_num; rec[SO_EP_RET] = A.ep_ret; rec[SO_EP_LEN] = A.ep_len;
    rec[SO_XVEL] = A.xvel; rec[SO_YVEL] = A.yvel;
    rec[SO_HEAD_N] = (double)A.head_n; rec[SO_HEAD_POS] = (double)A.head_pos;
  }
}

struct StepIO {
  double* state;        // [N][STATE_STRIDE]
  double* heading;      // [N][HEADING_SLOTS]
  const double* ctrl64; // [N][6] or null
  const float* ctrl32;  // [N][6] or null
  double* obs;          // [N][obs_dim] or null
  float* obs32;         // [N][obs_dim] or null
  double* reward;       // [N] or null
  uint8_t* done;        // [N] or null (terminated | truncated)
  double* info;         // [N][INFO_DIM] or null
  double* term_obs;     // [N][obs_dim] or null : observation before an auto reset
  double* draws;        // [N][NDRAW]: reset draws (in: explicit, out: generated)
  const uint8_t* mask;  // reset: which envs ; null = all
  unsigned long long seed;
  long long env_id_base;
  int n_envs;
  int explicit_draws;
  int n_pool;           // background reset pool slots stored after the n_envs records
  double* pool_obs;     // [n_pool][obs_dim]: reset observation of each ready pool slot
  double* real_obs;     // [N][obs_dim] or null: noise-free observation when use_obs_noise (info["real_observation"])
};

TSG_FN void write_obs(const EnvScratch& S, const EnvCfg& c, const StepIO& io, int e, int lane) {
  if (io.obs) { LANE_FOR(i, c.obs_dim) io.obs[(size_t)e * c.obs_dim + i] = S.u.post.obs[i]; }
  if (io.obs32) { LANE_FOR(i, c.obs_dim) io.obs32[(size_t)e * c.obs_dim + i] = (float)S.u.post.obs[i]; }
}

// the body of the step kernel for env e
TSG_FN void run_step(EnvScratch& S, const DevModel& m, const EnvCfg& c, const StepIO& io, int e, int lane) {
  Aux A; StepOut O;
  bool need_head = (c.task == TASK_TURN || c.task == TASK_AIMING);
  load_env(S, A, io.state + (size_t)e * STATE_STRIDE, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
  LANE_FOR(i, NACT) S.action[i] = io.ctrl64 ? io.ctrl64[(size_t)e * NACT + i] : (double)io.ctrl32[(size_t)e * NACT + i];
  WSYNC();
  env_step(A, O, CTX_ARGS);
  compute_obs(S, m, c, A, lane);
  A.ep_len += 1; A.ep_ret += O.reward;
  TSG_OBS_NOISE(e, io.env_id_base + e, io.state[(size_t)e * STATE_STRIDE + SO_NRESET], A.ep_len)
  int truncated = (c.max_episode_steps > 0 && A.ep_len >= c.max_episode_steps) ? 1 : 0;
  write_obs(S, c, io, e, lane);
  if (lane == 0) {
    if (io.reward) io.reward[e] = O.reward;
    if (io.done) io.done[e] = (uint8_t)((O.terminated || truncated) ? 1 : 0);
  }
  if (io.info) {
    double* I = io.info + (size_t)e * INFO_DIM;
    LANE_FOR(i, INFO_DIM) {
      double v = 0;
      switch (i) {
        case IO_REW_FWD: v = O.fwd; break;
        case IO_REW_CTRL: v = -O.ctrl_cost; break;
        case IO_REW_SURVIVE: v = O.healthy; break;
        case IO_X: v = O.xy[0]; break;
        case IO_Y: v = O.xy[1]; break;
        case IO_PSI: v = O.psi; break;
        case IO_XVEL: v = A.xvel; break;
        case IO_YVEL: v = A.yvel; break;
        case IO_TERMINATED: v = O.terminated; break;
        case IO_TRUNCATED: v = truncated; break;
        case IO_NCON: v = S.nact; break;
        case IO_NITER: v = S.niter_total; break;
        case IO_NLS: v = S.nls_total; break;
        case IO_BARFORCE: v = O.barforce; break;
        case IO_MAXCFRC: v = O.maxcfrc; break;
        case IO_WAYPT: v = A.waypt[0]; break;
        case IO_WAYPT + 1: v = A.waypt[1]; break;
        case IO_ORI: v = A.ori[0]; break;
        case IO_ORI + 1: v = A.ori[1]; break;
        case IO_OVERFLOW: v = S.overflow; break;
        case IO_BAD: v = S.bad; break;
        case IO_NMPR: v = S.nmpr_total; break;
        case IO_RESET_PSI: v = A.reset_psi; break;
        default: if (i >= IO_TEN && i < IO_TEN + 9) v = S.tlen[i - IO_TEN];
      }
      I[i] = v;
    }
  }
  store_env(S, A, io.state + (size_t)e * STATE_STRIDE, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
}

TSG_FN void make_draws(double* d, unsigned long long seed, unsigned long long env_id, unsigned long long nreset) {
  double un[12];
  for (int k = 0; k < 6; k++) {
    uint32_t r[4];
    philox(seed, env_id, (nreset << 8) | (unsigned long long)k, r);
    un[2 * k] = u01(r[0], r[1]); un[2 * k + 1] = u01(r[2], r[3]);
  }
  d[0] = un[0]; d[1] = un[1]; d[8] = un[2]; d[9] = un[3];
  for (int k = 0; k < 3; k++) {  // Box-Muller
    double u1 = 1.0 - un[4 + 2 * k], u2 = un[5 + 2 * k];
    double r = tsg_sqrt(-2.0 * log(u1));
    d[2 + 2 * k] = r * cos(2 * PI * u2); d[3 + 2 * k] = r * sin(2 * PI * u2);
  }
}

// the body of the reset kernel for env e (mask already checked)
TSG_FN void run_reset(EnvScratch& S, const DevModel& m, const EnvCfg& c, const StepIO& io, int e, int lane) {
  Aux A;
  bool need_head = (c.task == TASK_TURN || c.task == TASK_AIMING);
  double* rec = io.state + (size_t)e * STATE_STRIDE;
  load_env(S, A, rec, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
  if (io.term_obs && io.obs) { LANE_FOR(i, c.obs_dim) io.term_obs[(size_t)e * c.obs_dim + i] = io.obs[(size_t)e * c.obs_dim + i]; }
  double nreset = rec[SO_NRESET];
  if (io.explicit_draws) { LANE_FOR(i, NDRAW) S.draws[i] = io.draws[(size_t)e * NDRAW + i]; }
  else {
    if (lane == 0) make_draws(S.draws, io.seed, (unsigned long long)(io.env_id_base + e), (unsigned long long)nreset);
    WSYNC();
    if (io.draws) { LANE_FOR(i, NDRAW) io.draws[(size_t)e * NDRAW + i] = S.draws[i]; }
  }
  WSYNC();
  env_reset(S, m, c, A, lane);
  compute_obs(S, m, c, A, lane);
  TSG_OBS_NOISE(e, io.env_id_base + e, nreset + 1, 0)
  write_obs(S, c, io, e, lane);
  store_env(S, A, rec, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
  if (lane == 0) rec[SO_NRESET] = nreset + 1;
}

// background reset pool: slot p (record n_envs + p) advances by one warm-up step per launch until it holds a
// completely reset env (state, heading ring, reset observation); tsg_assign_kernel then hands ready slots to envs
// that are done, so an env reset costs no latency tail.  phase (SO_FLAGS) = warm-up steps done, warmup_steps+1 = ready.
TSG_FN void run_pool(EnvScratch& S, const DevModel& m, const EnvCfg& c, const StepIO& io, int p, bool finish_now, int lane) {
  Aux A;
  size_t row = (size_t)io.n_envs + p;
  double* rec = io.state + row * STATE_STRIDE;
  int phase = (int)rec[SO_FLAGS];
  if (phase > c.warmup_steps) return;
  load_env(S, A, rec, io.heading + row * HEADING_SLOTS, true, lane);
  double nreset = rec[SO_NRESET];
  if (phase == 0) {
    if (lane == 0) make_draws(S.draws, io.seed, (1ull << 40) + (unsigned long long)(io.env_id_base + p), (unsigned long long)nreset);
    WSYNC();
    LANE_FOR(i, NDRAW) io.draws[row * NDRAW + i] = S.draws[i];
  } else { LANE_FOR(i, NDRAW) S.draws[i] = io.draws[row * NDRAW + i]; }
  WSYNC();
  // only the single warm-up step of a launch follows the aligned barrier protocol (exactly one simulate());
  // the begin / finish parts contain extra forward passes and steps and run unaligned
  int aligned = S.align && !finish_now;
  WSYNC();
  if (lane == 0) S.align = 0;
  WSYNC();
  if (phase == 0) reset_begin(S, m, c, A, lane); else reset_setpoints(S, c, lane);
  do {
    if (lane == 0) S.align = aligned;
    WSYNC();
    reset_warm_step(S, m, c, A, lane);
    if (lane == 0) S.align = 0;
    WSYNC();
    phase++;
  } while (finish_now && phase < c.warmup_steps);
  if (phase >= c.warmup_steps) {
    reset_finish(S, m, c, A, lane);
    compute_obs(S, m, c, A, lane);
    TSG_OBS_NOISE(-1, (1ull << 40) + (unsigned long long)(io.env_id_base + p), nreset + 1, 0)
    LANE_FOR(i, c.obs_dim) io.pool_obs[(size_t)p * c.obs_dim + i] = S.u.post.obs[i];
    phase = c.warmup_steps + 1;
  }
  store_env(S, A, rec, io.heading + row * HEADING_SLOTS, true, lane);
  if (lane == 0) rec[SO_FLAGS] = (double)phase;
}

// mj_forward on the stored state (after tsg_set_state): refresh kinematics bookkeeping and obs
TSG_FN void run_forward(EnvScratch& S, const DevModel& m, const EnvCfg& c, const StepIO& io, int e, int lane) {
  Aux A;
  bool need_head = (c.task == TASK_TURN || c.task == TASK_AIMING);
  double* rec = io.state + (size_t)e * STATE_STRIDE;
  load_env(S, A, rec, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
  forward(CTX_ARGS);
  stage_cfrc(S, m, lane);
  aux_from_forward(S, A);
  compute_obs(S, m, c, A, lane);
  write_obs(S, c, io, e, lane);
  if (io.info) {
    double* I = io.info + (size_t)e * INFO_DIM;
    LANE_FOR(i, INFO_DIM) {
      double v = 0;
      if (i >= IO_TEN && i < IO_TEN + 9) v = S.tlen[i - IO_TEN];
      else if (i == IO_NCON) v = S.nact;
      else if (i == IO_X) v = A.xy_prev[0];
      else if (i == IO_Y) v = A.xy_prev[1];
      else if (i == IO_PSI) v = A.psi_prev;
      I[i] = v;
    }
  }
  store_env(S, A, rec, io.heading + (size_t)e * HEADING_SLOTS, need_head, lane);
}

}  // namespace tsg
