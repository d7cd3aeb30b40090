// tsg_host.h -- host-side preparation of the device constants from the C-ABI structs.
#pragma once
#include <math.h>
#include <string.h>

#include <string>

#include "../../include/tsg_model.h"
#include "tsg_env.cuh"

namespace tsg {

// returns "" on success, else an error message
inline std::string make_dev_model(const TsgModel& t, DevModel& m, const float* hdata_dev) {
  if (t.struct_bytes != (int)sizeof(TsgModel)) return "TsgModel.struct_bytes mismatch";
  memset(&m, 0, sizeof(m));
  m.h = t.timestep;
  for (int k = 0; k < 3; k++) m.grav[k] = t.gravity[k];
  m.tol = t.tolerance; m.ls_tol = t.ls_tolerance; m.mpr_tol = t.mpr_tolerance; m.meaninertia = t.meaninertia;
  m.iterations = t.iterations; m.ls_iterations = t.ls_iterations; m.mpr_iterations = t.mpr_iterations;
  m.flags = t.flags;
  if (t.flags & TSG_FLAG_CROSSBAR_DERIV) return "TSG_FLAG_CROSSBAR_DERIV is an oracle-only switch";
  for (int b = 0; b < NBAR; b++) {
    for (int j = 0; j < 6; j++) {
      double M = j < 3 ? t.body_mass[b] : t.body_inertia[b][j - 3];
      m.M[6 * b + j] = M; m.invM[6 * b + j] = 1.0 / M;
    }
    for (int k = 0; k < 3; k++) m.inertia[b][k] = t.body_inertia[b][k];
    m.invw_tran[b] = t.body_invweight0[b][0];
    for (int g = 0; g < 5; g++) {
      int G = 5 * b + g;
      m.gtype[G] = t.geom_type[b][g];
      if (m.gtype[G] != GEOM_SPHERE && m.gtype[G] != GEOM_CYL) return "geom type must be sphere or cylinder";
      m.gsize[G][0] = t.geom_size[b][g][0]; m.gsize[G][1] = t.geom_size[b][g][1];
      for (int k = 0; k < 3; k++) m.gpos[G][k] = t.geom_pos[b][g][k];
      // geom frames must be the body frame up to axis flips (cylinders/spheres are symmetric under those)
      const double* q = t.geom_quat[b][g];
      int big = 0;
      for (int k = 0; k < 4; k++) if (fabs(q[k]) > 1 - 1e-9) big++;
      if (big != 1) return "geom quaternions must be axis flips of the body frame";
    }
  }
  for (int k = 0; k < NQ; k++) m.qpos0[k] = t.qpos0[k];
  for (int b = 0; b < NBAR; b++) m.nends[b] = 0;
  for (int tt = 0; tt < NTEN; tt++) {
    m.ten_act[tt] = -1;
    m.tk[tt] = t.ten_stiffness[tt]; m.tdamp[tt] = t.ten_damping[tt];
    m.tls[tt][0] = t.ten_lengthspring[tt][0]; m.tls[tt][1] = t.ten_lengthspring[tt][1];
    if (t.ten_body[tt][0] == t.ten_body[tt][1]) return "tendon sites must be on different bars";
    for (int e = 0; e < 2; e++) {
      int end = 2 * tt + e, b = t.ten_body[tt][e];
      if (b < 0 || b >= NBAR) return "bad tendon body";
      m.tbody[end] = b;
      for (int k = 0; k < 3; k++) m.tsite[end][k] = t.ten_site[tt][e][k];
      if (m.nends[b] >= 8) return "too many tendon ends on one bar";
      m.ends[b][m.nends[b]++] = end;
    }
  }
  for (int a = 0; a < NACT; a++) {
    m.act_tendon[a] = t.act_tendon[a];
    if (t.act_tendon[a] < 0 || t.act_tendon[a] >= NTEN || m.ten_act[t.act_tendon[a]] >= 0) return "bad actuator tendon";
    m.ten_act[t.act_tendon[a]] = a;
  }
  m.dyntype = t.act_dyntype; m.ctrllimited = t.ctrllimited; m.forcelimited = t.forcelimited;
  m.dynprm0 = t.act_dynprm0; m.gain = t.act_gain;
  for (int k = 0; k < 3; k++) m.bias[k] = t.act_bias[k];
  for (int k = 0; k < 2; k++) { m.ctrlrange[k] = t.ctrlrange[k]; m.forcerange[k] = t.forcerange[k]; }
  if (!(t.solref[0] < 0 && t.solref[1] < 0)) return "only direct (negative) solref is supported";
  if (t.condim != 6) return "condim must be 6";
  double dmax = fmin(MAXIMP, fmax(MINIMP, t.solimp[1]));
  m.K = -t.solref[0] / (dmax * dmax); m.B = -t.solref[1] / dmax;
  for (int k = 0; k < 5; k++) { m.solimp[k] = t.solimp[k]; m.fr[k] = t.friction[k]; }
  m.mu = t.friction[0] / sqrt(t.impratio);
  m.inv_mu2 = 1.0 / (m.mu * m.mu * (1 + m.mu * m.mu));
  m.solscale = 1.0 / (m.meaninertia * NV);
  m.dscale[0] = 1; m.fscale[0] = m.mu;
  for (int j = 1; j < 6; j++) {
    // R_j = (R_0 / impratio) * f0^2 / f_{j-1}^2  ->  D_j = D_0 * dscale_j
    m.dscale[j] = t.impratio * (t.friction[j - 1] * t.friction[j - 1]) / (t.friction[0] * t.friction[0]);
    m.fscale[j] = t.friction[j - 1];
  }
  for (int r = 0; r < 6; r++) { m.wtab[0][r] = m.dscale[r]; m.wtab[1][r] = r ? t.friction[r - 1] * t.friction[r - 1] : 0.0; }
  for (int i = 0, e = 0; i < NV; i++) for (int j = 0; j <= i; j++, e++) { m.tri_i[e] = (unsigned char)i; m.tri_j[e] = (unsigned char)j; }
  m.floor_type = t.floor_type;
  for (int k = 0; k < 3; k++) { m.fpos[k] = t.floor_pos[k]; m.fnormal[k] = t.floor_mat[3 * k + 2]; }
  if (t.floor_type == TSG_FLOOR_HFIELD) {
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < 9; k++) if (fabs(t.floor_mat[k] - I[k]) > 1e-12) return "height field frame must be axis aligned";
    m.nrow = t.hf_nrow; m.ncol = t.hf_ncol;
    for (int k = 0; k < 4; k++) m.hsize[k] = t.hf_size[k];
    m.hdx = 2.0 * m.hsize[0] / (double)(m.ncol - 1); m.hdy = 2.0 * m.hsize[1] / (double)(m.nrow - 1);
    m.hdata = hdata_dev;
    if (!hdata_dev || m.nrow < 2 || m.ncol < 2) return "height field data missing";
  }
  return "";
}

inline std::string make_env_cfg(const TsgEnvConfig& t, const TsgModel& mod, EnvCfg& c) {
  if (t.struct_bytes != (int)sizeof(TsgEnvConfig)) return "TsgEnvConfig.struct_bytes mismatch";
  memset(&c, 0, sizeof(c));
  c.env_kind = t.env_kind; c.task = t.task; c.frame_skip = t.frame_skip; c.obs_dim = t.obs_dim;
  c.use_cap_velocity = t.use_cap_velocity; c.terminate_when_unhealthy = t.terminate_when_unhealthy;
  c.is_test = t.is_test; c.reward_delay_steps = t.reward_delay_steps; c.max_episode_steps = t.max_episode_steps;
  c.warmup_steps = t.warmup_steps; c.npose = t.npose;
  if (c.obs_dim < 1 || c.obs_dim > 64) return "obs_dim out of range";
  if (c.npose < 1 || c.npose > TSG_NPOSE) return "npose out of range";
  if (c.warmup_steps < 1) return "warmup_steps must be >= 1";
  if (c.reward_delay_steps < 1 || c.reward_delay_steps + 1 > HEADING_SLOTS) return "reward_delay_steps out of range";
  if (c.env_kind == ENV_LEGACY && c.task > TASK_TURN) return "tensegrity_env supports straight/turn only";
  c.desired_direction = t.desired_direction; c.ctrl_cost_weight = t.ctrl_cost_weight;
  c.healthy_reward = t.healthy_reward; c.yaw_reward_weight = t.yaw_reward_weight;
  c.min_reset_heading = t.min_reset_heading; c.max_reset_heading = t.max_reset_heading;
  c.tendon_reset_mean = t.tendon_reset_mean; c.tendon_reset_stdev = t.tendon_reset_stdev;
  c.tendon_min_length = t.tendon_min_length; c.tendon_max_length = t.tendon_max_length;
  for (int k = 0; k < 2; k++) { c.waypt_range[k] = t.waypt_range[k]; c.waypt_angle_range[k] = t.waypt_angle_range[k]; }
  c.ditch_reward_max = t.ditch_reward_max; c.ditch_reward_stdev = t.ditch_reward_stdev;
  c.waypt_reward_amplitude = t.waypt_reward_amplitude; c.waypt_reward_stdev = t.waypt_reward_stdev;
  c.kill_force = t.kill_force;
  c.use_obs_noise = t.use_obs_noise ? 1 : 0;
  c.obs_noise_tendon_stdev = t.obs_noise_tendon_stdev; c.obs_noise_cap_pos_stdev = t.obs_noise_cap_pos_stdev;
  if (c.use_obs_noise && c.env_kind != ENV_TR) return "use_obs_noise exists for tr_env only";
  if (c.use_obs_noise && (c.obs_noise_tendon_stdev < 0 || c.obs_noise_cap_pos_stdev < 0)) return "negative obs noise stdev";
  c.dt = mod.timestep * t.frame_skip;
  for (int p = 0; p < TSG_NPOSE; p++) for (int k = 0; k < NQ; k++) c.reset_pose[p][k] = t.reset_pose[p][k];
  return "";
}

}  // namespace tsg
