/* tsg_model.h -- plain-data description of the 3-bar tensegrity model and of the
 * env semantics layered on it.  Shared (as a DATA declaration only) by the CUDA
 * product library (libtsg.so) and by the CPU oracle (oracle/, test
 * infrastructure).  It is what MjModel.from_xml_path() produces for the two
 * reference XMLs, restricted to the fields the hot path reads:
 *   /root/reference/3prism_jonathan_steady_side.xml:13-23,30-39,61-62,71-124,127-164,204-210
 *   /root/reference/3prism_jonathan_steady_side_uneven_ground.xml:38-39,48,56,65-118,122-158
 * All arrays are fixed size: the topology (3 free bars, 5 geoms per bar, 9
 * two-site spatial tendons, 6 tendon actuators) is the reference's.
 */
#ifndef TSG_MODEL_H_
#define TSG_MODEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSG_NBAR 3
#define TSG_NGEOM_BAR 5 /* rXY cylinder, s+ sphere, s- sphere, b+ cylinder, b- cylinder */
#define TSG_NTEN 9
#define TSG_NACT 6
#define TSG_NQ 21
#define TSG_NV 18
#define TSG_NBODY 4 /* world + 3 bars (cfrc_ext rows) */

#define TSG_GEOM_SPHERE 2   /* mjGEOM_SPHERE   */
#define TSG_GEOM_CYLINDER 5 /* mjGEOM_CYLINDER */

#define TSG_FLOOR_PLANE 0
#define TSG_FLOOR_HFIELD 1

#define TSG_DYN_NONE 0
#define TSG_DYN_FILTER 2 /* mjDYN_FILTER */

/* model flags: details of MuJoCo 2.3.7 that could not be verified offline
 * (SURVEY.md Appendix E); defaults are the believed-2.3.7 behaviour. */
#define TSG_FLAG_ACTVEL_WHEN_CLAMPED 1u /* E.2: keep d(force)/d(vel) in qDeriv while force is clamped */
#define TSG_FLAG_CROSSBAR_DERIV 2u      /* E.1: (off) keep cross-bar J'BJ coupling in implicitfast */
#define TSG_FLAG_FIXNORMAL 4u           /* sphere-specific normal after MPR (mjc_fixNormal) */

typedef struct TsgModel {
  int32_t struct_bytes; /* sizeof(TsgModel), checked by tsg_create / the oracle */
  uint32_t flags;

  /* <option> */
  double timestep;
  double gravity[3];
  double tolerance;
  double ls_tolerance;
  double impratio;
  double mpr_tolerance;
  int32_t iterations;
  int32_t ls_iterations;
  int32_t mpr_iterations;
  int32_t pad0_;

  /* bodies: free joint each; COM at body origin, principal axes = body axes */
  double body_mass[TSG_NBAR];
  double body_inertia[TSG_NBAR][3];
  double body_invweight0[TSG_NBAR][2]; /* translational, rotational */
  double meaninertia;
  double qpos0[TSG_NQ];

  /* geoms, per bar, body-local frames */
  int32_t geom_type[TSG_NBAR][TSG_NGEOM_BAR];
  int32_t pad1_;
  double geom_size[TSG_NBAR][TSG_NGEOM_BAR][3]; /* sphere: r; cylinder: r, half-length */
  double geom_pos[TSG_NBAR][TSG_NGEOM_BAR][3];
  double geom_quat[TSG_NBAR][TSG_NGEOM_BAR][4];
  double geom_rbound[TSG_NBAR][TSG_NGEOM_BAR];

  /* tendons: two sites each */
  int32_t ten_body[TSG_NTEN][2];        /* bar index 0..2 of each site */
  double ten_site[TSG_NTEN][2][3];      /* site position, body-local */
  double ten_stiffness[TSG_NTEN];
  double ten_damping[TSG_NTEN];
  double ten_lengthspring[TSG_NTEN][2]; /* lower, upper (equal for a single springlength) */

  /* actuators: actuator i drives tendon act_tendon[i], gear 1 */
  int32_t act_tendon[TSG_NACT];
  int32_t act_dyntype;
  int32_t ctrllimited;
  int32_t forcelimited;
  int32_t pad2_;
  double act_dynprm0;  /* filter time constant */
  double act_gain;     /* gainprm[0], gaintype fixed */
  double act_bias[3];  /* biasprm[0..2], affine (all zero for biastype none) */
  double ctrlrange[2];
  double forcerange[2];

  /* contact parameters (identical for all geom pairs after MuJoCo's mixing) */
  double solref[2];
  double solimp[5];
  double friction[5]; /* expanded: slide, slide, spin, roll, roll */
  int32_t condim;

  /* floor */
  int32_t floor_type;
  double floor_pos[3];
  double floor_mat[9];
  int32_t hf_nrow, hf_ncol;
  double hf_size[4];     /* radius_x, radius_y, elevation_z, base_z */
  const float *hf_data;  /* nrow*ncol, row-major, normalised to [0,1]; host pointer, copied */
} TsgModel;

/* ---- env semantics (tr_env.py / tensegrity_env.py) ---------------------- */

#define TSG_ENV_TR 0     /* /root/reference/tr_env/tr_env/envs/tr_env.py */
#define TSG_ENV_LEGACY 1 /* /root/reference/tensegrity_env/tensegrity_env/envs/tensegrity_env.py */

#define TSG_TASK_STRAIGHT 0
#define TSG_TASK_TURN 1
#define TSG_TASK_AIMING 2
#define TSG_TASK_TRACKING 3
#define TSG_TASK_VEL_TRACK 4

#define TSG_HEADING_SLOTS 32 /* ring buffer capacity >= reward_delay_steps + 1 */
#define TSG_NPOSE 6          /* rolling_qpos table, tr_env.py:723-728 */

typedef struct TsgEnvConfig {
  int32_t struct_bytes;
  int32_t env_kind;           /* TSG_ENV_* */
  int32_t task;               /* TSG_TASK_* (desired_action) */
  int32_t frame_skip;         /* 20: tr_env.py:273, tensegrity_env.py:238 */
  int32_t obs_dim;            /* 27/45/48 (tr_env.py:262-268) or 39 (tensegrity_env.py:231) */
  int32_t use_cap_velocity;   /* tr_env.py:141 */
  int32_t terminate_when_unhealthy;
  int32_t is_test;            /* tr_env.py:145 */
  int32_t reward_delay_steps; /* int(reward_delay_seconds/dt): 1 (tr_env) / 25 (legacy) */
  int32_t max_episode_steps;  /* TimeLimit, tr_env/tr_env/__init__.py:6 (5000); 0 = off */
  int32_t warmup_steps;       /* 50: tr_env.py:811, tensegrity_env.py:495 */
  int32_t npose;              /* number of rows used in reset_pose (6 tr_env, 1 legacy=qpos0) */
  double desired_direction;
  double ctrl_cost_weight;    /* 0.01 tr_env.py:148 ; 0.001 tensegrity_env.py:163 */
  double healthy_reward;      /* 0.1 */
  double yaw_reward_weight;   /* tr_env.py:171 */
  double min_reset_heading, max_reset_heading;
  double tendon_reset_mean, tendon_reset_stdev, tendon_min_length, tendon_max_length;
  double waypt_range[2];       /* tr_env.py:164 */
  double waypt_angle_range[2]; /* tr_env.py:165 */
  double ditch_reward_max, ditch_reward_stdev;  /* tr_env.py:167-168 */
  double waypt_reward_amplitude, waypt_reward_stdev; /* tr_env.py:169-170 */
  double kill_force;           /* 1500: tr_env.py:480 */
  double reset_pose[TSG_NPOSE][TSG_NQ];
  /* observation noise (tr_env.py:142,161-162,552-644): obs = real_obs + N(0, stdev) per component, the tracking
   * vector / yaw re-derived from the noisy cap positions; rewards and termination always use the true state */
  int32_t use_obs_noise;
  int32_t pad_noise_;
  double obs_noise_tendon_stdev;  /* 0.02 */
  double obs_noise_cap_pos_stdev; /* 0.05: end-cap positions AND end-cap velocities (tr_env.py:606-617) */
  /* reset noise (tr_env.py:734-743, tensegrity_env.py:436-445): qpos += U(-s, s) per component, qvel = s * N(0, 1);
   * off (0) in the reference defaults.  Draws: Philox(seed, env id, reset count), reproducible */
  double reset_noise_scale;
  /* contact cost (tr_env.py:292-304, 513-516): reward -= weight * sum(clip(cfrc_ext, range)^2) over the 4 x 6 external
   * contact wrench rows, and info["reward_ctrl"] = -contact_cost; off in the reference defaults */
  int32_t use_contact_forces;
  int32_t pad_contact_;
  double contact_cost_weight;     /* 5e-4 */
  double contact_force_range[2];  /* (-1, 1); (-1000, 1000) for desired_action "turn" (tr_env.py:255-256) */
} TsgEnvConfig;

/* draws consumed by one reset (reference: unseeded numpy, tr_env.py:730,775,802-804,831-832) */
#define TSG_NDRAW 10 /* pose u01, heading u01, 6 x N(0,1), waypoint length u01, waypoint yaw u01 */

#ifdef __cplusplus
}
#endif
#endif /* TSG_MODEL_H_ */
