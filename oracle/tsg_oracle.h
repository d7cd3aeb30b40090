/* tsg_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, not product code).
 *
 * Plain-C, single-env, dense, fp64 restatement of what the reference's hot path
 * executes inside MuJoCo 2.3.7 (`mj_step`, `mj_forward`, `mj_rnePostConstraint`)
 * for the 3-bar tensegrity model, reached from
 *   /root/reference/tr_env/tr_env/envs/tr_env.py:346,744,763,800,812
 *   /root/reference/tensegrity_env/tensegrity_env/envs/tensegrity_env.py:297,448,484
 * through gym 0.26.2 MujocoEnv.do_simulation / set_state.
 *
 * PARITY UNPINNED: MuJoCo 2.3.7 (requirements.txt:5) is a third-party dependency
 * that is neither vendored under /root/reference nor installable here, and the
 * reference holds no tests or golden vectors.  This file restates the published
 * algorithm from knowledge; it was never compared with a
 * MuJoCo binary.  What IS pinned: analytic/KKT properties (tests/test_oracle_*.py)
 * and kinematic consistency with the `_last_obs` vectors stored in the
 * reference's SB3 checkpoints (tests/golden/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.
 */
#ifndef TSG_ORACLE_H_
#define TSG_ORACLE_H_

#include "../include/tsg_model.h"

#ifdef __cplusplus
extern "C" {
#endif

#define TSGO_MAXCON 96
#define TSGO_MAXEFC (6 * TSGO_MAXCON)

typedef struct TsgoContact {
  double dist;
  double pos[3];
  double frame[9];
  int32_t geom1, geom2; /* 0 = floor, 1 + 5*bar + k otherwise */
  int32_t body1, body2; /* 0 = world, 1..3 bars */
  int32_t exclude;      /* dist >= 0: detected but no constraint rows */
  int32_t efc_address;  /* first row, or -1 */
} TsgoContact;

typedef struct TsgoData {
  /* state (mjData: qpos, qvel, act, ctrl, qacc_warmstart, time) */
  double qpos[TSG_NQ];
  double qvel[TSG_NV];
  double act[TSG_NACT];
  double ctrl[TSG_NACT];
  double qacc_warmstart[TSG_NV];
  double time;

  /* position stage */
  double xpos[TSG_NBAR][3];
  double xquat[TSG_NBAR][4];
  double xmat[TSG_NBAR][9];
  double geom_xpos[TSG_NBAR][TSG_NGEOM_BAR][3];
  double geom_xmat[TSG_NBAR][TSG_NGEOM_BAR][9];
  double site_xpos[TSG_NTEN][2][3];
  double com_world[3]; /* subtree_com of the world body */
  double ten_length[TSG_NTEN];
  double ten_J[TSG_NTEN][TSG_NV];

  /* velocity / force stage */
  double ten_velocity[TSG_NTEN];
  double qfrc_passive[TSG_NV];
  double qfrc_bias[TSG_NV];
  double qfrc_actuator[TSG_NV];
  double actuator_force[TSG_NACT];
  double act_dot[TSG_NACT];
  double qfrc_smooth[TSG_NV];
  double qacc_smooth[TSG_NV];

  /* constraint stage */
  int32_t ncon, nefc;
  TsgoContact contact[TSGO_MAXCON];
  double efc_J[TSGO_MAXEFC][TSG_NV];
  double efc_pos[TSGO_MAXEFC];
  double efc_vel[TSGO_MAXEFC];
  double efc_aref[TSGO_MAXEFC];
  double efc_R[TSGO_MAXEFC];
  double efc_D[TSGO_MAXEFC];
  double efc_force[TSGO_MAXEFC];
  int32_t efc_state[TSGO_MAXEFC];
  double qacc[TSG_NV];
  double qfrc_constraint[TSG_NV];
  double solver_cost;
  int32_t solver_iter;
  int32_t ls_evals;   /* total line-search evaluations of the last solve */
  int32_t mpr_calls;  /* narrow-phase MPR invocations of the last collision pass */
  int32_t con_overflow;

  /* mj_rnePostConstraint */
  double cfrc_ext[TSG_NBODY][6]; /* [torque; force] about the root's subtree com */

  int32_t warning; /* bit0 bad qpos, bit1 bad qvel, bit2 bad qacc (=> auto reset) */
  int32_t pad_;
} TsgoData;

int tsgo_sizeof_data(void);
int tsgo_sizeof_model(void);

/* mj_resetData */
void tsgo_reset_data(const TsgModel *m, TsgoData *d);
/* mj_forward (position, velocity, actuation, acceleration, constraint) */
void tsgo_forward(const TsgModel *m, TsgoData *d);
/* mj_step(m, d, nstep) */
void tsgo_step(const TsgModel *m, TsgoData *d, int nstep);
/* mj_rnePostConstraint: fills cfrc_ext from the current contacts / efc_force */
void tsgo_rne_post_constraint(const TsgModel *m, TsgoData *d);
/* mj_contactForce(m, d, id, out[6]) */
void tsgo_contact_force(const TsgModel *m, const TsgoData *d, int id, double out[6]);

/* pieces exposed for unit tests */
void tsgo_kinematics(const TsgModel *m, TsgoData *d);
void tsgo_tendon(const TsgModel *m, TsgoData *d);
void tsgo_collision(const TsgModel *m, TsgoData *d);
void tsgo_make_constraint(const TsgModel *m, TsgoData *d);
/* primal cost of a candidate acceleration (Gauss + constraint), and its gradient */
double tsgo_primal_cost(const TsgModel *m, TsgoData *d, const double *qacc, double *grad);

/* MPR on two convex objects; type: 2 sphere, 5 cylinder, 100 prism (size = 18 vertex coords).
 * Returns 1 if penetrating; fills depth, dir (obj1 -> obj2), pos. */
int tsgo_mpr(int type1, const double *pos1, const double *mat1, const double *size1,
             int type2, const double *pos2, const double *mat2, const double *size2,
             double tolerance, int max_iterations, double *depth, double *dir, double *pos);

/* CPU baseline: step n_envs independent envs n_steps env-steps (frame_skip substeps each)
 * with uniform random ctrl in [lo,hi], a pthread pool over envs.  Returns env-steps done. */
long tsgo_bench(const TsgModel *m, int n_envs, int n_steps, int frame_skip, int warm_steps,
                double ctrl_lo, double ctrl_hi, unsigned long long seed, int n_threads,
                double *out_checksum);
int tsgo_max_threads(void);
/* persistent batch of envs stepped by a pthread pool (bench.py CPU arm) */
typedef struct TsgoBatch TsgoBatch;
TsgoBatch *tsgo_batch_create(const TsgModel *m, int n_envs, unsigned long long seed);
void tsgo_batch_destroy(TsgoBatch *b);
long tsgo_batch_step(TsgoBatch *b, int n_steps, int frame_skip, double lo, double hi, int n_threads);
void tsgo_batch_get(const TsgoBatch *b, int e, double *qpos, double *qvel);
/* n states stepped nstep substeps each from identical inputs (threaded): the checker of the full-batch parity tests */
void tsgo_step_states(const TsgModel *m, int n, int nstep, const double *qpos, const double *qvel, const double *act,
                      const double *warm, const double *ctrl, double *out_qpos, double *out_qvel, double *out_ten,
                      int *ncon_minmax, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
