// tsg_emul.cpp -- TEST HARNESS ONLY.  Compiles the device source (tsg_core.cuh / tsg_env.cuh) as plain
// C++ with TSG_HOST_EMUL: every LANE_FOR runs serially, i.e. one warp is emulated by one host thread.
// Used by tests/test_emul_vs_oracle.py to check the lane logic against the oracle where no GPU exists.
// It is not part of libtsg.so and is never reachable from the product package.
#define TSG_HOST_EMUL 1
#include "../../tensegrity_rl_b200/csrc/tsg_host.h"

#include <stdlib.h>

using namespace tsg;

struct Emul {
  DevModel m;
  EnvCfg c;
  float* hdata;
  EnvScratch S;
  Con spill[MAXC - MAXC_S];
  unsigned long long seed; long long env_id; double* real_obs;   // observation-noise keys / true-observation output
};

extern "C" {

const char* emul_create(const TsgModel* model, const TsgEnvConfig* cfg, void** out) {
  static std::string err;
  Emul* E = new Emul();
  E->hdata = nullptr;
  if (model->floor_type == TSG_FLOOR_HFIELD) {
    size_t n = (size_t)model->hf_nrow * model->hf_ncol;
    E->hdata = (float*)malloc(n * sizeof(float));
    memcpy(E->hdata, model->hf_data, n * sizeof(float));
  }
  err = make_dev_model(*model, E->m, E->hdata);
  if (err.empty()) err = make_env_cfg(*cfg, *model, E->c);
  if (!err.empty()) { delete E; return err.c_str(); }
  memset(&E->S, 0, sizeof(E->S));
  E->S.spill = E->spill;
  E->seed = 0; E->env_id = 0; E->real_obs = nullptr;
  *out = E;
  return nullptr;
}
void emul_destroy(void* h) { Emul* E = (Emul*)h; free(E->hdata); delete E; }
int emul_scratch_bytes() { return (int)sizeof(EnvScratch); }

static StepIO make_io(double* rec, double* heading, const Emul* E = nullptr) {
  StepIO io;
  memset(&io, 0, sizeof(io));
  io.state = rec; io.heading = heading; io.n_envs = 1;
  if (E) { io.seed = E->seed; io.env_id_base = E->env_id; io.real_obs = E->real_obs; }
  return io;
}
void emul_set_noise(void* h, unsigned long long seed, long long env_id, double* real_obs) {
  Emul* E = (Emul*)h; E->seed = seed; E->env_id = env_id; E->real_obs = real_obs;
}
void emul_obs_normals(unsigned long long seed, unsigned long long stream, unsigned long long nreset,
                      unsigned long long step, int n, double* out) {
  for (int pr = 0; pr < (n + 1) / 2; pr++) {
    double z0, z1;
    noise_pair(seed, stream, nreset, step, pr, z0, z1);
    out[2 * pr] = z0;
    if (2 * pr + 1 < n) out[2 * pr + 1] = z1;
  }
}
void emul_step(void* h, double* rec, double* heading, const double* ctrl, double* obs, double* reward,
               uint8_t* done, double* info) {
  Emul* E = (Emul*)h;
  StepIO io = make_io(rec, heading, E);
  io.ctrl64 = ctrl; io.obs = obs; io.reward = reward; io.done = done; io.info = info;
  run_step(E->S, E->m, E->c, io, 0, 0);
}
void emul_reset(void* h, double* rec, double* heading, double* draws, int explicit_draws, unsigned long long seed,
                long long env_id, double* obs) {
  Emul* E = (Emul*)h;
  StepIO io = make_io(rec, heading, E);
  io.draws = draws; io.explicit_draws = explicit_draws; io.seed = seed; io.env_id_base = env_id; io.obs = obs;
  E->seed = seed; E->env_id = env_id;
  run_reset(E->S, E->m, E->c, io, 0, 0);
}
void emul_forward(void* h, double* rec, double* heading, double* obs, double* info) {
  Emul* E = (Emul*)h;
  StepIO io = make_io(rec, heading);
  io.obs = obs; io.info = info;
  run_forward(E->S, E->m, E->c, io, 0, 0);
}
// raw physics: nstep x mj_step on the record with the given ctrl (no env semantics), then cfrc_ext
void emul_mj_step(void* h, double* rec, const double* ctrl, int nstep, double* ten_length, double* cfrc_ext,
                  int* stats) {
  Emul* E = (Emul*)h;
  double heading[HEADING_SLOTS] = {0};
  Aux A;
  load_env(E->S, A, rec, heading, false, 0);
  for (int i = 0; i < NACT; i++) E->S.ctrl[i] = ctrl[i];
  for (int s = 0; s < nstep; s++) substep(E->S, E->m, E->c, 0);
  stage_cfrc(E->S, E->m, 0);
  if (ten_length) for (int i = 0; i < NTEN; i++) ten_length[i] = E->S.tlen[i];
  if (cfrc_ext) for (int i = 0; i < 24; i++) cfrc_ext[i] = E->S.u.post.cfrc[i / 6][i % 6];
  if (stats) { stats[0] = E->S.nact; stats[1] = E->S.niter_total; stats[2] = E->S.nls_total; stats[3] = E->S.nmpr_total; stats[4] = E->S.overflow; stats[5] = E->S.bad; }
  store_env(E->S, A, rec, heading, false, 0);
}
void emul_make_draws(double* d, unsigned long long seed, unsigned long long env_id, unsigned long long nreset) {
  make_draws(d, seed, env_id, nreset);
}
}
