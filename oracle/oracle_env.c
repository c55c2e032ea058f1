/* TEST INFRASTRUCTURE (see oracle_internal.h).  Env layer of the oracle: the Python glue of the
 * reference restated per env, driving the fp64 physics substeps.
 *   before_step / unnormalize_so100    gym_so100/tasks/single_arm.py:33-38, constants.py:44-47,78-86
 *   physics.step(10) + mj_step1        gym_so100/env.py:174 via dm_control
 *   get_reward (CubeToBin)             gym_so100/tasks/single_arm.py:322-380
 *   get_observation / _format_raw_obs  gym_so100/tasks/single_arm.py:82-114, env.py:137-145
 *   initialize_episode                 gym_so100/tasks/single_arm.py:299-309
 *   GoalEnv step / compute_reward      gym_so100/env.py:341-358, 372-406
 */
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "so100_oracle.h"

struct so100o {
  so100_model m;
  int n;
  int task;            /* 0 cube_to_bin (SO100Env), 1 goal env (SO100GoalEnv), 2 touch_cube, 3 touch_cube_sparse (SO100Env) */
  uint64_t seed;
  int64_t env_offset;
  oenv* env;
};

/* ------------------------------------------------------------------ Philox4x32-10 (shared spec with the CUDA reset) */
static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }

so100o* so100o_create(const void* blob, size_t nbytes, int num_envs, int task, uint64_t seed, int64_t env_offset) {
  if (nbytes != sizeof(so100_model)) return 0;
  so100o* h = (so100o*)calloc(1, sizeof(so100o));
  memcpy(&h->m, blob, sizeof(so100_model));
  if (h->m.magic != SO100_MODEL_MAGIC || h->m.version != SO100_MODEL_VERSION) { free(h); return 0; }
  for (int b = 0; b < h->m.nbody; b++)
    if (h->m.body_jtype[b] == SO100_JNT_FREE && norm3(h->m.body_ipos[b]) != 0) { free(h); return 0; }
  h->n = num_envs; h->task = task; h->seed = seed; h->env_offset = env_offset;
  h->env = (oenv*)calloc(num_envs, sizeof(oenv));
  for (int i = 0; i < num_envs; i++) {
    memcpy(h->env[i].qpos, h->m.qpos0, sizeof(h->m.qpos0));
  }
  return h;
}
void so100o_destroy(so100o* h) { if (h) { free(h->env); free(h); } }
int so100o_num_envs(const so100o* h) { return h->n; }

/* ------------------------------------------------------------------ observations / reward */
static void write_obs(const so100o* h, const oenv* e, float* obs, float* achieved, float* desired) {
  const so100_model* m = &h->m;
  if (obs) {   /* env.py:137-145  [box(3), bin(3), ee(3), qpos(6)] as float32 */
    for (int k = 0; k < 3; k++) {
      obs[k] = (float)e->site[m->site_cube][k];
      obs[3 + k] = (float)e->site[m->site_bin][k];
      obs[6 + k] = (float)e->site[m->site_ee][k];
    }
    for (int k = 0; k < 6; k++) obs[9 + k] = (float)e->qpos[k];
  }
  if (achieved) for (int k = 0; k < 3; k++) achieved[k] = (float)e->site[m->site_cube][k];
  if (desired) for (int k = 0; k < 3; k++) desired[k] = e->goal[k];
}

/* single_arm.py:322-380 */
static float cube_to_bin_reward(const so100o* h, const oenv* e) {
  const so100_model* m = &h->m;
  int touch_gripper = 0, touch_table = 0;
  for (int c = 0; c < e->ncon; c++) {
    int g1 = e->con[c].g1, g2 = e->con[c].g2;
    if ((g2 == m->cg_cube && ((m->pad_mask >> g1) & 1)) || (g1 == m->cg_cube && ((m->pad_mask >> g2) & 1))) touch_gripper = 1;
    if (g1 == m->cg_cube && g2 == m->cg_table) touch_table = 1;    /* ordered pair ("red_box","table") */
  }
  float cube[3];
  for (int k = 0; k < 3; k++) cube[k] = (float)e->site[m->site_cube][k];
  double bmin[3], bmax[3];
  const double* c = e->site[m->site_bin];
  bmin[0] = c[0] - m->bin_hw; bmin[1] = c[1] - m->bin_hw; bmin[2] = c[2] + 0.0;
  bmax[0] = c[0] + m->bin_hw; bmax[1] = c[1] + m->bin_hw; bmax[2] = c[2] + m->bin_h;
  int over_bin = (bmin[0] < (double)cube[0] && (double)cube[0] < bmax[0]) && (bmin[1] < (double)cube[1] && (double)cube[1] < bmax[1]);
  int inside = 1;
  float half = (float)m->cube_half;
  for (int k = 0; k < 3; k++) {
    volatile float lower = cube[k] - half, upper = cube[k] + half;   /* float32 array arithmetic */
    if (!((double)lower > bmin[k])) inside = 0;
    if (!((double)upper < bmax[k])) inside = 0;
  }
  int released = inside && !touch_gripper;
  float reward = 0.0f;
  if (touch_gripper) reward = 1.0f;
  if (touch_gripper && !touch_table) reward = 2.0f;
  if (over_bin) reward = 2.5f;
  if (inside) reward = 3.0f;
  if (released) reward = 4.0f;
  return reward;
}

/* single_arm.py:149-215 (dense = 1, SO100TouchCubeTask) and 246-285 (dense = 0, SO100TouchCubeSparseTask): float64
 * throughout, like the reference (site_xpos is float64); success = pad contact and |ee - cube| < 0.05 -> max_reward 4 */
static float touch_reward(const so100o* h, const oenv* e, int dense) {
  const so100_model* m = &h->m;
  int touch_gripper = 0;
  for (int c = 0; c < e->ncon; c++) {
    int g1 = e->con[c].g1, g2 = e->con[c].g2;
    if ((g2 == m->cg_cube && ((m->pad_mask >> g1) & 1)) || (g1 == m->cg_cube && ((m->pad_mask >> g2) & 1))) touch_gripper = 1;
  }
  const double* ee = e->site[m->site_ee];
  const double* cu = e->site[m->site_cube];
  double d = sqrt((ee[0] - cu[0]) * (ee[0] - cu[0]) + (ee[1] - cu[1]) * (ee[1] - cu[1]) + (ee[2] - cu[2]) * (ee[2] - cu[2]));
  double reward = 0.0;
  if (dense) {
    const double thr[5] = {0.7, 0.5, 0.3, 0.1, 0.05}, w[5] = {0.1, 0.2, 0.5, 1.0, 2.0};
    for (int k = 0; k < 5; k++)
      if (d < thr[k]) { double r = w[k] * (1 - d / thr[k]); if (r > reward) reward = r; }
    if (touch_gripper) reward += 1.0;
  }
  if (touch_gripper && d < 0.05) return 4.0f;
  reward -= 0.2;
  return (float)reward;
}

/* env.py:341-358: float32 distance, threshold 0.01 */
static float goal_distance(const float* a, const float* d) {
  volatile float dx = a[0] - d[0], dy = a[1] - d[1], dz = a[2] - d[2];
  volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  volatile float s = xx + yy;
  s = s + zz;
  return sqrtf(s);
}

/* ------------------------------------------------------------------ reset */
static void sample_goal(so100o* h, oenv* e, int64_t gid, const double* box_pose, int32_t total_steps) {
  const so100_model* m = &h->m;
  uint32_t r[4];
  philox((uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), e->episode, 1u, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), r);
  float lo[3], hi[3];
  if (total_steps < m->goal_curriculum_steps) {   /* env.py:324-330 */
    lo[0] = (float)(box_pose[0] - m->lift_goal_xy); hi[0] = (float)(box_pose[0] + m->lift_goal_xy);
    lo[1] = (float)(box_pose[1] - m->lift_goal_xy); hi[1] = (float)(box_pose[1] + m->lift_goal_xy);
    lo[2] = (float)m->lift_goal_zlo; hi[2] = (float)m->lift_goal_zhi;
  } else {
    for (int k = 0; k < 3; k++) { lo[k] = (float)m->bin_goal_lo[k]; hi[k] = (float)m->bin_goal_hi[k]; }
  }
  for (int k = 0; k < 3; k++) {
    volatile float range = hi[k] - lo[k];
    e->goal[k] = fmaf(u01(r[k]), range, lo[k]);
  }
}

static void reset_one(so100o* h, int i, const double* box_pose_or_null, int32_t total_steps) {
  const so100_model* m = &h->m;
  oenv* e = &h->env[i];
  int64_t gid = h->env_offset + i;
  double pose[7] = {0, 0, 0, 1, 0, 0, 0};
  if (box_pose_or_null) {
    memcpy(pose, box_pose_or_null, sizeof(pose));
  } else {   /* utils.py:18-29 with Philox instead of MT19937 (float32 draws) */
    uint32_t r[4];
    philox((uint32_t)gid, (uint32_t)((uint64_t)gid >> 32), e->episode, 0u, (uint32_t)h->seed, (uint32_t)(h->seed >> 32), r);
    for (int k = 0; k < 3; k++) {
      float lo = (float)m->box_lo[k], hi = (float)m->box_hi[k];
      volatile float range = hi - lo;
      pose[k] = (double)fmaf(u01(r[k]), range, lo);
    }
  }
  /* mj_resetData then single_arm.py:303-307 */
  memcpy(e->qpos, m->qpos0, sizeof(m->qpos0));
  memset(e->qvel, 0, sizeof(e->qvel));
  memset(e->warm, 0, sizeof(e->warm));
  memset(e->qacc, 0, sizeof(e->qacc));
  for (int k = 0; k < 6; k++) { e->qpos[k] = m->start_pose[k]; e->ctrl[k] = m->start_pose[k]; }
  for (int k = 0; k < 7; k++) e->qpos[m->nq - 7 + k] = pose[k];
  e->step_count = 0;
  if (h->task == 1) sample_goal(h, e, gid, pose, total_steps);
  e->episode++;
  o_position(m, e);   /* dm_control reset_context exit: mj_forward (position part is what obs/reward read) */
}

int so100o_reset(so100o* h, const uint8_t* mask, const double* box_pose, int32_t* total_steps_io, float* obs, float* achieved, float* desired) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < h->n; i++) {
    if (mask && !mask[i]) continue;
    reset_one(h, i, box_pose ? box_pose + 7 * i : 0, total_steps_io ? total_steps_io[i] : 0);
  }
  for (int i = 0; i < h->n; i++)
    write_obs(h, &h->env[i], obs ? obs + 15 * i : 0, achieved ? achieved + 3 * i : 0, desired ? desired + 3 * i : 0);
  return 0;
}

/* ------------------------------------------------------------------ physics */
static void substep(const so100_model* m, oenv* e) {
  o_position(m, e);
  o_velocity_actuation(m, e);
  o_make_constraints(m, e);
  o_solve(m, e);
  o_integrate(m, e);
}

int so100o_substeps(so100o* h, int nsub) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < h->n; i++)
    for (int s = 0; s < nsub; s++) substep(&h->m, &h->env[i]);
  return 0;
}

/* mj_forward without integration: everything up to qacc / constraint forces */
int so100o_forward(so100o* h) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < h->n; i++) {
    oenv* e = &h->env[i];
    o_position(&h->m, e);
    o_velocity_actuation(&h->m, e);
    o_make_constraints(&h->m, e);
    o_solve(&h->m, e);
  }
  return 0;
}

int so100o_step(so100o* h, const float* action, int autoreset, int32_t* total_steps_io, float* obs, float* achieved,
                float* desired, float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* success, float* final_obs) {
  const so100_model* m = &h->m;
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < h->n; i++) {
    oenv* e = &h->env[i];
    /* constants.py:44-47 in float32 (numpy >= 2 keeps float32 for float32-scalar (op) python-float) */
    for (int k = 0; k < 6; k++) {
      float lo = (float)m->act_lo[k], hi = (float)m->act_hi[k];
      float range = (float)(m->act_hi[k] - m->act_lo[k]);
      volatile float t = action[6 * i + k] + 1.0f;
      t = t / 2.0f;
      t = t * range;
      t = t + lo;
      float u = t < lo ? lo : (t > hi ? hi : t);
      e->ctrl[k] = (double)u;
    }
    for (int s = 0; s < m->nsubstep; s++) substep(m, e);
    o_position(m, e);     /* dm_control legacy step: trailing mj_step1 */
    e->step_count++;
    int32_t total = total_steps_io ? ++total_steps_io[i] : 0;
    float r; int term, trunc, succ;
    float ag[3], dg[3];
    write_obs(h, e, 0, ag, dg);
    if (h->task == 0) {
      r = cube_to_bin_reward(h, e);
      succ = r == 4.0f; term = succ;                       /* env.py:175 */
      trunc = e->step_count >= m->max_episode_steps;       /* TimeLimit wrapper, __init__.py:27 */
    } else if (h->task >= 2) {
      r = touch_reward(h, e, h->task == 2);
      succ = r == 4.0f; term = succ;                       /* env.py:175 */
      trunc = e->step_count >= 300;                        /* TimeLimit wrapper, __init__.py:7,17 */
    } else {
      float d = goal_distance(ag, dg);
      succ = d < (float)m->goal_threshold;
      r = succ ? 0.0f : -1.0f;
      term = succ;
      trunc = e->step_count >= 300;                        /* env.py:200, 398 */
    }
    if (reward) reward[i] = r;
    if (terminated) terminated[i] = (uint8_t)term;
    if (truncated) truncated[i] = (uint8_t)trunc;
    if (success) success[i] = (uint8_t)succ;
    if (final_obs) write_obs(h, e, final_obs + 15 * i, 0, 0);
    if (autoreset && (term || trunc)) reset_one(h, i, 0, total);
    write_obs(h, e, obs ? obs + 15 * i : 0, achieved ? achieved + 3 * i : 0, desired ? desired + 3 * i : 0);
  }
  return 0;
}

/* ------------------------------------------------------------------ state access */
int so100o_set_state(so100o* h, const double* qpos, const double* qvel, const double* ctrl, const double* warm) {
  for (int i = 0; i < h->n; i++) {
    oenv* e = &h->env[i];
    if (qpos) memcpy(e->qpos, qpos + 13 * i, 13 * sizeof(double));
    if (qvel) memcpy(e->qvel, qvel + 12 * i, 12 * sizeof(double));
    if (ctrl) memcpy(e->ctrl, ctrl + 6 * i, 6 * sizeof(double));
    if (warm) memcpy(e->warm, warm + 12 * i, 12 * sizeof(double));
  }
  return 0;
}
int so100o_get_state(const so100o* h, double* qpos, double* qvel, double* ctrl, double* warm) {
  for (int i = 0; i < h->n; i++) {
    const oenv* e = &h->env[i];
    if (qpos) memcpy(qpos + 13 * i, e->qpos, 13 * sizeof(double));
    if (qvel) memcpy(qvel + 12 * i, e->qvel, 12 * sizeof(double));
    if (ctrl) memcpy(ctrl + 6 * i, e->ctrl, 6 * sizeof(double));
    if (warm) memcpy(warm + 12 * i, e->warm, 12 * sizeof(double));
  }
  return 0;
}
int so100o_set_goal(so100o* h, const float* goal) {
  for (int i = 0; i < h->n; i++) memcpy(h->env[i].goal, goal + 3 * i, 3 * sizeof(float));
  return 0;
}
int so100o_set_counters(so100o* h, const int32_t* step_count, const uint32_t* episode) {
  for (int i = 0; i < h->n; i++) {
    if (step_count) h->env[i].step_count = step_count[i];
    if (episode) h->env[i].episode = episode[i];
  }
  return 0;
}

/* debug / parity accessors for env i */
int so100o_get_dyn(const so100o* h, int i, double* M, double* bias, double* qfrc_act, double* qacc_smooth, double* qacc,
                   double* sites, double* xpos, double* xquat) {
  const oenv* e = &h->env[i];
  if (M) for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) M[r * 12 + c] = e->M[r * NVMAX + c];
  if (bias) memcpy(bias, e->bias, 12 * sizeof(double));
  if (qfrc_act) memcpy(qfrc_act, e->qfrc_act, 12 * sizeof(double));
  if (qacc_smooth) memcpy(qacc_smooth, e->qacc_smooth, 12 * sizeof(double));
  if (qacc) memcpy(qacc, e->qacc, 12 * sizeof(double));
  if (sites) memcpy(sites, e->site, h->m.nsite * 3 * sizeof(double));
  if (xpos) memcpy(xpos, e->xpos, h->m.nbody * 3 * sizeof(double));
  if (xquat) memcpy(xquat, e->xquat, h->m.nbody * 4 * sizeof(double));
  return 0;
}
/* contacts of env i: returns ncon; geom[2*c..] = MuJoCo geom ids; data[c] = dist, pos3, normal3, force4 (11 doubles) */
int so100o_get_contacts(const so100o* h, int i, int maxc, int32_t* geom, double* data) {
  const oenv* e = &h->env[i];
  int n = e->ncon < maxc ? e->ncon : maxc;
  for (int c = 0; c < n; c++) {
    const ocontact* k = &e->con[c];
    geom[2 * c] = h->m.geom_mjid[k->g1]; geom[2 * c + 1] = h->m.geom_mjid[k->g2];
    double* d = data + 11 * c;
    d[0] = k->dist; memcpy(d + 1, k->pos, 3 * sizeof(double)); memcpy(d + 4, k->frame, 3 * sizeof(double));
    memcpy(d + 7, k->force, 4 * sizeof(double));
  }
  return e->ncon;
}
int so100o_get_solver(const so100o* h, int i, int* nefc, int* iters, double* grad, int* overflow) {
  const oenv* e = &h->env[i];
  if (nefc) *nefc = e->nefc;
  if (iters) *iters = e->solver_iter;
  if (grad) *grad = e->solver_grad;
  if (overflow) *overflow = e->overflow;
  return 0;
}
int so100o_get_efc(const so100o* h, int i, int maxr, double* J, double* aref, double* R, double* force, double* jar) {
  const oenv* e = &h->env[i];
  int n = e->nefc < maxr ? e->nefc : maxr;
  for (int r = 0; r < n; r++) {
    if (J) memcpy(J + 12 * r, e->J[r], 12 * sizeof(double));
    if (aref) aref[r] = e->earef[r];
    if (R) R[r] = e->eR[r];
    if (force) force[r] = e->eforce[r];
    if (jar) jar[r] = e->ejar[r];
  }
  return e->nefc;
}
/* reward truth-table hook for the golden vectors: synthetic contact list (MuJoCo geom ids, geom1 first)
   and cube_site position on env 0, everything else as left by the last position stage */
float so100o_test_reward(so100o* h, int ncon, const int32_t* geom_mjid_pairs, const double* cube_site) {
  oenv* e = &h->env[0];
  const so100_model* m = &h->m;
  e->ncon = 0;
  for (int c = 0; c < ncon && c < MAXCON; c++) {
    int g[2] = {-1, -1};
    for (int s = 0; s < 2; s++)
      for (int k = 0; k < m->ngeom; k++)
        if (m->geom_mjid[k] == geom_mjid_pairs[2 * c + s]) g[s] = k;
    if (g[0] < 0 || g[1] < 0) return -1000.0f;
    e->con[e->ncon].g1 = g[0]; e->con[e->ncon].g2 = g[1];
    e->ncon++;
  }
  memcpy(e->site[m->site_cube], cube_site, 3 * sizeof(double));
  return cube_to_bin_reward(h, e);
}

/* the same for the touch tasks (task 2 = so100_touch_cube, 3 = so100_touch_cube_sparse) with an injected ee_site */
float so100o_test_touch_reward(so100o* h, int task, int ncon, const int32_t* geom_mjid_pairs, const double* cube_site,
                               const double* ee_site) {
  oenv* e = &h->env[0];
  const so100_model* m = &h->m;
  e->ncon = 0;
  for (int c = 0; c < ncon && c < MAXCON; c++) {
    int g[2] = {-1, -1};
    for (int s = 0; s < 2; s++)
      for (int k = 0; k < m->ngeom; k++)
        if (m->geom_mjid[k] == geom_mjid_pairs[2 * c + s]) g[s] = k;
    if (g[0] < 0 || g[1] < 0) return -1000.0f;
    e->con[e->ncon].g1 = g[0]; e->con[e->ncon].g2 = g[1];
    e->ncon++;
  }
  memcpy(e->site[m->site_cube], cube_site, 3 * sizeof(double));
  memcpy(e->site[m->site_ee], ee_site, 3 * sizeof(double));
  return touch_reward(h, e, task == 2);
}

/* worker threads of the OpenMP loops over envs (torchrun exports OMP_NUM_THREADS=1; the CPU arm of bench.py uses all cores) */
int so100o_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* batched HER reward, env.py:346-349 */
int so100o_compute_reward(const float* ag, const float* dg, int n, float thr, float* out) {
  for (int i = 0; i < n; i++) out[i] = goal_distance(ag + 3 * i, dg + 3 * i) < thr ? 0.0f : -1.0f;
  return 0;
}
/* constants.py:78-86 on a batch, float32 */
int so100o_unnormalize(const so100o* h, const float* action, int n, float* out) {
  const so100_model* m = &h->m;
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++) {
      float lo = (float)m->act_lo[k], hi = (float)m->act_hi[k], range = (float)(m->act_hi[k] - m->act_lo[k]);
      volatile float t = action[6 * i + k] + 1.0f;
      t = t / 2.0f; t = t * range; t = t + lo;
      out[6 * i + k] = t < lo ? lo : (t > hi ? hi : t);
    }
  return 0;
}
