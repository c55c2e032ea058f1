/* TEST INFRASTRUCTURE -- CPU fp64 oracle of the bin-a-cube env step.
 *
 * A plain-C restatement of what the reference executes per `env.step` / `env.reset`
 * (gym_so100/env.py:148-182, 302-406; gym_so100/tasks/single_arm.py:33-38, 82-114,
 * 299-380) including the physics it delegates to the un-vendored MuJoCo 3.3.3 wheel
 * (`mj_step` x10 + `mj_step1`).  PARITY UNPINNED against MuJoCo
 * itself: neither this container nor the GPU box has a MuJoCo build, and the reference
 * ships no golden vectors for the physics.  It is pinned only on the numpy-only pieces of
 * the reference (tests/golden) and on analytic known answers (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product path (gym_so100_c_b200/csrc) never does.
 */
#ifndef SO100_ORACLE_INTERNAL_H_
#define SO100_ORACLE_INTERNAL_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../include/so100_model.h"

#define NB SO100_MAXBODY
#define NVMAX SO100_MAXDOF
#define NG SO100_MAXGEOM
#define NS SO100_MAXSITE
#define MAXCON 96
#define MAXEFC (12 + 6 + 4 * MAXCON)
#define MJMINVAL 1e-15
#define MJMINIMP 0.0001
#define MJMAXIMP 0.9999

enum { EFC_FRICTION = 0, EFC_LIMIT = 1, EFC_CONTACT = 2, EFC_CONTACT_CONT = 3 };

typedef struct {
  int pair, g1, g2, dim; /* g1/g2: collidable-geom indices, g1 has the lower (type,id) */
  double dist, pos[3], frame[9];
  double friction[3], solref[2], solimp[5], mu;
  int efc;
  double force[4];
} ocontact;

typedef struct {
  /* state */
  double qpos[SO100_MAXQ], qvel[NVMAX], ctrl[SO100_MAXACT], warm[NVMAX], qacc[NVMAX];
  float goal[3];
  int32_t step_count;
  uint32_t episode;
  /* position stage */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  double gpos[NG][3], gmat[NG][9], gcen[NG][3];
  double site[NS][3];
  double M[NVMAX * NVMAX], Lchol[NVMAX * NVMAX];
  int ncon, overflow;
  ocontact con[MAXCON];
  /* velocity / force stage */
  double bias[NVMAX], qfrc_act[NVMAX], qfrc_smooth[NVMAX], qacc_smooth[NVMAX], qfrc_constraint[NVMAX];
  /* constraints */
  int nefc;
  int etype[MAXEFC], eid[MAXEFC];
  double J[MAXEFC][NVMAX], epos[MAXEFC], ediag[MAXEFC], eR[MAXEFC], eD[MAXEFC], earef[MAXEFC],
      evel[MAXEFC], efloss[MAXEFC], eforce[MAXEFC], ejar[MAXEFC];
  int solver_iter;
  double solver_grad;
} oenv;

/* ---- small vector helpers ---- */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void sub3(double* r, const double* a, const double* b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void add3(double* r, const double* a, const double* b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void addscl3(double* r, const double* a, const double* b, double s) { r[0] = a[0] + s * b[0]; r[1] = a[1] + s * b[1]; r[2] = a[2] + s * b[2]; }
static inline void scl3(double* r, const double* a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void copy3(double* r, const double* a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MJMINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  a[0] /= n; a[1] /= n; a[2] /= n;
  return n;
}
/* row-major 3x3: r = M v, r = M^T v */
static inline void mulmv(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
         z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void mulmtv(double* r, const double* m, const double* v) {
  double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
         z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void quat_mul(double* r, const double* a, const double* b) {
  double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static inline void quat2mat(double* m, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
static inline void quat_normalize(double* q) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MJMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

/* ---- module entry points ---- */
void o_position(const so100_model* m, oenv* e);                 /* kinematics, M, factor, collision */
void o_collide(const so100_model* m, oenv* e);                  /* oracle_collide.c */
void o_velocity_actuation(const so100_model* m, oenv* e);       /* bias, actuator, qacc_smooth */
void o_make_constraints(const so100_model* m, oenv* e);
void o_solve(const so100_model* m, oenv* e);                    /* oracle_solve.c */
void o_integrate(const so100_model* m, oenv* e);
void o_jac(const so100_model* m, const oenv* e, int body, const double* point, double* jp, double* jr);
int o_chol(double* L, const double* A, int n);
void o_chol_solve(const double* L, double* x, int n);

#endif
