/* TEST INFRASTRUCTURE -- public entry points of the CPU fp64 oracle (see oracle_internal.h for
 * what it restates and why parity with MuJoCo itself is unpinned). */
#ifndef SO100_ORACLE_H_
#define SO100_ORACLE_H_
#include <stddef.h>
#include "oracle_internal.h"

typedef struct so100o so100o;

so100o* so100o_create(const void* blob, size_t nbytes, int num_envs, int task, uint64_t seed, int64_t env_offset);
void so100o_destroy(so100o* h);
int so100o_num_envs(const so100o* h);
int so100o_reset(so100o* h, const uint8_t* mask, const double* box_pose, int32_t* total_steps_io, float* obs,
                 float* achieved, float* desired);
int so100o_substeps(so100o* h, int nsub);
int so100o_forward(so100o* h);
int so100o_step(so100o* h, const float* action, int autoreset, int32_t* total_steps_io, float* obs, float* achieved,
                float* desired, float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* success, float* final_obs);
int so100o_set_state(so100o* h, const double* qpos, const double* qvel, const double* ctrl, const double* warm);
int so100o_get_state(const so100o* h, double* qpos, double* qvel, double* ctrl, double* warm);
int so100o_set_goal(so100o* h, const float* goal);
int so100o_set_counters(so100o* h, const int32_t* step_count, const uint32_t* episode);
int so100o_get_dyn(const so100o* h, int i, double* M, double* bias, double* qfrc_act, double* qacc_smooth, double* qacc,
                   double* sites, double* xpos, double* xquat);
int so100o_get_contacts(const so100o* h, int i, int maxc, int32_t* geom, double* data);
int so100o_get_solver(const so100o* h, int i, int* nefc, int* iters, double* grad, int* overflow);
int so100o_get_efc(const so100o* h, int i, int maxr, double* J, double* aref, double* R, double* force, double* jar);
int so100o_compute_reward(const float* ag, const float* dg, int n, float thr, float* out);
/* 0: checker tolerances (default), 1: MuJoCo's solver settings (bench.py's CPU baseline), see oracle_solve.c */
void so100o_set_solver_mode(int mode);
int so100o_unnormalize(const so100o* h, const float* action, int n, float* out);
#endif
