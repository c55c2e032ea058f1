// Kernels: fused step, reset, physics-only substeps, forward (parity/debug), HER reward,
// state pack/unpack.  Env e is handled by tile (e % EPB) of block (e / EPB).
#pragma once
#include "so100_solve.cuh"
#include "so100_collide.cuh"

namespace so100 {

struct StepArgs {
  float* state;            // [N, STATE_WORDS]
  const float* action;     // [N,6]
  float *obs, *achieved, *desired, *reward, *final_obs;
  uint8_t *terminated, *truncated, *success;
  int n, autoreset, task;
  uint32_t seed_lo, seed_hi;
  long long env_offset;
};

template <unsigned LPE> __device__ __forceinline__ void load_state(const Tile<LPE>& t, EnvS* S, const float* state, int env) {
  const float* src = state + (size_t)env * STATE_WORDS;
  for (int k = t.thread_rank(); k < STATE_WORDS; k += LPE) S->st[k] = src[k];
  t.sync();
}
template <unsigned LPE> __device__ __forceinline__ void store_state(const Tile<LPE>& t, const EnvS* S, float* state, int env) {
  t.sync();
  float* dst = state + (size_t)env * STATE_WORDS;
  for (int k = t.thread_rank(); k < STATE_WORDS; k += LPE) dst[k] = S->st[k];
}

// one full MuJoCo substep (mj_step) on the state in S->st
#ifdef SO100_PHASE_SYNC
#define PHASE_SYNC() __syncthreads()
#else
#define PHASE_SYNC() do { } while (0)
#endif
template <unsigned LPE> __device__ void substep(const Tile<LPE>& t, EnvS* S, const DevTables& T) {
  PROF_BEGIN();
  PHASE_SYNC();
  kinematics(t, S);
  mass_matrix(t, S);
  t.sync();
  PROF_MARK(0);
  PHASE_SYNC();
  smooth_forces(t, S);
  const float qas = smooth_acc(t, S);
  PROF_MARK(1);
  PHASE_SYNC();
  collide_env(t, S, T);
  PROF_MARK(2);
  PHASE_SYNC();
  make_contact_rows(t, S, T);
  PROF_MARK(3);
  solve(t, S, T, qas, reinterpret_cast<uint32_t*>(&S->st[S_DIAG]));
  PROF_MARK(4);
  PHASE_SYNC();
  integrate(t, S);
  PROF_MARK(5);
}

// utils.py:18-29 / single_arm.py:299-309 / env.py:322-334 on the device
template <unsigned LPE>
__device__ void reset_env(const Tile<LPE>& t, EnvS* S, long long gid, const float* box_pose, int task, uint32_t seed_lo,
                          uint32_t seed_hi) {
  const int lane = t.thread_rank();
  t.sync();
  if (lane == 0) {
    uint32_t episode = __float_as_uint(S->st[S_EPISODE]);
    float pose[7] = {0, 0, 0, 1, 0, 0, 0};
    if (box_pose) {
#pragma unroll
      for (int k = 0; k < 7; k++) pose[k] = box_pose[k];
    } else {
      uint32_t r[4];
      philox4x32((uint32_t)gid, (uint32_t)((unsigned long long)gid >> 32), episode, 0u, seed_lo, seed_hi, r);
#pragma unroll
      for (int k = 0; k < 3; k++) pose[k] = __fmaf_rn(u01(r[k]), c_m.box_range[k], c_m.box_lo[k]);
    }
#pragma unroll
    for (int k = 0; k < NL; k++) { S->st[S_QPOS + k] = c_m.start_pose[k]; S->st[S_CTRL + k] = c_m.start_pose[k]; }
#pragma unroll
    for (int k = 0; k < 7; k++) S->st[S_QPOS + 6 + k] = pose[k];
#pragma unroll
    for (int k = 0; k < NV; k++) { S->st[S_QVEL + k] = 0.0f; S->st[S_WARM + k] = 0.0f; }
    S->st[S_STEP] = __int_as_float(0);
    if (task == 1) {
      uint32_t r[4];
      philox4x32((uint32_t)gid, (uint32_t)((unsigned long long)gid >> 32), episode, 1u, seed_lo, seed_hi, r);
      float lo[3], hi[3];
      if (__float_as_int(S->st[S_TOTAL]) < c_m.curriculum_steps) {
        lo[0] = __fsub_rn(pose[0], c_m.lift_xy); hi[0] = __fadd_rn(pose[0], c_m.lift_xy);
        lo[1] = __fsub_rn(pose[1], c_m.lift_xy); hi[1] = __fadd_rn(pose[1], c_m.lift_xy);
        lo[2] = c_m.lift_zlo; hi[2] = c_m.lift_zhi;
      } else {
#pragma unroll
        for (int k = 0; k < 3; k++) { lo[k] = c_m.bin_goal_lo[k]; hi[k] = c_m.bin_goal_hi[k]; }
      }
#pragma unroll
      for (int k = 0; k < 3; k++) S->st[S_GOAL + k] = __fmaf_rn(u01(r[k]), __fsub_rn(hi[k], lo[k]), lo[k]);
    }
    S->st[S_EPISODE] = __uint_as_float(episode + 1u);
  }
  t.sync();
}

struct SiteOut { V3 cube, ee; };
__device__ __forceinline__ SiteOut sites(const EnvS* S) {
  SiteOut o;
  o.cube = ld3(S->lpos[NL]) + mulmv(S->lmat[NL], ld3(c_m.cube_site_off));
  o.ee = ld3(S->lpos[4]) + mulmv(S->lmat[4], ld3(c_m.ee_off));
  return o;
}

template <unsigned LPE>
__device__ __forceinline__ void write_obs(const Tile<LPE>& t, const EnvS* S, int env, float* obs, float* achieved,
                                          float* desired) {
  const int lane = t.thread_rank();
  const SiteOut so = sites(S);
  for (int k = lane; k < 15; k += LPE) {
    float v;
    if (k < 3) v = comp(so.cube, k);
    else if (k < 6) v = c_m.bin_center[k - 3];
    else if (k < 9) v = comp(so.ee, k - 6);
    else v = S->st[S_QPOS + k - 9];
    if (obs) obs[(size_t)env * 15 + k] = v;
  }
  if (lane < 3) {
    if (achieved) achieved[(size_t)env * 3 + lane] = comp(so.cube, lane);
    if (desired) desired[(size_t)env * 3 + lane] = S->st[S_GOAL + lane];
  }
}

#ifndef SO100_MINB
#define SO100_MINB 3     // resident 128-thread blocks per SM the register allocation must allow
#endif
#ifndef SO100_BLOCK
#define SO100_BLOCK 128  // threads per block of the fused step kernel
#endif
template <unsigned LPE> __global__ void __launch_bounds__(SO100_BLOCK, SO100_MINB) step_kernel(StepArgs A, DevTables T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int EPB = SO100_BLOCK / LPE;
  cg::thread_block blk = cg::this_thread_block();
  Tile<LPE> t = cg::tiled_partition<LPE>(blk);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= A.n) return;
  EnvS* S = reinterpret_cast<EnvS*>(smem_raw) + t.meta_group_rank();
  const int lane = t.thread_rank();
  load_state(t, S, A.state, env);
  // before_step: unnormalize_so100 in float32 (constants.py:44-47, 78-86)
  if (lane < NL) {
    float v = __fadd_rn(A.action[(size_t)env * 6 + lane], 1.0f);
    v = __fdiv_rn(v, 2.0f);
    v = __fmul_rn(v, c_m.act_range[lane]);
    v = __fadd_rn(v, c_m.act_lo[lane]);
    S->st[S_CTRL + lane] = fminf(fmaxf(v, c_m.act_lo[lane]), c_m.act_hi[lane]);
  }
  t.sync();
  for (int s = 0; s < c_m.nsub; s++) substep(t, S, T);
  // trailing mj_step1: positions + contacts of the new state
  kinematics(t, S);
  collide_env(t, S, T);
  // ---- task layer
  uint32_t* diag = reinterpret_cast<uint32_t*>(&S->st[S_DIAG]);
  const int ncon_raw = S->ncon, ncon = min(ncon_raw, NC);
  bool bad = false;
  for (int k = lane; k < S_GOAL; k += LPE) bad |= !isfinite(S->st[k]);
  bad = t.any(bad);
  int tg = 0, tt = 0;
  for (int c = lane; c < ncon; c += LPE) {
    const DevPair& P = T.pair[S->cpair[c]];
    if ((P.g2 == c_m.cg_cube && ((c_m.pad_mask >> P.g1) & 1u)) || (P.g1 == c_m.cg_cube && ((c_m.pad_mask >> P.g2) & 1u))) tg = 1;
    if (P.g1 == c_m.cg_cube && P.g2 == c_m.cg_table) tt = 1;
  }
  const bool touch_gripper = t.any(tg), touch_table = t.any(tt);
  const SiteOut so = sites(S);
  const int step_count = __float_as_int(S->st[S_STEP]) + 1;
  const int total = __float_as_int(S->st[S_TOTAL]) + 1;
  float reward;
  bool succ, trunc;
  if (A.task == 0) {
    // single_arm.py:322-380; float32 cube_pos compared against float64 bin bounds
    const double cx = (double)so.cube.x, cy = (double)so.cube.y;
    const bool over_bin = (c_m.bin_min[0] < cx && cx < c_m.bin_max[0]) && (c_m.bin_min[1] < cy && cy < c_m.bin_max[1]);
    bool inside = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float lower = __fsub_rn(comp(so.cube, k), c_m.cube_half), upper = __fadd_rn(comp(so.cube, k), c_m.cube_half);
      inside = inside && ((double)lower > c_m.bin_min[k]) && ((double)upper < c_m.bin_max[k]);
    }
    const bool released = inside && !touch_gripper;
    reward = 0.0f;
    if (touch_gripper) reward = 1.0f;
    if (touch_gripper && !touch_table) reward = 2.0f;
    if (over_bin) reward = 2.5f;
    if (inside) reward = 3.0f;
    if (released) reward = 4.0f;
    succ = reward == 4.0f;
    trunc = step_count >= c_m.max_episode_steps;
  } else {
    // env.py:341-358, float32, ((dx^2 + dy^2) + dz^2)
    const float dx = __fsub_rn(so.cube.x, S->st[S_GOAL]), dy = __fsub_rn(so.cube.y, S->st[S_GOAL + 1]),
                dz = __fsub_rn(so.cube.z, S->st[S_GOAL + 2]);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    succ = d < c_m.goal_threshold;
    reward = succ ? 0.0f : -1.0f;
    trunc = step_count >= c_m.goal_max_steps;
  }
  if (bad) { succ = false; trunc = true; reward = 0.0f; }
  const bool term = succ;
  t.sync();
  if (lane == 0) {
    S->st[S_STEP] = __int_as_float(step_count);
    S->st[S_TOTAL] = __int_as_float(total);
    if (ncon_raw > NC) diag[0] += 1u;
    if (bad) diag[2] += 1u;
    if (term || trunc) diag[3] += 1u;
    if (succ) diag[4] += 1u;
    if (A.reward) A.reward[env] = reward;
    if (A.terminated) A.terminated[env] = term ? 1 : 0;
    if (A.truncated) A.truncated[env] = trunc ? 1 : 0;
    if (A.success) A.success[env] = succ ? 1 : 0;
  }
  if (A.final_obs) write_obs(t, S, env, A.final_obs, nullptr, nullptr);
  if ((A.autoreset && (term || trunc)) || bad) {
    reset_env(t, S, A.env_offset + env, nullptr, A.task, A.seed_lo, A.seed_hi);
    kinematics(t, S);
  }
  write_obs(t, S, env, A.obs, A.achieved, A.desired);
  store_state(t, S, A.state, env);
}

template <unsigned LPE>
__global__ void __launch_bounds__(128)
reset_kernel(float* state, const uint8_t* mask, const float* box_pose, float* obs, float* achieved, float* desired, int n,
             int task, uint32_t seed_lo, uint32_t seed_hi, long long env_offset, DevTables T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int EPB = 128 / LPE;
  cg::thread_block blk = cg::this_thread_block();
  Tile<LPE> t = cg::tiled_partition<LPE>(blk);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  EnvS* S = reinterpret_cast<EnvS*>(smem_raw) + t.meta_group_rank();
  load_state(t, S, state, env);
  if (!mask || mask[env]) reset_env(t, S, env_offset + env, box_pose ? box_pose + (size_t)env * 7 : nullptr, task, seed_lo, seed_hi);
  kinematics(t, S);
  write_obs(t, S, env, obs, achieved, desired);
  store_state(t, S, state, env);
}

template <unsigned LPE> __global__ void __launch_bounds__(128) substeps_kernel(float* state, int n, int nsub, DevTables T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int EPB = 128 / LPE;
  cg::thread_block blk = cg::this_thread_block();
  Tile<LPE> t = cg::tiled_partition<LPE>(blk);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  EnvS* S = reinterpret_cast<EnvS*>(smem_raw) + t.meta_group_rank();
  load_state(t, S, state, env);
  for (int s = 0; s < nsub; s++) substep(t, S, T);
  store_state(t, S, state, env);
}

// mj_forward on the stored state: qacc, contacts (+ forces) and sites, nothing integrated
template <unsigned LPE>
__global__ void __launch_bounds__(128)
forward_kernel(const float* state, int n, float* qacc, int32_t* ncon_out, int32_t* con_geom, float* con_data, float* sites_out,
               DevTables T) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int EPB = 128 / LPE;
  cg::thread_block blk = cg::this_thread_block();
  Tile<LPE> t = cg::tiled_partition<LPE>(blk);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  EnvS* S = reinterpret_cast<EnvS*>(smem_raw) + t.meta_group_rank();
  const int lane = t.thread_rank();
  load_state(t, S, state, env);
  kinematics(t, S);
  mass_matrix(t, S);
  t.sync();
  smooth_forces(t, S);
  const float qas = smooth_acc(t, S);
  collide_env(t, S, T);
  make_contact_rows(t, S, T);
  solve(t, S, T, qas, nullptr);
  t.sync();
  const int ncon = min(S->ncon, NC);
  if (qacc && lane < NV) qacc[(size_t)env * NV + lane] = S->a[lane];
  if (ncon_out && lane == 0) ncon_out[env] = S->ncon;
  for (int c = lane; c < NC; c += LPE) {
    const bool live = c < ncon;
    if (con_geom) {
      int g1 = -1, g2 = -1;
      if (live) { const DevPair& P = T.pair[S->cpair[c]]; g1 = T.geom[P.g1].mjid; g2 = T.geom[P.g2].mjid; }
      con_geom[((size_t)env * NC + c) * 2] = g1;
      con_geom[((size_t)env * NC + c) * 2 + 1] = g2;
    }
    if (con_data) {
      float* d = con_data + ((size_t)env * NC + c) * 11;
      d[0] = live ? S->cdist[c] : 0.0f;
      for (int k = 0; k < 3; k++) { d[1 + k] = live ? S->cpos[c][k] : 0.0f; d[4 + k] = live ? S->cnrm[c][k] : 0.0f; }
      for (int k = 0; k < 4; k++) d[7 + k] = live ? S->cfrc[c][k] : 0.0f;
    }
  }
  if (sites_out) {
    const SiteOut so = sites(S);
    for (int k = lane; k < 9; k += LPE) {
      float v = k < 3 ? comp(so.cube, k) : (k < 6 ? c_m.bin_center[k - 3] : comp(so.ee, k - 6));
      sites_out[(size_t)env * 9 + k] = v;
    }
  }
}

// env.py:346-349 on a batch
__global__ void compute_reward_kernel(const float* ag, const float* dg, long long n, float thr, float* out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float dx = __fsub_rn(ag[3 * i], dg[3 * i]), dy = __fsub_rn(ag[3 * i + 1], dg[3 * i + 1]),
              dz = __fsub_rn(ag[3 * i + 2], dg[3 * i + 2]);
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
  out[i] = d < thr ? 0.0f : -1.0f;
}

// record <-> user arrays.  dir = 0: record -> arrays (get), 1: arrays -> record (set)
__global__ void state_io_kernel(float* state, int n, int dir, float* qpos, float* qvel, float* ctrl, float* warm, float* goal,
                                int32_t* step_count, int32_t* total_steps, uint32_t* episode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int env = i / STATE_WORDS, k = i % STATE_WORDS;
  if (env >= n) return;
  float* rec = state + (size_t)env * STATE_WORDS + k;
  float* p = nullptr;
  if (k < S_QVEL) { if (qpos) p = qpos + (size_t)env * 13 + k; }
  else if (k < S_CTRL) { if (qvel) p = qvel + (size_t)env * 12 + (k - S_QVEL); }
  else if (k < S_WARM) { if (ctrl) p = ctrl + (size_t)env * 6 + (k - S_CTRL); }
  else if (k < S_GOAL) { if (warm) p = warm + (size_t)env * 12 + (k - S_WARM); }
  else if (k < S_STEP) { if (goal) p = goal + (size_t)env * 3 + (k - S_GOAL); }
  else if (k == S_STEP) { if (step_count) p = reinterpret_cast<float*>(step_count) + env; }
  else if (k == S_TOTAL) { if (total_steps) p = reinterpret_cast<float*>(total_steps) + env; }
  else if (k == S_EPISODE) { if (episode) p = reinterpret_cast<float*>(episode) + env; }
  if (!p) return;
  if (dir == 0) *p = *rec; else *rec = *p;
}

__global__ void init_state_kernel(float* state, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int env = i / STATE_WORDS, k = i % STATE_WORDS;
  if (env >= n) return;
  float v = 0.0f;
  if (k == S_QPOS + 9) v = 1.0f;
  state[(size_t)env * STATE_WORDS + k] = v;
}

// sums the per-env uint32 counters into out[8] (uint64)
__global__ void diag_reduce_kernel(const float* state, int n, unsigned long long* out) {
  __shared__ unsigned long long acc[SO100_NDIAG_K];
  if (threadIdx.x < SO100_NDIAG_K) acc[threadIdx.x] = 0;
  __syncthreads();
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < n; env += gridDim.x * blockDim.x) {
    const uint32_t* d = reinterpret_cast<const uint32_t*>(state + (size_t)env * STATE_WORDS + S_DIAG);
    for (int k = 0; k < SO100_NDIAG_K; k++)
      if (d[k]) atomicAdd(&acc[k], (unsigned long long)d[k]);
  }
  __syncthreads();
  if (threadIdx.x < SO100_NDIAG_K) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

}  // namespace so100
