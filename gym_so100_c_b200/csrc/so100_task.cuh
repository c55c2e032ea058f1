// Task layer on the device: episode reset (utils.py:18-29, single_arm.py:299-309, env.py:322-334), sites and
// observations (single_arm.py:82-114, env.py:137-145), staged reward / success / truncation
// (single_arm.py:322-380, env.py:172-182, 341-358, 372-406), same-call auto-reset.
#pragma once
#include "so100_solve.cuh"

namespace so100 {

struct StepArgs {
  float* state;            // [N, STATE_WORDS]
  const float* action;     // [N,6]
  float *obs, *achieved, *desired, *reward, *final_obs;
  uint8_t *terminated, *truncated, *success;
  float* ep_return;        // [N] or null: return of the episode an env has just finished (untouched otherwise)
  int32_t* ep_length;      // [N] or null: its length in env steps
  int n, autoreset, task;
  int trace;               // development builds (-DSO100_TRACE): base record id
  uint32_t seed_lo, seed_hi;
  long long env_offset;
};

template <unsigned LPE>
__device__ void reset_env(const Tile<LPE>& t, TaskS* S, long long gid, const float* box_pose, int task, uint32_t seed_lo,
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
    S->st[S_EPRET] = 0.0f;
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
__device__ __forceinline__ SiteOut sites(const FrameBlock& f) {
  SiteOut o;
  o.cube = ld3(f.lpos[NL]) + mulmv(f.lmat[NL], ld3(c_m.cube_site_off));
  o.ee = ld3(f.lpos[4]) + mulmv(f.lmat[4], ld3(c_m.ee_off));
  return o;
}

template <unsigned LPE>
__device__ __forceinline__ void write_obs(const Tile<LPE>& t, const TaskS* S, int env, float* obs, float* achieved, float* desired) {
  const int lane = t.thread_rank();
  const SiteOut so = sites(S->f);
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

// reward / success / termination / observation / same-call auto-reset on the post-step state whose frames and
// contact list are in workspace record `w`
template <unsigned LPE> __device__ void task_env(const Tile<LPE>& t, TaskS* S, const StepArgs& A, float* w, int env, const DevTables& T) {
  const int lane = t.thread_rank();
  uint32_t* diag = reinterpret_cast<uint32_t*>(&S->st[S_DIAG]);
  const int ncon_raw = __float_as_int(w[W_HDR]), ncon = min(ncon_raw, NC);
  bool bad = false;
  for (int k = lane; k < S_GOAL; k += LPE) bad |= !isfinite(S->st[k]);
  bad = t.any(bad);
  int tg = 0, tt = 0;
  for (int c = lane; c < ncon; c += LPE) {
    const DevPair& P = T.pair[__float_as_int(w[W_CON + c * CON_WORDS + 7])];
    if ((P.g2 == c_m.cg_cube && ((c_m.pad_mask >> P.g1) & 1u)) || (P.g1 == c_m.cg_cube && ((c_m.pad_mask >> P.g2) & 1u))) tg = 1;
    if (P.g1 == c_m.cg_cube && P.g2 == c_m.cg_table) tt = 1;     // ordered pair ("red_box", "table")
  }
  const bool touch_gripper = t.any(tg), touch_table = t.any(tt);
  const SiteOut so = sites(S->f);
  const int step_count = __float_as_int(S->st[S_STEP]) + 1;
  const int total = __float_as_int(S->st[S_TOTAL]) + 1;
  float reward;
  bool succ, trunc;
  if (A.task == 0) {
    // float32 cube_pos compared against float64 bin bounds, exactly as numpy does in the reference
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
  } else if (A.task >= 2) {
    // SO100TouchCubeTask (2, single_arm.py:149-215) / SO100TouchCubeSparseTask (3, single_arm.py:246-285); the reference
    // evaluates this in float64 on float64 sites, here float32 sites and arithmetic (reward within 1e-5)
    const V3 d3 = so.ee - so.cube;
    const float d = sqrtf(dot(d3, d3));
    reward = 0.0f;
    if (A.task == 2) {
      if (d < 0.7f) reward = fmaxf(reward, 0.1f * (1.0f - d / 0.7f));
      if (d < 0.5f) reward = fmaxf(reward, 0.2f * (1.0f - d / 0.5f));
      if (d < 0.3f) reward = fmaxf(reward, 0.5f * (1.0f - d / 0.3f));
      if (d < 0.1f) reward = fmaxf(reward, 1.0f * (1.0f - d / 0.1f));
      if (d < 0.05f) reward = fmaxf(reward, 2.0f * (1.0f - d / 0.05f));
      if (touch_gripper) reward += 1.0f;
    }
    succ = touch_gripper && d < 0.05f;
    reward = succ ? 4.0f : reward - 0.2f;
    trunc = step_count >= c_m.goal_max_steps;   // TimeLimit 300 (__init__.py:7,17)
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
    if (ncon_raw > NC || (__float_as_int(w[W_HDR + 2]) & HDR_OVERFLOW)) diag[0] += 1u;
    if (bad) diag[2] += 1u;
    if (term || trunc) diag[3] += 1u;
    if (succ) diag[4] += 1u;
    // RecordEpisodeStatistics (scripts/train_sac.py:290, train_sac_her.py:226): running return of the episode; on its end the
    // return / length go to the per-env outputs and into the per-env sums that so100_episode_stats reduces
    const float epret = S->st[S_EPRET] + reward;
    if (term || trunc) {
      S->st[S_RETSUM] += epret;
      S->st[S_LENSUM] = __uint_as_float(__float_as_uint(S->st[S_LENSUM]) + (uint32_t)step_count);
      if (A.ep_return) A.ep_return[env] = epret;
      if (A.ep_length) A.ep_length[env] = step_count;
      S->st[S_EPRET] = 0.0f;
    } else {
      S->st[S_EPRET] = epret;
    }
    if (A.reward) A.reward[env] = reward;
    if (A.terminated) A.terminated[env] = term ? 1 : 0;
    if (A.truncated) A.truncated[env] = trunc ? 1 : 0;
    if (A.success) A.success[env] = succ ? 1 : 0;
  }
  if (A.final_obs) write_obs(t, S, env, A.final_obs, nullptr, nullptr);
  if ((A.autoreset && (term || trunc)) || bad) {
    reset_env(t, S, A.env_offset + env, nullptr, A.task, A.seed_lo, A.seed_hi);
    kinematics<false>(t, S);
    // the contact list in the workspace belongs to the finished episode: the next step's first substep must not reuse it
    if (lane == 0) reinterpret_cast<int*>(w + W_HDR)[3] = HDR_STALE;
  }
  write_obs(t, S, env, A.obs, A.achieved, A.desired);
}

}  // namespace so100
