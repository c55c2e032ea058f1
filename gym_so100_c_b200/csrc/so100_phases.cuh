// Phase-split step pipeline.
//
// ncu on the single fused step kernel (profiles/r01_fused_step_ncu.txt) showed the dominant stall
// to be `no_instruction` (10.2 stalled warps per issue): one substep executes ~10 k distinct SASS
// instructions (~160 KB), and the resident warps of an SM sit in different regions of that code, so
// the instruction cache thrashes.  The step is therefore issued as a short sequence of small
// kernels -- all warps of a launch run the same few KB of code -- and the per-env intermediate
// data (link frames, mass matrix, smooth forces, contact list: <= 1.4 KB) is streamed through a
// workspace record that stays L2-resident (16384 envs x 1.4 KB = 23 MB of the 126 MB L2):
//
//   per substep:  K1 kinematics + mass matrix + RNE bias + actuators      state -> work
//                 K2 collision (broad phase, box SAT, hull GJK/EPA)       work  -> work
//                 K3 contact rows + Newton solve + Euler                  state, work -> state
//   per step:     K1 (kinematics only) + K2 + K4 task layer (reward / flags / obs / auto-reset)
#pragma once
#include "so100_kernels.cuh"

namespace so100 {

// workspace record (float words)
constexpr int W_FRAMES = 0;                 // lpos[7][3] lmat[7][9] axis[6][3]  (contiguous in EnvS)
constexpr int W_FRAMES_N = 21 + 63 + 18;    // 102
constexpr int W_DYN = W_FRAMES + W_FRAMES_N;  // Marm[21] qfs[12]              (contiguous in EnvS)
constexpr int W_DYN_N = 21 + 12;
constexpr int W_QAS = W_DYN + W_DYN_N;      // qacc_smooth[12]
constexpr int W_NCON = W_QAS + NV;
constexpr int W_CON = W_NCON + 1;           // contact c: pos[3] nrm[3] dist pair
constexpr int WORK_WORDS = ((W_CON + 8 * NC + 31) / 32) * 32;

static_assert(offsetof(EnvS, lmat) == offsetof(EnvS, lpos) + 21 * sizeof(float), "frames must be contiguous");
static_assert(offsetof(EnvS, axis) == offsetof(EnvS, lpos) + 84 * sizeof(float), "frames must be contiguous");
static_assert(offsetof(EnvS, qfs) == offsetof(EnvS, Marm) + 21 * sizeof(float), "dyn must be contiguous");

template <unsigned LPE> __device__ __forceinline__ void copy_words(const Tile<LPE>& t, float* dst, const float* src, int n) {
  for (int k = t.thread_rank(); k < n; k += LPE) dst[k] = src[k];
}

template <unsigned LPE> __device__ __forceinline__ void store_contacts(const Tile<LPE>& t, const EnvS* S, float* w) {
  const int lane = t.thread_rank();
  const int ncon = min(S->ncon, NC);
  if (lane == 0) w[W_NCON] = __int_as_float(S->ncon);
  for (int k = lane; k < ncon * 8; k += LPE) {
    const int c = k >> 3, f = k & 7;
    float v;
    if (f < 3) v = S->cpos[c][f];
    else if (f < 6) v = S->cnrm[c][f - 3];
    else if (f == 6) v = S->cdist[c];
    else v = __int_as_float((int)S->cpair[c]);
    w[W_CON + k] = v;
  }
}
template <unsigned LPE> __device__ __forceinline__ void load_contacts(const Tile<LPE>& t, EnvS* S, const float* w) {
  const int lane = t.thread_rank();
  const int nraw = __float_as_int(w[W_NCON]);
  const int ncon = min(nraw, NC);
  if (lane == 0) S->ncon = nraw;
  for (int k = lane; k < ncon * 8; k += LPE) {
    const int c = k >> 3, f = k & 7;
    const float v = w[W_CON + k];
    if (f < 3) S->cpos[c][f] = v;
    else if (f < 6) S->cnrm[c][f - 3] = v;
    else if (f == 6) S->cdist[c] = v;
    else S->cpair[c] = (unsigned char)__float_as_int(v);
  }
}

#define SO100_PHASE_PROLOGUE(LPE_)                                             \
  extern __shared__ __align__(16) unsigned char smem_raw[];                    \
  constexpr int EPB = 128 / LPE_;                                              \
  cg::thread_block blk = cg::this_thread_block();                              \
  Tile<LPE_> t = cg::tiled_partition<LPE_>(blk);                               \
  const int env = blockIdx.x * EPB + t.meta_group_rank();                      \
  if (env >= n) return;                                                        \
  EnvS* S = reinterpret_cast<EnvS*>(smem_raw) + t.meta_group_rank();           \
  const int lane = t.thread_rank();                                            \
  (void)lane

// K1: state -> frames (+ mass matrix, smooth forces, unconstrained acceleration)
template <unsigned LPE>
__global__ void __launch_bounds__(128) phase_kin_dyn(float* state, float* work, const float* action, int n, int with_dyn) {
  SO100_PHASE_PROLOGUE(LPE);
  float* rec = state + (size_t)env * STATE_WORDS;
  float* w = work + (size_t)env * WORK_WORDS;
  copy_words(t, S->st, rec, S_GOAL);       // qpos qvel ctrl warm
  t.sync();
  if (action) {
    // before_step: unnormalize_so100 in float32 (constants.py:44-47, 78-86)
    if (lane < NL) {
      float v = __fadd_rn(action[(size_t)env * 6 + lane], 1.0f);
      v = __fdiv_rn(v, 2.0f);
      v = __fmul_rn(v, c_m.act_range[lane]);
      v = __fadd_rn(v, c_m.act_lo[lane]);
      v = fminf(fmaxf(v, c_m.act_lo[lane]), c_m.act_hi[lane]);
      S->st[S_CTRL + lane] = v;
      rec[S_CTRL + lane] = v;
    }
    t.sync();
  }
  kinematics(t, S);
  if (with_dyn) {
    mass_matrix(t, S);
    t.sync();
    smooth_forces(t, S);
    const float qas = smooth_acc(t, S);
    if (lane < NV) w[W_QAS + lane] = qas;
    copy_words(t, w + W_DYN, S->Marm, W_DYN_N);
  }
  copy_words(t, w + W_FRAMES, &S->lpos[0][0], W_FRAMES_N);
}

// K2: frames -> contact list
template <unsigned LPE> __global__ void __launch_bounds__(128) phase_collide(float* work, int n, DevTables T) {
  SO100_PHASE_PROLOGUE(LPE);
  float* w = work + (size_t)env * WORK_WORDS;
  copy_words(t, &S->lpos[0][0], w + W_FRAMES, W_FRAMES_N);
  t.sync();
  collide_env(t, S, T);
  store_contacts(t, S, w);
}

// K3: contact rows, Newton solve, semi-implicit Euler
#ifndef SO100_MINB_K3
#define SO100_MINB_K3 4
#endif
template <unsigned LPE> __global__ void __launch_bounds__(128, SO100_MINB_K3) phase_solve(float* state, const float* work, int n, DevTables T) {
  SO100_PHASE_PROLOGUE(LPE);
  float* rec = state + (size_t)env * STATE_WORDS;
  const float* w = work + (size_t)env * WORK_WORDS;
  copy_words(t, S->st, rec, STATE_WORDS);
  copy_words(t, &S->lpos[0][0], w + W_FRAMES, W_FRAMES_N);
  copy_words(t, S->Marm, w + W_DYN, W_DYN_N);
  load_contacts(t, S, w);
  const float qas = lane < NV ? w[W_QAS + lane] : 0.0f;
  t.sync();
  make_contact_rows(t, S, T);
  solve(t, S, T, qas, reinterpret_cast<uint32_t*>(&S->st[S_DIAG]));
  integrate(t, S);
  // qpos qvel (ctrl unchanged) warm + diagnostics
  copy_words(t, rec, S->st, S_GOAL);
  for (int k = S_DIAG + lane; k < S_DIAG + SO100_NDIAG_K; k += LPE) rec[k] = S->st[k];
}

// K4: reward / success / termination / observation / same-call auto-reset on the post-step state
// (single_arm.py:322-380, env.py:137-145, 172-182, 372-406)
template <unsigned LPE> __global__ void __launch_bounds__(128) phase_task(StepArgs A, const float* work, DevTables T) {
  const int n = A.n;
  SO100_PHASE_PROLOGUE(LPE);
  const float* w = work + (size_t)env * WORK_WORDS;
  load_state(t, S, A.state, env);
  copy_words(t, &S->lpos[0][0], w + W_FRAMES, W_FRAMES_N);
  load_contacts(t, S, w);
  t.sync();
  uint32_t* diag = reinterpret_cast<uint32_t*>(&S->st[S_DIAG]);
  const int ncon_raw = S->ncon, ncon = min(ncon_raw, NC);
  bool bad = false;
  for (int k = lane; k < S_GOAL; k += LPE) bad |= !isfinite(S->st[k]);
  bad = t.any(bad);
  int tg = 0, tt = 0;
  for (int c = lane; c < ncon; c += LPE) {
    const DevPair& P = T.pair[S->cpair[c]];
    if ((P.g2 == c_m.cg_cube && ((c_m.pad_mask >> P.g1) & 1u)) || (P.g1 == c_m.cg_cube && ((c_m.pad_mask >> P.g2) & 1u))) tg = 1;
    if (P.g1 == c_m.cg_cube && P.g2 == c_m.cg_table) tt = 1;     // ordered pair ("red_box", "table")
  }
  const bool touch_gripper = t.any(tg), touch_table = t.any(tt);
  const SiteOut so = sites(S);
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
  } else {
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

}  // namespace so100
