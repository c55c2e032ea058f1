// The step pipeline: a short sequence of small kernels per physics substep, linked by an L2-resident
// workspace record per env (so100_scratch.cuh), with the rare heavy work classes pulled out of the
// regular grid into queue-driven persistent kernels:
//
//   per substep:  K1  phase_kin_dyn       kinematics, mass matrix, RNE bias, actuators      state -> work
//                 K2a phase_collide_box   broad phase, box SAT + clipping, hull-pair cull   work  -> work (+ hull queue)
//                 K2b phase_collide_hull  GJK/EPA for queued envs (~14 % of envs)           work  -> work (+ heavy queue)
//                 K3l phase_solve_light   contact rows, Newton, Euler; envs <= 8 contacts   state, work -> state
//                 K3h phase_solve_heavy   the same for queued envs with 9..24 contacts or an arm-cube contact (dense Hessian)
//   per step:     K1 (kinematics only) + K2a + K2b + K4 phase_task (reward / flags / obs / auto-reset)
//
// Why not one fused kernel: ncu on the fused step (profiles/r01_fused_step_ncu.txt) showed it to be instruction-fetch
// bound (356 KB of SASS, stall_no_instruction = 10 warps per issue); small kernels keep every resident warp in the same
// few KB of code, and per-phase scratch layouts (1.2-4 KB per env instead of 10.4 KB) lift the shared-memory limit on
// resident warps.  Why queues: GJK/EPA and many-contact solves take 5-20x the time of the common case; run inside
// the regular grid they pin a whole block (and its registers / shared memory) while three of its four tiles idle.
#pragma once
#include "so100_gjk.cuh"
#include "so100_task.cuh"

namespace so100 {

#ifndef SO100_TPB_K3L
#define SO100_TPB_K3L 32      // threads per block of the light solve kernel: one warp per block, so a slow env pins at most its warp sibling
#endif
#ifndef SO100_WARPS_K3L
#define SO100_WARPS_K3L 16    // resident warps per SM the register allocation of the light solve kernel must allow (128 registers with
                              // two envs per warp: no spills; 20 warps = 96 registers spills 132 bytes and is 4 % slower)
#endif

// K1 for one env: state -> frames (+ mass matrix, smooth forces, unconstrained acceleration) in the workspace
template <unsigned LPE>
__device__ __forceinline__ void kin_dyn_env(const Tile<LPE>& t, KinS* S, float* rec, float* w, const float* action, int env, int with_dyn) {
  const int lane = t.thread_rank();
  copy_vec<LPE, 32>(t, S->st, rec);       // qpos qvel ctrl (+ 1 word of warm)
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
  kinematics<true>(t, S);
  if (with_dyn) {
    mass_matrix(t, S);
    smooth_forces(t, S);
    smooth_acc(t, S);
    copy_vec<LPE, W_DYN_N>(t, w + W_DYN, reinterpret_cast<const float*>(&S->d));
  }
  copy_vec<LPE, W_FRAMES_N>(t, w + W_FRAMES, reinterpret_cast<const float*>(&S->f));
}

// first kernel of a stage: re-arm the NEXT stage's copy of the stage words; slow-lane words (`stage` = 0 for the first stage of a
// step, which also re-arms the slow lane; every stage notes where the slow-lane queue stands: the entries a stage adds belong to
// that stage's slow-lane kernel)
__device__ __forceinline__ void rearm_queues(const Queues& Q, int stage) {
  if (blockIdx.x != 0) return;
  if (threadIdx.x < Q_WORDS) Q.ctl_next[threadIdx.x] = 0;
  if (threadIdx.x == Q_WORDS) {
    if (stage == 0) Q.lanectl[Q_LANE_COUNT] = 0;
    if (stage >= 0 && stage < 12) Q.lanectl[Q_LANE_CURSOR + stage] = stage == 0 ? 0 : Q.lanectl[Q_LANE_COUNT];
  }
}

// K1: regular grid over a group's envs
template <unsigned LPE>
__global__ void __launch_bounds__(128) phase_kin_dyn(float* state, float* work, const float* action, int n, int with_dyn, Queues Q, int stage) {
  SO100_TILE_PROLOGUE(LPE, 128, KinS);
  SO100_TRACE_SCOPE(Q.trace + TR_KIN);
  rearm_queues(Q, stage);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  if (Q.slowlane) {
    if (stage == 0) { if (lane == 0) Q.lane[env] = 0; }
    else if (Q.lane[env] != 0) return;
  }
  kin_dyn_env(t, S, state + (size_t)env * STATE_WORDS, work + (size_t)env * WORK_WORDS, action, env, with_dyn);
}

// K2a: frames -> box contacts + surviving hull pairs; classifies the env
// `reuse`: the workspace already holds the complete contact lists of exactly this state (the trailing collision stage of
// the previous so100_step), except for envs whose header says HDR_STALE (reset since): everyone else only re-enters the
// heavy queue, which K1 has just re-armed
// K2a for one env whose frames are in S->f
template <unsigned LPE>
__device__ __forceinline__ void collide_box_route(const Tile<LPE>& t, BoxS* S, float* w, int env, const DevTables& T, const Queues& Q) {
  const int lane = t.thread_rank();
  int ncon;
  bool coupled;
  const int nsurv = collide_box_env(t, S, w, T, &ncon, &coupled);
  if (nsurv > 0) {
    int base = 0;
    if (lane == 0) base = atomicAdd(&Q.ctl[Q_HULL_COUNT], nsurv);
    base = t.shfl(base, 0);
    if (lane < nsurv) Q.hull[base + lane] = env * NHP + lane;
  } else if (lane == 0) {
    Q.route(env, ncon, coupled, false);
  }
}

template <unsigned LPE> __global__ void __launch_bounds__(128) phase_collide_box(float* work, int n, DevTables T, Queues Q, int reuse) {
  SO100_TILE_PROLOGUE(LPE, 128, BoxS);
  SO100_TRACE_SCOPE(Q.trace + TR_BOX);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  if (Q.suspended(env)) return;
  float* w = work + (size_t)env * WORK_WORDS;
  if (reuse) {
    const int4 hdr = *reinterpret_cast<const int4*>(w + W_HDR);
    if (hdr.w == 0) {
      if (lane == 0) Q.route(env, hdr.x, (hdr.z & HDR_COUPLED) != 0, hdr.y > 0);
      return;
    }
  }
  copy_vec<LPE, W_FRAMES_N>(t, reinterpret_cast<float*>(&S->f), w + W_FRAMES);
  t.sync();
  collide_box_route(t, S, w, env, T, Q);
}

// K1 + K2a in one kernel (both run two envs per warp): the frames go from the kinematics scratch to the collision scratch through
// registers instead of through the workspace and a kernel boundary.  The queue words K2a appends to were re-armed by the previous
// stage (rearm_queues), so no block has to wait for block 0.
constexpr size_t KINBOX_SMEM = sizeof(KinS) > sizeof(BoxS) ? sizeof(KinS) : sizeof(BoxS);
template <unsigned LPE>
__global__ void __launch_bounds__(128) phase_kin_box(float* state, float* work, const float* action, int n, int with_dyn, DevTables T, Queues Q,
                                                     int stage, int reuse) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int EPB = 128 / LPE;
  cg::thread_block blk = cg::this_thread_block();
  Tile<LPE> t = cg::tiled_partition<LPE>(blk);
  unsigned char* smem = smem_raw + (size_t)t.meta_group_rank() * KINBOX_SMEM;
  const int lane = t.thread_rank();
  SO100_TRACE_SCOPE(Q.trace + TR_KIN);
  rearm_queues(Q, stage);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  if (Q.slowlane) {
    if (stage == 0) { if (lane == 0) Q.lane[env] = 0; }
    else if (Q.lane[env] != 0) return;
  }
  float* w = work + (size_t)env * WORK_WORDS;
  KinS* K = reinterpret_cast<KinS*>(smem);
  kin_dyn_env(t, K, state + (size_t)env * STATE_WORDS, w, action, env, with_dyn);
  if (reuse) {
    const int4 hdr = *reinterpret_cast<const int4*>(w + W_HDR);
    if (hdr.w == 0) {
      if (lane == 0) Q.route(env, hdr.x, (hdr.z & HDR_COUPLED) != 0, hdr.y > 0);
      return;
    }
  }
  t.sync();
  BoxS* B = reinterpret_cast<BoxS*>(smem);
  {
    // the two layouts overlap in the scratch: every lane holds its part of the frame block in registers before anyone stores
    constexpr int NV4 = W_FRAMES_N / 4, R = (NV4 + LPE - 1) / LPE;
    const float4* src = reinterpret_cast<const float4*>(&K->f);
    float4* dst = reinterpret_cast<float4*>(&B->f);
    float4 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) if (lane + r * (int)LPE < NV4) v[r] = src[lane + r * LPE];
    t.sync();
#pragma unroll
    for (int r = 0; r < R; r++) if (lane + r * (int)LPE < NV4) dst[lane + r * LPE] = v[r];
  }
  t.sync();
  collide_box_route(t, B, w, env, T, Q);
}

// K2b: persistent tiles drain the hull-pair queue
#ifndef SO100_K2B_MINB
#define SO100_K2B_MINB 5
#endif
template <unsigned LPE> __global__ void __launch_bounds__(128, SO100_K2B_MINB) phase_collide_hull(float* work, DevTables T, Queues Q) {
  SO100_TILE_PROLOGUE(LPE, 128, HullS);
  SO100_TRACE_SCOPE(Q.trace + TR_HULL);
#ifdef SO100_DEV_SKIP_HULL
  return;      // timing experiment only (wrong physics): what the step costs without any GJK/EPA work
#endif
  const int count = *reinterpret_cast<volatile int*>(&Q.ctl[Q_HULL_COUNT]);
  if (blockIdx.x == 0 && threadIdx.x == 0) Q.note(1, count);
  // every tile's first item is its own index (no 4096 atomics on one word before anyone starts); the rest is handed out dynamically
  const int ntiles = gridDim.x * (128 / LPE);
  int i = blockIdx.x * (128 / LPE) + t.meta_group_rank();
  const int max_gjk = Q.slowlane ? Q.budget_gjk : 48, max_epa = Q.slowlane ? Q.budget_epa : EPA_MAXV - 4;
  for (;;) {
    if (i >= count) break;
    const int item = Q.hull[i], env = item / NHP;
    float* w = work + (size_t)env * WORK_WORDS;
    copy_vec<LPE, W_FRAMES_N>(t, reinterpret_cast<float*>(&S->f), w + W_FRAMES);
    if (lane == 0) i = ntiles + atomicAdd(&Q.ctl[Q_HULL_NEXT], 1);     // next item: in flight during this one
    t.sync();
    bool coupled = false;
    const int ncon = collide_hull_item(t, S, w, item % NHP, T, &coupled, max_gjk, max_epa);     // one call site: the item code is 60 KB
    if (lane == 0 && ncon >= 0) Q.route(env, ncon, coupled, true);
    if (lane == 0 && ncon == -2) Q.suspend(env);        // over budget: the slow lane redoes this env's collision stage
    i = t.shfl(i, 0);
    t.sync();
  }
}

// forward-mode outputs of the solve kernels (so100_forward): nothing is integrated or stored
struct SolveOut {
  float* qacc;       // [N,12] or null
  float* con_data;   // [N,NC,11] or null: forces go to columns 7..10
  int forward;       // 0: integrate and store the state (the step path)
};

// `active` = false (only with several tiles per warp): the tile has no env to solve but its warp sibling does; it loads a valid
// record, takes part in the warp votes of the solver loop with "done", and stores nothing
// `max_it`: Newton budget (regular light kernel with the slow lane on); an over-budget solve returns -1 and stores nothing.
template <bool DENSE, unsigned LPE, class ES>
__device__ int solve_env(const Tile<LPE>& t, ES* S, float* rec, const float* w, int env, int ncon_raw, const DevTables& T,
                         const SolveOut& O, bool active = true, int max_it = NEWTON_MAXIT) {
  const int lane = t.thread_rank();
  copy_vec<LPE, 48>(t, S->st, rec);
  copy_vec<LPE, W_FRAMES_N>(t, reinterpret_cast<float*>(&S->f), w + W_FRAMES);
  copy_vec<LPE, W_DYN_N>(t, reinterpret_cast<float*>(&S->d), w + W_DYN);
  const int ncon = min(ncon_raw, ES::NCAP);
  for (int k = lane; k < ncon * 2; k += LPE)
    reinterpret_cast<float4*>(&S->con[0][0])[k] = reinterpret_cast<const float4*>(w + W_CON)[k];
  if (lane == 0) S->ncon = ncon;
  t.sync();
  make_contact_rows(t, S, T);
#ifdef SO100_SOLVE_TRACE
  const int iters = solve<DENSE>(t, S, T, (O.forward || !active) ? nullptr : reinterpret_cast<uint32_t*>(rec + S_DIAG), active, max_it, O.forward ? env : -1);
#else
  const int iters = solve<DENSE>(t, S, T, (O.forward || !active) ? nullptr : reinterpret_cast<uint32_t*>(rec + S_DIAG), active, max_it);
#endif
  if (!active || iters < 0) return iters;
  if (O.forward) {
    t.sync();
    if (O.qacc && lane < NV) O.qacc[(size_t)env * NV + lane] = S->a[lane];
    if (O.con_data)
      for (int k = lane; k < ncon * 4; k += LPE) O.con_data[((size_t)env * NC + (k >> 2)) * 11 + 7 + (k & 3)] = S->cfrc[k >> 2][k & 3];
    return iters;
  }
  integrate(t, S);
  copy_vec<LPE, 48>(t, rec, S->st);   // qpos qvel (ctrl unchanged) warm (+ goal / counters unchanged)
  return iters;
}

// K3l: regular grid, one tile per env; envs with more than NCL contacts or an arm-cube contact are left to K3h
template <unsigned LPE>
__global__ void __launch_bounds__(SO100_TPB_K3L, SO100_WARPS_K3L * 32 / SO100_TPB_K3L) phase_solve_light(float* state, const float* work, int n, DevTables T, Queues Q, SolveOut O) {
  SO100_TILE_PROLOGUE(LPE, SO100_TPB_K3L, SolS<NCL>);
  SO100_TRACE_SCOPE(Q.trace + TR_LIGHT_A);
  int slot = blockIdx.x * EPB + t.meta_group_rank();
  bool live = slot < n;
  if (LPE == 16 && Q.pair_far) {
    // the two tiles of warp W take slots W and n - 1 - W: paired stragglers run their data-dependent branches (line-search trip counts,
    // cone zones) one after the other, 7.8 us per Newton iteration instead of ~6 for a straggler whose sibling is done after one
    const int W = slot >> 1;                  // EPB is even: bit 0 of the slot is the tile's position in its warp
    const bool second = (slot & 1) != 0;
    slot = second ? n - 1 - W : W;
    live = second ? slot > W : 2 * W <= n - 1;      // an odd n's middle env belongs to the first tile
  }
  if (LPE == 32 && !live) return;            // one tile per warp: nobody to vote with
  const int env = Q.order_in[live ? slot : n - 1];
  const float* w = work + (size_t)env * WORK_WORDS;
  const int4 hdr = *reinterpret_cast<const int4*>(w + W_HDR);
  // hdr.y = hull pairs of the env (written by K2a only; K2b, which may be running beside this kernel, never touches it): such
  // envs are solved by the queue kernel below once K2b has completed their contact list.  hdr.x / hdr.z of the others are final.
  const bool hull_env = Q.split && hdr.y > 0;
  const int ncon_raw = (hdr.z & HDR_COUPLED) ? NC + 2 : hdr.x;      // arm-cube contacts: heavy kernel (dense Hessian)
  const bool mine = live && !hull_env && ncon_raw <= NCL && !Q.suspended(env);
  const int max_it = Q.slowlane ? Q.budget_newton : NEWTON_MAXIT;
  int iters = hull_env ? 0 : 1000;           // heavy / medium envs count as slow; hull envs are not this kernel's business
#ifdef SO100_SOLVE_CLOCK
  unsigned long long t0_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_));
#endif
#ifdef SO100_SOLVE_CLOCK
  if (lane < 4) S->clk2[lane] = 0;
#endif
  // several tiles per warp: every tile enters the solver (its loop votes warp-wide), the ones without work as inactive
  if (LPE < 32) { const int it_ = solve_env<false>(t, S, state + (size_t)env * STATE_WORDS, w, env, min(ncon_raw, NCL), T, O, mine, max_it); if (mine) iters = it_; }
  else if (mine) iters = solve_env<false>(t, S, state + (size_t)env * STATE_WORDS, w, env, ncon_raw, T, O, true, max_it);
  if (iters < 0) {                           // over the Newton budget: nothing was stored; the slow lane takes the env from here
    if (lane == 0) Q.suspend(env);
    iters = 1000;
  }
#ifdef SO100_SOLVE_CLOCK
  // development build: duration (ns) and iteration count of this env's last solve in the spare words of its state record
  unsigned long long t1_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
  if (lane == 0) {
    float* rec_ = state + (size_t)env * STATE_WORDS;
    rec_[57] = __int_as_float((int)(t1_ - t0_));
    rec_[58] = __int_as_float(iters);
    rec_[59] = __int_as_float(ncon_raw | (S->coupled << 8) | (S->clk2[3] << 12));     // bits 12..: line-search iterations of the solve
    for (int k_ = 0; k_ < 4; k_++) rec_[60 + k_] = __int_as_float(S->clk[k_]);
  }
#endif
  if (!O.forward && live && lane == 0) {
    const int pos = iters >= 3 ? atomicAdd(&Q.ctl[Q_SLOW], 1) : n - 1 - atomicAdd(&Q.ctl[Q_FAST], 1);
    Q.order_out[pos] = env;
  }
}

// K2b + solve, fused (schedule 3): the tile that completes an env's last hull pair goes straight on to solve that env -- the light
// class on a 16-lane sub-tile with an inactive sibling (the same code and tile width as the regular light grid), an arm-cube
// contact with the dense 8-contact code -- instead of queueing it for a kernel behind a grid-wide boundary.  For the envs with
// hull pairs (the ones with the long solves) the stage then costs the slowest env's OWN GJK/EPA + solve, not the slowest
// GJK/EPA item of the group plus the slowest solve of the group.  Envs with more than NCL contacts still go to the heavy queue.
// After the hull-pair queue the tiles drain light queue b / medium queue b (filled only when the first stage of a step reuses
// the previous step's contact lists).
constexpr size_t hs_max2(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t HULL_SOLVE_SMEM = hs_max2(sizeof(HullS), 2 * sizeof(SolS<NCL>));
template <unsigned LPE_LIGHT>
__device__ __forceinline__ void hull_solve_env(const Tile<32>& t, const Tile<LPE_LIGHT>& tl, unsigned char* smem, float* state, const float* work,
                                               int env, int ncon, bool coupled, const DevTables& T, const Queues& Q, const SolveOut& O) {
  float* rec = state + (size_t)env * STATE_WORDS;
  const float* w = work + (size_t)env * WORK_WORDS;
  if (ncon > NCL) {
    if (t.thread_rank() == 0) Q.heavy[atomicAdd(&Q.ctl[Q_HEAVY_COUNT], 1)] = env;
  } else if (coupled) {
    solve_env<true>(t, reinterpret_cast<SolS<NCL>*>(smem), rec, w, env, ncon, T, O);
  } else if constexpr (LPE_LIGHT < 32) {
    const int half = tl.meta_group_rank() & 1;
    solve_env<false>(tl, reinterpret_cast<SolS<NCL>*>(smem) + half, rec, w, env, ncon, T, O, half == 0);
  } else {
    solve_env<false>(t, reinterpret_cast<SolS<NCL>*>(smem), rec, w, env, ncon, T, O);
  }
  t.sync();
}
#ifndef SO100_K2BS_MINB
#define SO100_K2BS_MINB 3
#endif
template <unsigned LPE_LIGHT>
__global__ void __launch_bounds__(128, SO100_K2BS_MINB) phase_hull_solve(float* state, float* work, DevTables T, Queues Q, SolveOut O) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::thread_block blk = cg::this_thread_block();
  Tile<32> t = cg::tiled_partition<32>(blk);
  Tile<LPE_LIGHT> tl = cg::tiled_partition<LPE_LIGHT>(blk);
  unsigned char* smem = smem_raw + (size_t)t.meta_group_rank() * HULL_SOLVE_SMEM;
  HullS* S = reinterpret_cast<HullS*>(smem);
  const int lane = t.thread_rank();
  SO100_TRACE_SCOPE(Q.trace + TR_HULL);
  const int count = *reinterpret_cast<volatile int*>(&Q.ctl[Q_HULL_COUNT]);
  if (blockIdx.x == 0 && threadIdx.x == 0) Q.note(1, count);
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.ctl[Q_HULL_NEXT], 1);
    i = t.shfl(i, 0);
    if (i >= count) break;
    const int item = Q.hull[i], env = item / NHP;
    float* w = work + (size_t)env * WORK_WORDS;
    copy_vec<32, W_FRAMES_N>(t, reinterpret_cast<float*>(&S->f), w + W_FRAMES);
    t.sync();
    bool coupled = false;
    const int ncon = collide_hull_item(t, S, w, item % NHP, T, &coupled);
    t.sync();
    if (ncon >= 0) hull_solve_env<LPE_LIGHT>(t, tl, smem, state, work, env, ncon, coupled, T, Q, O);
  }
  // the queues the reuse path of K2a fills for envs with hull pairs
  const int nlb = *reinterpret_cast<volatile int*>(&Q.ctl[Q_LB_COUNT]);
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.ctl[Q_LB_NEXT], 1);
    i = t.shfl(i, 0);
    if (i >= nlb) break;
    const int env = Q.light_b[i];
    hull_solve_env<LPE_LIGHT>(t, tl, smem, state, work, env, __float_as_int(work[(size_t)env * WORK_WORDS + W_HDR]), false, T, Q, O);
  }
  const int nmb = *reinterpret_cast<volatile int*>(&Q.ctl[Q_MEDB_COUNT]);
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.ctl[Q_MEDB_NEXT], 1);
    i = t.shfl(i, 0);
    if (i >= nmb) break;
    const int env = Q.medium_b[i];
    hull_solve_env<LPE_LIGHT>(t, tl, smem, state, work, env, __float_as_int(work[(size_t)env * WORK_WORDS + W_HDR]), true, T, Q, O);
  }
}

// K3l-b: the light-class envs whose contact list K2b completed (light queue b), same solver code and tile width as the regular
// grid above.  Persistent warps; the tiles of a warp pull their items independently but enter the solver together (its loop
// votes warp-wide), a tile that found the queue empty as inactive.
template <unsigned LPE>
__global__ void __launch_bounds__(SO100_TPB_K3L, SO100_WARPS_K3L * 32 / SO100_TPB_K3L) phase_solve_light_queue(float* state, const float* work, DevTables T, Queues Q, SolveOut O) {
  SO100_TILE_PROLOGUE(LPE, SO100_TPB_K3L, SolS<NCL>);
  SO100_TRACE_SCOPE(Q.trace + TR_LIGHT_B);
  const int count = *reinterpret_cast<volatile int*>(&Q.ctl[Q_LB_COUNT]);
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.ctl[Q_LB_NEXT], 1);
    i = t.shfl(i, 0);
    const bool have = i < count;
    if (!warp_any<LPE>(t, have)) break;
    const int env = Q.light_b[have ? i : 0];
    const float* w = work + (size_t)env * WORK_WORDS;
    if (LPE < 32) solve_env<false>(t, S, state + (size_t)env * STATE_WORDS, w, env, min(__float_as_int(w[W_HDR]), NCL), T, O, have);
    else if (have) solve_env<false>(t, S, state + (size_t)env * STATE_WORDS, w, env, __float_as_int(w[W_HDR]), T, O);
    t.sync();
  }
}

// K3h / K3m: persistent tiles drain the heavy queue (NCAP = NC: more than NCL contacts) or the medium queue (NCAP = NCL: an
// arm-cube contact among at most NCL), both with the dense Hessian.  The medium instantiation carries one contact row per
// lane instead of three: fewer registers (more resident tiles when a policy holds thousands of cubes at once) and shorter
// Newton iterations for the coupled stragglers that end every group's solve stage.
#ifndef SO100_K3H_MINB
#define SO100_K3H_MINB 1      // resident blocks per SM the heavy solve kernel's register allocation must allow (168 registers, 3
                              // blocks; measured on B200: 4 (128 registers, spills) -8 %, 5 (96 registers) -14 % env-steps/s)
#endif
#ifndef SO100_K3M_MINB
#define SO100_K3M_MINB 4      // the same for the medium instantiation: 128 registers without spills since the Cholesky factors moved out of local
                              // memory (round 2; 3 blocks = 160 registers: +1.5 % step time when a policy holds thousands of cubes)
#endif
// `which` (medium instantiation only): 0 = medium queue a (complete after K2a), 1 = medium queue b (complete after K2b)
template <unsigned LPE, int NCAP>
__global__ void __launch_bounds__(128, NCAP == NCL ? SO100_K3M_MINB : SO100_K3H_MINB) phase_solve_heavy(float* state, const float* work, DevTables T, Queues Q, SolveOut O, int which) {
  SO100_TILE_PROLOGUE(LPE, 128, SolS<NCAP>);
  constexpr bool MED = NCAP == NCL;
  SO100_TRACE_SCOPE(Q.trace + (MED ? (which ? TR_MED_B : TR_MED_A) : TR_HEAVY));
  const int* queue = MED ? (which ? Q.medium_b : Q.medium_a) : Q.heavy;
  const int qc = MED ? (which ? Q_MEDB_COUNT : Q_MEDA_COUNT) : Q_HEAVY_COUNT;
  const int count = *reinterpret_cast<volatile int*>(&Q.ctl[qc]);
  if (MED && blockIdx.x == 0 && threadIdx.x == 0) Q.note(0, count);
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.ctl[qc + 1], 1);
    i = t.shfl(i, 0);
    if (i >= count) break;
    const int env = queue[i];
    const float* w = work + (size_t)env * WORK_WORDS;
#ifdef SO100_SOLVE_CLOCK
    unsigned long long t0_, t1_;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0_));
    if (lane < 4) S->clk2[lane] = 0;
    const int iters_ = solve_env<true>(t, S, state + (size_t)env * STATE_WORDS, w, env, __float_as_int(w[W_HDR]), T, O);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1_));
    if (lane == 0) {       // same record as the light kernel writes; bit 11: dense kernel
      float* rec_ = state + (size_t)env * STATE_WORDS;
      rec_[57] = __int_as_float((int)(t1_ - t0_));
      rec_[58] = __int_as_float(iters_);
      rec_[59] = __int_as_float(min(__float_as_int(w[W_HDR]), 255) | ((S->coupled ? 1 : 0) << 8) | (1 << 11) | (S->clk2[3] << 12));
#if SO100_SOLVE_CLOCK == 2
      // split of the dense direction: Hessian + factorisation total, then assembly / arm block + W / Schur complement + solves
      rec_[60] = __int_as_float(S->clk[2]);
      for (int k_ = 0; k_ < 3; k_++) rec_[61 + k_] = __int_as_float(S->clk2[k_]);
#else
      for (int k_ = 0; k_ < 4; k_++) rec_[60 + k_] = __int_as_float(S->clk[k_]);
#endif
    }
#else
    solve_env<true>(t, S, state + (size_t)env * STATE_WORDS, w, env, __float_as_int(w[W_HDR]), T, O);
#endif
    t.sync();
  }
}

// Slow lane: persistent one-warp blocks take the envs suspended during stage `stage` (slow-lane queue entries from this
// stage's cursor up to the queue length at launch) and run each of them, alone, through the rest of the step: the collision
// stage of `stage` again (cheap, and it makes every suspension reason look the same), solve + integrate, then for every later
// stage kinematics / dynamics -> collision -> solve -> integrate, and finally the trailing position stage.  The same device
// functions with the same tile widths as the regular kernels run here (kinematics on a 16-lane sub-tile, the light solve on a
// 16-lane sub-tile with an inactive sibling), so an env's result does not depend on which lane computed it.
#ifndef SO100_SLOW_MINB
#define SO100_SLOW_MINB 8
#endif
constexpr size_t slow_max2(size_t a, size_t b) { return a > b ? a : b; }
constexpr size_t SLOW_SMEM = slow_max2(slow_max2(sizeof(KinS), sizeof(BoxS)), slow_max2(slow_max2(sizeof(HullS), 2 * sizeof(SolS<NCL>)), sizeof(SolS<NC>)));
template <unsigned LPE_KIN, unsigned LPE_BOX, unsigned LPE_LIGHT>
__global__ void __launch_bounds__(32, SO100_SLOW_MINB) phase_slow_lane(float* state, float* work, DevTables T, Queues Q, int stage, int nsub, int trailing) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::thread_block blk = cg::this_thread_block();
  Tile<32> t = cg::tiled_partition<32>(blk);
  Tile<LPE_KIN> tk = cg::tiled_partition<LPE_KIN>(blk);
  Tile<LPE_LIGHT> tl = cg::tiled_partition<LPE_LIGHT>(blk);
  Tile<LPE_BOX> tb = cg::tiled_partition<LPE_BOX>(blk);
  const int lane = t.thread_rank();
  SO100_TRACE_SCOPE(Q.trace + TR_SLOW);
  const int count = *reinterpret_cast<volatile int*>(&Q.lanectl[Q_LANE_COUNT]);
  if (blockIdx.x == 0 && threadIdx.x == 0) Q.note(0, count >> 2);      // a quarter of the envs suspended so far this step
  const SolveOut O{nullptr, nullptr, 0};
  for (;;) {
    int i = 0;
    if (lane == 0) i = atomicAdd(&Q.lanectl[Q_LANE_CURSOR + stage], 1);
    i = t.shfl(i, 0);
    if (i >= count) break;
    const int env = Q.slow[i];
    float* rec = state + (size_t)env * STATE_WORDS;
    float* w = work + (size_t)env * WORK_WORDS;
    for (int s = stage; s <= nsub; s++) {
      const bool position_only = s == nsub;           // the trailing mj_step1 of the step
      if (position_only && !trailing) break;
      if (s > stage) {
        if (tk.meta_group_rank() == 0) kin_dyn_env(tk, reinterpret_cast<KinS*>(smem_raw), rec, w, nullptr, env, position_only ? 0 : 1);
        t.sync();
      }
      // collision stage from the frames in the workspace
      BoxS* B = reinterpret_cast<BoxS*>(smem_raw);
      copy_vec<32, W_FRAMES_N>(t, reinterpret_cast<float*>(&B->f), w + W_FRAMES);
      t.sync();
      int ncon = 0, nsurv = 0;
      bool coupled = false;
      if (tb.meta_group_rank() == 0) nsurv = collide_box_env(tb, B, w, T, &ncon, &coupled);     // on a sub-tile of the regular kernel's width
      t.sync();
      nsurv = t.shfl(nsurv, 0); ncon = t.shfl(ncon, 0); coupled = t.shfl((int)coupled, 0) != 0;
      if (nsurv > 0) {
        HullS* H = reinterpret_cast<HullS*>(smem_raw);     // its frame block is the one just loaded (both layouts start with it)
        for (int slot = 0; slot < nsurv; slot++) {
          bool cpl = false;
          const int r = collide_hull_item(t, H, w, slot, T, &cpl);
          if (r >= 0) { ncon = r; coupled = cpl; }
          t.sync();
        }
      }
      if (position_only) break;
      if (ncon > NCL) solve_env<true>(t, reinterpret_cast<SolS<NC>*>(smem_raw), rec, w, env, ncon, T, O);
      else if (coupled) solve_env<true>(t, reinterpret_cast<SolS<NCL>*>(smem_raw), rec, w, env, ncon, T, O);
      else if constexpr (LPE_LIGHT < 32)
        solve_env<false>(tl, reinterpret_cast<SolS<NCL>*>(smem_raw) + tl.meta_group_rank(), rec, w, env, ncon, T, O, tl.meta_group_rank() == 0);
      else solve_env<false>(t, reinterpret_cast<SolS<NCL>*>(smem_raw), rec, w, env, ncon, T, O);
      t.sync();
    }
  }
}

// K4: task layer on the post-step state
template <unsigned LPE> __global__ void __launch_bounds__(128) phase_task(StepArgs A, float* work, DevTables T) {
  SO100_TILE_PROLOGUE(LPE, 128, TaskS);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= A.n) return;
  SO100_TRACE_SCOPE(A.trace + TR_TASK);
  float* w = work + (size_t)env * WORK_WORDS;
  float* rec = A.state + (size_t)env * STATE_WORDS;
  copy_vec<LPE, STATE_WORDS>(t, S->st, rec);
  copy_vec<LPE, W_FRAMES_N>(t, reinterpret_cast<float*>(&S->f), w + W_FRAMES);
  t.sync();
  task_env(t, S, A, w, env, T);
  t.sync();
  copy_vec<LPE, STATE_WORDS>(t, rec, S->st);
}

template <unsigned LPE>
__global__ void __launch_bounds__(128)
reset_kernel(float* state, const uint8_t* mask, const float* box_pose, float* obs, float* achieved, float* desired, int n,
             int task, uint32_t seed_lo, uint32_t seed_hi, long long env_offset) {
  SO100_TILE_PROLOGUE(LPE, 128, TaskS);
  const int env = blockIdx.x * EPB + t.meta_group_rank();
  if (env >= n) return;
  float* rec = state + (size_t)env * STATE_WORDS;
  copy_vec<LPE, STATE_WORDS>(t, S->st, rec);
  t.sync();
  if (!mask || mask[env]) reset_env(t, S, env_offset + env, box_pose ? box_pose + (size_t)env * 7 : nullptr, task, seed_lo, seed_hi);
  kinematics<false>(t, S);
  write_obs(t, S, env, obs, achieved, desired);
  t.sync();
  copy_vec<LPE, STATE_WORDS>(t, rec, S->st);
}

// so100_forward: contact list and sites of the workspace in the caller's layout (forces are filled in by the solve kernels)
__global__ void export_forward_kernel(const float* work, int n, int32_t* ncon_out, int32_t* con_geom, float* con_data, float* sites_out,
                                      DevTables T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int env = i / 32, c = i % 32;
  if (env >= n) return;
  const float* w = work + (size_t)env * WORK_WORDS;
  const int ncon_raw = __float_as_int(w[W_HDR]), ncon = min(ncon_raw, NC);
  if (c == 0 && ncon_out) ncon_out[env] = ncon_raw;
  if (c < NC) {
    const bool live = c < ncon;
    const float* q = w + W_CON + c * CON_WORDS;
    if (con_geom) {
      int g1 = -1, g2 = -1;
      if (live) { const DevPair& P = T.pair[__float_as_int(q[7])]; g1 = T.geom[P.g1].mjid; g2 = T.geom[P.g2].mjid; }
      con_geom[((size_t)env * NC + c) * 2] = g1;
      con_geom[((size_t)env * NC + c) * 2 + 1] = g2;
    }
    if (con_data) {
      float* d = con_data + ((size_t)env * NC + c) * 11;
      d[0] = live ? q[6] : 0.0f;
      for (int k = 0; k < 6; k++) d[1 + k] = live ? q[k] : 0.0f;
      for (int k = 0; k < 4; k++) d[7 + k] = 0.0f;
    }
  } else if (c == NC && sites_out) {
    const FrameBlock& f = *reinterpret_cast<const FrameBlock*>(w + W_FRAMES);
    const SiteOut so = sites(f);
    float* o = sites_out + (size_t)env * 9;
    o[0] = so.cube.x; o[1] = so.cube.y; o[2] = so.cube.z;
    o[3] = c_m.bin_center[0]; o[4] = c_m.bin_center[1]; o[5] = c_m.bin_center[2];
    o[6] = so.ee.x; o[7] = so.ee.y; o[8] = so.ee.z;
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

// so100_measure_fp32_peak: 8 independent FFMA chains per thread, everything in registers
__global__ void __launch_bounds__(256) fma_peak_kernel(float* sink, int iters, float m) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const float c = 1e-3f;
#pragma unroll 4
  for (int i = 0; i < iters; i++) {
    a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
    a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
  }
  sink[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
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

// episode statistics over all envs: out[0] episodes finished, [1] successes, [2] sum of episode returns, [3] sum of episode lengths
__global__ void episode_reduce_kernel(const float* state, int n, double* out) {
  __shared__ double acc[4];
  if (threadIdx.x < 4) acc[threadIdx.x] = 0.0;
  __syncthreads();
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int env = blockIdx.x * blockDim.x + threadIdx.x; env < n; env += gridDim.x * blockDim.x) {
    const float* rec = state + (size_t)env * STATE_WORDS;
    const uint32_t* d = reinterpret_cast<const uint32_t*>(rec + S_DIAG);
    a[0] += (double)d[3]; a[1] += (double)d[4];
    a[2] += (double)rec[S_RETSUM]; a[3] += (double)__float_as_uint(rec[S_LENSUM]);
  }
  for (int k = 0; k < 4; k++)
    if (a[k] != 0.0) atomicAdd(&acc[k], a[k]);
  __syncthreads();
  if (threadIdx.x < 4) atomicAdd(&out[threadIdx.x], acc[threadIdx.x]);
}

}  // namespace so100
