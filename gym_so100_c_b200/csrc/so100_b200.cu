// Host side of libso100_b200.so: model narrowing/validation, device allocation, kernel
// launches behind the C ABI of include/so100_b200.h.  No CPU fallback exists: every entry
// point either launches the sm_100a kernels or fails with an error code.
#include <array>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/so100_b200.h"
#include "../../include/so100_model.h"

namespace so100 { constexpr int SO100_NDIAG_K = SO100_NDIAG; }
#include "so100_phases.cuh"
#include "so100_her.cuh"
#include "so100_render.cuh"

using namespace so100;

static_assert(NC == SO100_MAX_CONTACTS, "contact capacity mismatch");

// lanes per env (tile width) of each phase kernel
#ifndef SO100_LPE_K1
#define SO100_LPE_K1 16
#endif
#ifndef SO100_LPE_K3L
#define SO100_LPE_K3L 16     // two envs per warp in the light solve kernel (12 dof lanes + contact rows fit 16 lanes): measured on B200
                             // against one env per warp, both at their best register budget: +3 % env-steps/s at 16384 envs, +9.5 % at 65536
#endif
#ifndef SO100_LPE_K2A
#define SO100_LPE_K2A 16     // two envs per warp in the box collision stage too (one lane per geom / candidate pair; the 24 clipping
                             // candidates of a box pair in two passes)
#endif
constexpr unsigned LPE_K1 = SO100_LPE_K1, LPE_K2A = SO100_LPE_K2A, LPE_K2B = 32, LPE_K3L = SO100_LPE_K3L, LPE_K3H = 32, LPE_K4 = 32;
constexpr int BLOCK = 128, TPB_K3L = SO100_TPB_K3L;
// grids of the queue-driven persistent kernels (4 tiles per block): large enough for one item per tile in a 2048-env group,
// small enough that the blocks that find the queue empty do not crowd the SMs (the heavy solve kernel holds 27 k registers
// and 58 KB of shared memory per block)
#ifndef SO100_K2B_BLOCKS_PER_SM
#define SO100_K2B_BLOCKS_PER_SM 1     // and at least n / 16 blocks (one tile per 4 envs; 0.18 GJK items per env under random actions).
                                      // Measured with 6 groups at 16384 envs: 2 per SM or n / 8 (the earlier setting) 2.55 ms per
                                      // step, 1 per SM or n / 16 2.51, 1/2 per SM or n / 32 2.55; the scripted-grasp phase prefers
                                      // the larger grid by 5 %
#endif
#ifndef SO100_K3H_BLOCKS
#define SO100_K3H_BLOCKS 37           // heavy queue (> 8 contacts): 148 tiles; rare under any policy
#endif
#ifndef SO100_K3M_BLOCKS
#define SO100_K3M_BLOCKS 148          // medium queue (arm-cube contact): 0-10 envs per 2048-env group and substep under random
                                      // actions, half the group when a policy holds the cubes
#endif
enum { CLS_KIN = 0, CLS_BOX = 1, CLS_SOLVE = 2, CLS_TASK = 3, CLS_HULL = 4, CLS_HEAVY = 5, CLS_N = 6 };

static thread_local std::string g_err;
static std::mutex g_model_mutex;
// The uniform model constants live in __constant__ memory, and kernel attributes (opt-in shared memory) are per device:
// both are tracked per device index, so that one process may drive several GPUs (one handle each, or more with one model).
constexpr int MAX_DEVICES = 64;
static int g_live_handles[MAX_DEVICES] = {0};
static so100::DevModel g_live_model[MAX_DEVICES];
static bool g_configured[MAX_DEVICES] = {false};
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

// Every entry point runs on its handle's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1, dev = -1;
  explicit DeviceGuard(int d) : dev(d) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};
#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      (void)cudaGetLastError();   /* a non-sticky error (e.g. out of memory) must not resurface in the next call's check */ \
      return fail(SO100_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));             \
    }                                                                                              \
  } while (0)

// A contiguous range of envs that runs the step pipeline on its own stream.  Kernel durations are set by their slowest
// env (a 25-iteration Newton solve or a GJK/EPA pair is ~10x the typical one), so one grid over all envs leaves the GPU
// nearly idle during every kernel's tail; with several independent groups the tail of one group's kernel overlaps the
// bulk of another's.  The heavy solve queue of each group drains on a second stream beside the light kernel.
struct EnvGroup {
  int off = 0, n = 0;
  int index = 0, stage = 0;   // development builds (-DSO100_TRACE): trace record id = (index * 12 + stage) * 10 + kernel kind
  // st == nullptr: the caller's stream.  side: medium queue a (beside the light grid), hull: K2b -> light queue b,
  // side2: medium queue b + heavy queue (after K2b)
  static constexpr int NSLOW = 12;
  cudaStream_t slow[NSLOW] = {nullptr};                 // slow-lane kernel of stage s runs on slow[s % NSLOW], joined before the task kernel
  cudaEvent_t slow_fork[NSLOW] = {nullptr}, slow_done[NSLOW] = {nullptr};
  cudaStream_t st = nullptr, side = nullptr, side2 = nullptr, hull = nullptr;
  cudaEvent_t fork = nullptr, fork2 = nullptr, join = nullptr, join2 = nullptr, join3 = nullptr, done = nullptr;
  cudaEvent_t t_done = nullptr;                         // timing-enabled twin of `done` (so100_group_times)
  cudaEvent_t staged = nullptr;                         // recorded after the group's first position stage of a step (stagger)
  int* ctl = nullptr;                                   // this group's queue control words
  int* order = nullptr;                                 // two longest-first permutations of the group's envs, [2][n_total] apart
  int parity = 0;                                       // which of the two the next solve stage reads
};

struct so100_ctx {
  int n = 0, device = 0, task = 0, nsub = 10;
  bool registered = false;    // counted in g_live_handles[device]
  uint64_t seed = 0;
  int64_t env_offset = 0;
  float* state = nullptr;
  DevGeom* geom = nullptr;
  DevPair* pair = nullptr;
  float4* vert = nullptr;
  uchar4* bpair = nullptr;
  unsigned long long* diag = nullptr;
  float* work = nullptr;      // [N, WORK_WORDS] phase-pipeline workspace (L2-resident)
  int* qmem = nullptr;        // queue control words + heavy queue [N] + hull-pair queue [N * NHP] + medium queues a, b [N] each + light queue b [N]
  int* order = nullptr;       // solve order permutations: [2][N] for the env groups, [2][N] for the whole-batch group
  // so100_step replays a CUDA graph of its whole launch sequence (all groups, forks and joins): the host cost of a step
  // drops from several hundred launch / event calls to one cudaGraphLaunch.  Actions are staged into a fixed buffer so
  // that the graph's kernel arguments never change; one graph is cached per distinct set of output pointers.
  // The cache is bounded (GRAPH_SETS pointer sets, least recently used evicted).  A caller that keeps rotating its output
  // buffers beyond that bound is switched to library-owned staging outputs: one more graph, keyed on those stable buffers and
  // never re-captured, followed by device-to-device copies into whatever pointers the call names.
  static constexpr int NKEY = 20, GRAPH_SETS = 8;
  struct StepGraph { const void* key[NKEY]; cudaGraphExec_t exec; uint64_t used; };
  std::vector<StepGraph> graphs;
  uint64_t graph_clock = 0;
  int graph_captures = 0;     // captures since create (so100_graph_stats)
  bool stage_outputs = false; // rotating caller pointers: run the graph on the staging outputs below
  float *s_obs = nullptr, *s_ag = nullptr, *s_dg = nullptr, *s_rew = nullptr, *s_fin = nullptr;
  uint8_t *s_term = nullptr, *s_trunc = nullptr, *s_succ = nullptr;
  // optional per-env episode outputs (so100_set_episode_outputs): return / length of the episode an env has just finished
  float* ep_return = nullptr;
  int32_t* ep_length = nullptr;
  std::vector<std::array<const void*, 8>> host_sets;   // page-locked destination sets of so100_step_host seen so far
  double* ep_stats = nullptr;  // device [4] scratch of so100_episode_stats
  // renderer (so100_render_config / so100_render)
  bool render_ready = false;
  RenderCfg rcfg;
  float4* r_planes = nullptr;
  int *r_adr = nullptr, *r_num = nullptr;
  float* r_rgb = nullptr;
  cudaStream_t cap = nullptr;
  float* act_stage = nullptr;
  bool use_graph = true;
  EnvGroup whole;             // all envs on the caller's stream (small batches, forward / substeps, timing mode)
  std::vector<EnvGroup> groups;
  cudaEvent_t ev_start = nullptr, t_start = nullptr;
  bool group_times = false, capturing = false;   // SO100_GROUP_TIMES=1: per-group completion times of the last step
  // Groups that start a step together stay in phase (all in the solve bulk, then all in its tail) and overlap little;
  // stagger = 1 starts every odd group only after its even neighbour has finished its first position stage, so that the
  // two are about a third of a substep apart for the rest of the step; 2 chains all groups that way.
  // slow lane (so100_scratch.cuh: Queues): on by default for the small grid class; budgets of the regular kernels
  bool fuse_k12 = true;       // SO100_FUSE_K12=0: K1 and K2a as two kernels
  bool pair_far = true;       // SO100_PAIR_FAR=0: the light solve kernel pairs neighbours of the longest-first order in a warp
  bool slowlane_enabled = false, slow_on = false;   // SO100_SLOWLANE=1 enables it.  Measured (B200, 16384 envs): 3.6-4.9 ms per step against
                                                    // 2.5 without: one warp taking an env through kinematics, box collision, its hull pairs one
                                                    // after the other and the solve needs 250-450 us per stage, longer than the regular
                                                    // pipeline's whole stage, and the 255-register lane warps crowd the regular kernels
  int budget_newton = 6, budget_gjk = 10, budget_epa = 5;
  int prio_low = 0, prio_mid = 0, prio_high = 0;   // kernel scheduling priorities (launch_p); all equal with SO100_PRIO=0
  int dag = 0;       // SO100_DAG: schedule of a substep's kernels (launch_solve_stage); 0 = K2b before all solve classes (measured
                     // fastest: the envs K2b completes are the ones with the long solves, so running K2b beside the light grid
                     // (1, 2) only moves those solves behind a second kernel boundary: 3.3 vs 2.5 ms per step at 16384 envs)
  int stagger = 0;   // measured on B200: 1 and 2 are 1-9 % slower than 0 at 4096 / 16384 / 65536 envs (the groups drift apart on their own)
  int sm_count = 148;
  // Grid class of the queue kernels.  Their best grids depend on the workload: under random actions the queues hold a handful
  // of envs and small grids win (resident-but-idle persistent blocks cost the other groups' kernels), while a policy that holds
  // thousands of cubes at once fills them and wants twice the tiles.  Every step copies the largest relative queue lengths it saw
  // (Queues::note) to pinned host memory; so100_step reads the last completed copy, without synchronising, and replays the graph
  // captured for that class.  Results do not depend on the grids, so a stale reading costs time, never correctness.
  int* qstat = nullptr;             // device [2]
  volatile int* qstat_host = nullptr;   // pinned [2]
  int grid_class = 0;               // 0 small grids, 1 large grids
  bool adaptive_grids = true;       // SO100_ADAPTIVE_GRIDS=0 pins class 0
  int k2b_blocks = 148 * SO100_K2B_BLOCKS_PER_SM, k2b_div = 16;   // SO100_K2B_BLOCKS / SO100_K2B_DIV (environment) override
  int k3h_blocks = SO100_K3H_BLOCKS, k3m_blocks = SO100_K3M_BLOCKS;   // SO100_K3H_BLOCKS / SO100_K3M_BLOCKS (environment) override
  // so100_step ends with a collision stage on the post-step state (mj_step1), and the next so100_step starts with one on the
  // same state: while nothing else has touched the state in between, the first substep reuses those contact lists
  // (bit-identical by construction; envs auto-reset by the task kernel are marked stale and recomputed).  SO100_REUSE=0 disables.
  bool work_fresh = false, reuse_enabled = true;
  bool timing = false;        // so100_phase_timing: CUDA-event pairs around every phase-kernel launch
  std::vector<std::pair<cudaEvent_t, int>> events;   // (event, kernel class) begin markers, class -1 = end marker
  // staging for the host-buffer entry point
  float *h_action = nullptr, *h_obs = nullptr, *h_ag = nullptr, *h_dg = nullptr, *h_rew = nullptr, *h_fin = nullptr;
  uint8_t *h_term = nullptr, *h_trunc = nullptr, *h_succ = nullptr;
  DevTables tables() const { return DevTables{geom, pair, vert, bpair}; }
  static constexpr int CTL_WORDS = Q_STRIDE * 40;
  static size_t qmem_words(size_t n) { return CTL_WORDS + (6 + NHP) * n; }
  Queues queues(const EnvGroup& G) const {
    int* base = qmem + CTL_WORDS;
    const size_t N = (size_t)n;
    return Queues{G.ctl + (G.stage & 1) * Q_PARITY, G.ctl + ((G.stage + 1) & 1) * Q_PARITY, G.ctl + Q_LANE_BASE, base + N + (size_t)G.off * NHP, base + G.off, base + (1 + NHP) * N + G.off, base + (2 + NHP) * N + G.off,
                  base + (3 + NHP) * N + G.off, G.order + (size_t)G.parity * n, G.order + (size_t)(1 - G.parity) * n, qstat,
                  (1024 * 1024) / std::max(G.n, 1), (G.index < 8 || G.index == 32) ? ((G.index & 7) * 12 + G.stage) * 10 : -1000000, dag != 0 ? 1 : 0,
                  base + (4 + NHP) * N + G.off, base + (5 + NHP) * N + G.off, slow_on ? 1 : 0, budget_newton, budget_gjk, budget_epa, pair_far ? 1 : 0};
  }
};

// ------------------------------------------------------------------ small host math (double)
namespace {
struct D3 { double x, y, z; };
struct DQ { double w, x, y, z; };
DQ qmul_h(DQ a, DQ b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
void q2mat_h(DQ q, double* m) {
  double w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
D3 rot_h(const double* m, D3 v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
}
double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// frame of body b relative to `root` (an ancestor reached through joint-less bodies): pos + quat
bool rel_frame(const so100_model& m, int b, int root, D3& pos, DQ& q) {
  pos = {0, 0, 0}; q = {1, 0, 0, 0};
  while (b != root) {
    if (b <= 0 || m.body_jtype[b] >= 0) return false;
    double R[9];
    DQ qb = {m.body_quat[b][0], m.body_quat[b][1], m.body_quat[b][2], m.body_quat[b][3]};
    q2mat_h(qb, R);
    D3 p = rot_h(R, pos);
    pos = {m.body_pos[b][0] + p.x, m.body_pos[b][1] + p.y, m.body_pos[b][2] + p.z};
    q = qmul_h(qb, q);
    b = m.body_parent[b];
  }
  return true;
}
}  // namespace

// Narrow the packed fp64 model to the kernel's specialised float32 tables.
static int build_dev_model(const so100_model& m, DevModel& dm, std::vector<DevGeom>& geoms, std::vector<DevPair>& pairs,
                           std::vector<float4>& verts) {
  memset(&dm, 0, sizeof(dm));
  if (m.nv != NV || m.nq != NQ || m.nu != NL) return fail(SO100_ERR_MODEL, "kernel is specialised for nq=13 nv=12 nu=6");
  if (m.ngeom != NGEOM || m.npair > NPAIR_MAX || m.npair > 255) return fail(SO100_ERR_MODEL, "geom/pair count does not fit");
  // chain: hinge bodies in dof order, each the child of the previous one; one free body with dofs 6..11
  int link_body[NL], cube_body = -1, link_of_body[SO100_MAXBODY];
  for (int b = 0; b < SO100_MAXBODY; b++) link_of_body[b] = -2;
  for (int l = 0; l < NL; l++) link_body[l] = -1;
  for (int b = 1; b < m.nbody; b++) {
    if (m.body_jtype[b] == SO100_JNT_HINGE) {
      int d = m.body_dofadr[b];
      if (d < 0 || d >= NL || m.body_qposadr[b] != d) return fail(SO100_ERR_MODEL, "hinge dofs must be 0..5");
      link_body[d] = b; link_of_body[b] = d;
    } else if (m.body_jtype[b] == SO100_JNT_FREE) {
      if (m.body_dofadr[b] != NL || m.body_qposadr[b] != NL || m.body_parent[b] != 0) return fail(SO100_ERR_MODEL, "free body must own dofs 6..11");
      cube_body = b; link_of_body[b] = NL;
    }
  }
  if (cube_body < 0) return fail(SO100_ERR_MODEL, "no free body");
  for (int b = 0; b < m.nbody; b++)
    if (m.body_weldid[b] == 0) link_of_body[b] = -1;   // static
  D3 rp; DQ rq;
  for (int l = 0; l < NL; l++) {
    int b = link_body[l];
    if (b < 0) return fail(SO100_ERR_MODEL, "missing hinge link");
    int parent = m.body_parent[b];
    if (l == 0) {
      if (m.body_weldid[parent] != 0) return fail(SO100_ERR_MODEL, "first link must hang off a static body");
      if (!rel_frame(m, parent, 0, rp, rq)) return fail(SO100_ERR_MODEL, "base frame");
      dm.base_pos[0] = (float)rp.x; dm.base_pos[1] = (float)rp.y; dm.base_pos[2] = (float)rp.z;
      dm.base_quat[0] = (float)rq.w; dm.base_quat[1] = (float)rq.x; dm.base_quat[2] = (float)rq.y; dm.base_quat[3] = (float)rq.z;
    } else if (parent != link_body[l - 1]) {
      return fail(SO100_ERR_MODEL, "links must form a serial chain");
    }
    for (int k = 0; k < 3; k++) {
      dm.link_pos[l][k] = (float)m.body_pos[b][k];
      dm.link_axis[l][k] = (float)m.body_jaxis[b][k];
      dm.link_ipos[l][k] = (float)m.body_ipos[b][k];
    }
    for (int k = 0; k < 4; k++) dm.link_quat[l][k] = (float)m.body_quat[b][k];
    dm.link_mass[l] = (float)m.body_mass[b];
    // body-frame inertia tensor Ri diag Ri^T
    double Ri[9], I[9];
    q2mat_h({m.body_iquat[b][0], m.body_iquat[b][1], m.body_iquat[b][2], m.body_iquat[b][3]}, Ri);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += Ri[r * 3 + k] * m.body_inertia[b][k] * Ri[c * 3 + k];
        I[r * 3 + c] = s;
      }
    dm.link_Ib[l][0] = (float)I[0]; dm.link_Ib[l][1] = (float)I[1]; dm.link_Ib[l][2] = (float)I[2];
    dm.link_Ib[l][3] = (float)I[4]; dm.link_Ib[l][4] = (float)I[5]; dm.link_Ib[l][5] = (float)I[8];
    dm.armature[l] = (float)m.dof_armature[l];
    dm.lim_lo[l] = m.dof_limited[l] ? (float)m.dof_range[l][0] : -1e30f;
    dm.lim_hi[l] = m.dof_limited[l] ? (float)m.dof_range[l][1] : 1e30f;
    dm.lim_invw[l] = (float)m.dof_invweight0[l];
    if (m.act_dof[l] != l) return fail(SO100_ERR_MODEL, "actuator a must drive dof a");
    dm.kp[l] = (float)m.act_kp[l]; dm.kv[l] = (float)m.act_kv[l];
    dm.ctrl_lo[l] = (float)m.act_ctrlrange[l][0]; dm.ctrl_hi[l] = (float)m.act_ctrlrange[l][1];
    dm.frc_lo[l] = (float)m.act_forcerange[l][0]; dm.frc_hi[l] = (float)m.act_forcerange[l][1];
    dm.start_pose[l] = (float)m.start_pose[l];
    dm.act_lo[l] = (float)m.act_lo[l]; dm.act_hi[l] = (float)m.act_hi[l];
    dm.act_range[l] = (float)(m.act_hi[l] - m.act_lo[l]);
  }
  {
    int b = cube_body;
    for (int k = 0; k < 3; k++)
      if (m.body_ipos[b][k] != 0) return fail(SO100_ERR_MODEL, "free body needs its CoM at the origin");
    if (fabs(m.body_iquat[b][0]) != 1.0) return fail(SO100_ERR_MODEL, "free body needs a principal-axes frame");
    dm.cube_mass = (float)m.body_mass[b];
    for (int k = 0; k < 3; k++) dm.cube_I[k] = (float)m.body_inertia[b][k];
    for (int d = NL; d < NV; d++)
      if (m.dof_armature[d] != 0) return fail(SO100_ERR_MODEL, "armature on the free joint");
  }
  dm.timestep = (float)m.timestep;
  dm.gx = (float)m.gravity[0]; dm.gy = (float)m.gravity[1]; dm.gz = (float)m.gravity[2];
  dm.impratio = (float)m.impratio;
  dm.inv_scale = (float)(1.0 / (m.meaninertia * NV));
  dm.nsub = m.nsubstep; dm.npair = m.npair; dm.ngeom = m.ngeom;
  dm.max_episode_steps = m.max_episode_steps; dm.goal_max_steps = 300; dm.curriculum_steps = m.goal_curriculum_steps;
  // frictionloss / limit rows: default solref (0.02,1), solimp (0.9,0.95,0.001,0.5,2)
  const double tc = std::max(0.02, 2 * m.timestep), dmax = 0.95, d0 = 0.9;
  for (int d = 0; d < NV; d++) {
    double R = std::max(1e-15, (1 - d0) / d0 * m.dof_invweight0[d]);
    dm.fr_R[d] = (float)R; dm.fr_D[d] = (float)(1.0 / R); dm.fr_floss[d] = (float)m.dof_frictionloss[d];
  }
  dm.fr_B = (float)(2.0 / (dmax * tc));
  dm.lim_B = dm.fr_B;
  dm.lim_K = (float)(1.0 / (dmax * dmax * tc * tc));
  const float dsi[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.0f};
  memcpy(dm.lim_solimp, dsi, sizeof(dsi));

  // static body world frames
  auto world_frame = [&](int b, D3& p, DQ& q) { return rel_frame(m, b, 0, p, q); };
  // geoms
  geoms.assign(NGEOM, DevGeom{});
  verts.clear();
  for (int g = 0; g < NGEOM; g++) {
    DevGeom& G = geoms[g];
    int b = m.geom_body[g];
    G.link = link_of_body[b];
    if (G.link == -2) return fail(SO100_ERR_MODEL, "collidable geom on a welded child body");
    G.mjid = m.geom_mjid[g];
    G.rbound = (float)m.geom_rbound[g];
    double Rg[9];
    q2mat_h({m.geom_quat[g][0], m.geom_quat[g][1], m.geom_quat[g][2], m.geom_quat[g][3]}, Rg);
    bool ident = fabs(Rg[0] - 1) < 1e-12 && fabs(Rg[4] - 1) < 1e-12 && fabs(Rg[8] - 1) < 1e-12;
    if (!ident) return fail(SO100_ERR_MODEL, "geom frames must be aligned with their body");
    for (int k = 0; k < 3; k++) { G.center[k] = (float)m.geom_center[g][k]; G.half[k] = (float)m.geom_half[g][k]; }
    G.boxlike = m.geom_type[g] == SO100_GEOM_BOX;
    G.vadr = (int)verts.size(); G.vnum = m.geom_vnum[g];
    if (m.geom_type[g] == SO100_GEOM_MESH) {
      int hits = 0;
      for (int v = 0; v < G.vnum; v++) {
        const double* p = m.vert[m.geom_vadr[g] + v];
        verts.push_back(make_float4((float)p[0], (float)p[1], (float)p[2], 0.0f));
        bool corner = true;
        for (int k = 0; k < 3; k++) corner = corner && fabs(fabs(p[k] - m.geom_center[g][k]) - m.geom_half[g][k]) < 1e-9;
        hits += corner;
      }
      if (G.vnum == 8 && hits == 8) G.boxlike = 1;   // exact cuboid hull (the table): same support map as a box
    }
    if (G.link < 0) {
      if (!world_frame(b, rp, rq)) return fail(SO100_ERR_MODEL, "static geom frame");
      double Rw[9];
      q2mat_h(rq, Rw);
      D3 c = rot_h(Rw, {m.geom_center[g][0], m.geom_center[g][1], m.geom_center[g][2]});
      G.center[0] = (float)(rp.x + c.x); G.center[1] = (float)(rp.y + c.y); G.center[2] = (float)(rp.z + c.z);
      for (int k = 0; k < 9; k++) G.wmat[k] = (float)Rw[k];
      G.org[0] = (float)rp.x; G.org[1] = (float)rp.y; G.org[2] = (float)rp.z;
    }
  }
  // pairs
  pairs.assign(m.npair, DevPair{});
  for (int p = 0; p < m.npair; p++) {
    DevPair& P = pairs[p];
    P.g1 = m.pair_g1[p]; P.g2 = m.pair_g2[p]; P.dim = m.pair_condim[p];
    const bool box1 = m.geom_type[P.g1] == SO100_GEOM_BOX, box2 = m.geom_type[P.g2] == SO100_GEOM_BOX;
    if (geoms[P.g1].boxlike && geoms[P.g2].boxlike) P.mode = (box1 && box2) ? MODE_BOX_MULTI : MODE_BOX_SINGLE;
    else P.mode = MODE_HULL;
    P.f0 = (float)m.pair_friction[p][0]; P.f1 = (float)m.pair_friction[p][1];
    double si[5];
    for (int k = 0; k < 5; k++) si[k] = m.pair_solimp[p][k];
    si[0] = clampd(si[0], 1e-4, 0.9999); si[1] = clampd(si[1], 1e-4, 0.9999); si[2] = std::max(0.0, si[2]);
    si[3] = clampd(si[3], 1e-4, 0.9999); si[4] = std::max(1.0, si[4]);
    for (int k = 0; k < 5; k++) P.solimp[k] = (float)si[k];
    double tcp = m.pair_solref[p][0], dr = m.pair_solref[p][1];
    if (tcp <= 0) return fail(SO100_ERR_MODEL, "direct solref (negative) is not supported");
    tcp = std::max(tcp, 2 * m.timestep);
    P.K = (float)(1.0 / std::max(1e-15, si[1] * si[1] * tcp * tcp * dr * dr));
    P.B = (float)(2.0 / std::max(1e-15, si[1] * tcp));
    int b1 = m.geom_body[P.g1], b2 = m.geom_body[P.g2];
    P.dtran = (float)(m.body_invweight0[b1][0] + m.body_invweight0[b2][0]);
    P.drot = (float)(m.body_invweight0[b1][1] + m.body_invweight0[b2][1]);
    P.l1 = (short)geoms[P.g1].link; P.l2 = (short)geoms[P.g2].link;
    P.omd0 = (float)(1.0 - si[0]);
    P.dd = (float)(si[1] - si[0]);
  }
  // sites
  auto site_in_link = [&](int s, int& link, D3& p) {
    int b = m.site_body[s];
    D3 sp = {m.site_pos[s][0], m.site_pos[s][1], m.site_pos[s][2]};
    int root = b;
    while (root > 0 && m.body_jtype[root] < 0) root = m.body_parent[root];
    D3 bp; DQ bq;
    if (!rel_frame(m, b, root, bp, bq)) return false;
    double R[9];
    q2mat_h(bq, R);
    D3 r = rot_h(R, sp);
    p = {bp.x + r.x, bp.y + r.y, bp.z + r.z};
    link = root == 0 ? -1 : link_of_body[root];
    return true;
  };
  int lk; D3 sp;
  if (!site_in_link(m.site_ee, lk, sp) || lk != 4) return fail(SO100_ERR_MODEL, "ee_site must ride on link 4 (Fixed_Jaw)");
  dm.ee_off[0] = (float)sp.x; dm.ee_off[1] = (float)sp.y; dm.ee_off[2] = (float)sp.z;
  if (!site_in_link(m.site_cube, lk, sp) || lk != NL) return fail(SO100_ERR_MODEL, "cube_site must ride on the free body");
  dm.cube_site_off[0] = (float)sp.x; dm.cube_site_off[1] = (float)sp.y; dm.cube_site_off[2] = (float)sp.z;
  if (!site_in_link(m.site_bin, lk, sp) || lk != -1) return fail(SO100_ERR_MODEL, "bin_center must be static");
  const double bc[3] = {sp.x, sp.y, sp.z};
  for (int k = 0; k < 3; k++) dm.bin_center[k] = (float)bc[k];
  // single_arm.py:64-75 in float64
  dm.bin_min[0] = bc[0] - m.bin_hw; dm.bin_min[1] = bc[1] - m.bin_hw; dm.bin_min[2] = bc[2] + 0.0;
  dm.bin_max[0] = bc[0] + m.bin_hw; dm.bin_max[1] = bc[1] + m.bin_hw; dm.bin_max[2] = bc[2] + m.bin_h;
  for (int k = 0; k < 3; k++) {
    dm.box_lo[k] = (float)m.box_lo[k];
    dm.box_range[k] = (float)m.box_hi[k] - (float)m.box_lo[k];
    dm.bin_goal_lo[k] = (float)m.bin_goal_lo[k]; dm.bin_goal_hi[k] = (float)m.bin_goal_hi[k];
  }
  dm.cube_half = (float)m.cube_half; dm.goal_threshold = (float)m.goal_threshold;
  dm.lift_xy = (float)m.lift_goal_xy; dm.lift_zlo = (float)m.lift_goal_zlo; dm.lift_zhi = (float)m.lift_goal_zhi;
  dm.cg_cube = m.cg_cube; dm.cg_table = m.cg_table; dm.pad_mask = m.pad_mask;
  return SO100_OK;
}

// ------------------------------------------------------------------ launches
template <class ST> static size_t smem_of(unsigned lpe, int tpb = BLOCK) { return (size_t)(tpb / lpe) * sizeof(ST); }
static int grid_of(int n, unsigned lpe, int tpb = BLOCK) { const int epb = tpb / (int)lpe; return (n + epb - 1) / epb; }

static int configure_kernels(int device) {
  bool& done = g_configured[device];     // cudaFuncSetAttribute applies to the current device only
  if (done) return SO100_OK;
  CUDA_OK(cudaFuncSetAttribute(phase_kin_dyn<LPE_K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<KinS>(LPE_K1)));
  CUDA_OK(cudaFuncSetAttribute(phase_collide_box<LPE_K2A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<BoxS>(LPE_K2A)));
  CUDA_OK(cudaFuncSetAttribute(phase_kin_box<LPE_K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((BLOCK / LPE_K1) * KINBOX_SMEM)));
  CUDA_OK(cudaFuncSetAttribute(phase_collide_hull<LPE_K2B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<HullS>(LPE_K2B)));
  CUDA_OK(cudaFuncSetAttribute(phase_solve_light<LPE_K3L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<SolS<NCL>>(LPE_K3L, TPB_K3L)));
  CUDA_OK(cudaFuncSetAttribute(phase_slow_lane<LPE_K1, LPE_K2A, LPE_K3L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SLOW_SMEM));
  CUDA_OK(cudaFuncSetAttribute(phase_hull_solve<LPE_K3L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * HULL_SOLVE_SMEM)));
  CUDA_OK(cudaFuncSetAttribute(phase_solve_light_queue<LPE_K3L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<SolS<NCL>>(LPE_K3L, TPB_K3L)));
  CUDA_OK(cudaFuncSetAttribute(phase_solve_heavy<LPE_K3H, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<SolS<NC>>(LPE_K3H)));
  CUDA_OK(cudaFuncSetAttribute(phase_solve_heavy<LPE_K3H, NCL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<SolS<NCL>>(LPE_K3H)));
  done = true;
  return SO100_OK;
}

// optional per-kernel CUDA-event timing (so100_phase_timing): an event pair around each launch
static void mark(so100_ctx* h, cudaStream_t st, int cls, bool begin) {
  if (!h->timing) return;
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  cudaEventRecord(ev, st);
  h->events.push_back({ev, begin ? cls : -1});
}

// Kernel launch with a scheduling priority (cudaLaunchAttributePriority; also recorded in captured graph nodes).  The block
// scheduler hands freed SM slots to the pending kernel of the highest priority: the queue kernels (a few blocks that gate a whole
// env group's next substep) must not wait behind the thousands of pending blocks of the other groups' bulk kernels.
template <class... KArgs, class... Args>
static void launch_p(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, int prio, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePriority;
  at[0].val.priority = prio;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

static int k2b_grid(const so100_ctx* h, int n) {
  return std::min(grid_of(n, LPE_K2B), std::max(h->k2b_blocks << h->grid_class, n / std::max(1, h->k2b_div >> h->grid_class)));
}
static int k3lb_grid(const so100_ctx* h, int n) {
  // light queue b: ~14 % of a group's envs under random actions, two per block
  return std::min(grid_of(n, LPE_K3L, TPB_K3L), std::max(h->sm_count, n / (8 >> h->grid_class)));
}

// K1: kinematics (+ dynamics), K2a: box collision stage.  Leaves frames, M, qfrc_smooth, the box contacts and the hull-pair queue.
static void launch_kin_box(so100_ctx* h, EnvGroup& G, cudaStream_t st, const float* action, int with_dyn, int reuse = 0) {
  const int n = G.n;
  const DevTables T = h->tables();
  const Queues Q = h->queues(G);
  float* state = h->state + (size_t)G.off * STATE_WORDS;
  float* work = h->work + (size_t)G.off * WORK_WORDS;
  if (h->fuse_k12 && LPE_K1 == LPE_K2A && !h->timing) {       // timing mode keeps K1 / K2a apart: one duration per phase
    mark(h, st, CLS_KIN, true);
    launch_p(phase_kin_box<LPE_K1>, grid_of(n, LPE_K1), BLOCK, (BLOCK / LPE_K1) * KINBOX_SMEM, st, h->prio_mid, state, work,
             action ? action + (size_t)G.off * 6 : nullptr, n, with_dyn, T, Q, G.stage, reuse);
    mark(h, st, CLS_KIN, false);
    return;
  }
  mark(h, st, CLS_KIN, true);
  launch_p(phase_kin_dyn<LPE_K1>, grid_of(n, LPE_K1), BLOCK, smem_of<KinS>(LPE_K1), st, h->prio_mid, state, work, action ? action + (size_t)G.off * 6 : nullptr, n, with_dyn, Q, G.stage);
  mark(h, st, CLS_KIN, false); mark(h, st, CLS_BOX, true);
  launch_p(phase_collide_box<LPE_K2A>, grid_of(n, LPE_K2A), BLOCK, smem_of<BoxS>(LPE_K2A), st, h->prio_mid, work, n, T, Q, reuse);
  mark(h, st, CLS_BOX, false);
}
// K2b: GJK/EPA over the hull-pair queue
static void launch_hull(so100_ctx* h, EnvGroup& G, cudaStream_t st) {
  float* work = h->work + (size_t)G.off * WORK_WORDS;
  mark(h, st, CLS_HULL, true);
  launch_p(phase_collide_hull<LPE_K2B>, k2b_grid(h, G.n), BLOCK, smem_of<HullS>(LPE_K2B), st, h->prio_high, work, h->tables(), h->queues(G));
  mark(h, st, CLS_HULL, false);
}
// the whole position stage on one stream (trailing mj_step1, so100_forward)
static void launch_position_stage(so100_ctx* h, EnvGroup& G, cudaStream_t st, const float* action, int with_dyn, int reuse = 0) {
  launch_kin_box(h, G, st, action, with_dyn, reuse);
  launch_hull(h, G, st);
}

// One substep after K1 / K2a have been enqueued on `st`: K2b and the solve classes (+ Euler unless O.forward).
//   st     K3l-a (regular light grid: envs without hull pairs)                       |
//   side   K3m-a (medium queue a: arm-cube contact known after K2a)                  |  all rejoin `st`
//   hull   K2b -> K3l-b (light queue b)                                              |
//   side2  after K2b: K3m-b (medium queue b), K3h (heavy queue)                      |
static void launch_solve_stage(so100_ctx* h, EnvGroup& G, cudaStream_t st, const SolveOut& O, bool hull_done = false) {
  const int n = G.n;
  const DevTables T = h->tables();
  const Queues Q = h->queues(G);
  if (!O.forward) G.parity ^= 1;     // this launch writes the permutation the next one reads
  float* state = h->state + (size_t)G.off * STATE_WORDS;
  const float* work = h->work + (size_t)G.off * WORK_WORDS;
  const int med_grid = std::min(grid_of(n, LPE_K3H), h->k3m_blocks << h->grid_class), heavy_grid = std::min(grid_of(n, LPE_K3H), h->k3h_blocks);
  auto medium = [&](cudaStream_t s_, int which) {
    launch_p(phase_solve_heavy<LPE_K3H, NCL>, med_grid, BLOCK, smem_of<SolS<NCL>>(LPE_K3H), s_, h->prio_high, state, work, T, Q, O, which);
  };
  auto heavy = [&](cudaStream_t s_) {
    launch_p(phase_solve_heavy<LPE_K3H, NC>, heavy_grid, BLOCK, smem_of<SolS<NC>>(LPE_K3H), s_, h->prio_high, state, work, T, Q, O, 0);
  };
  auto light_a = [&](cudaStream_t s_) {
    launch_p(phase_solve_light<LPE_K3L>, grid_of(n, LPE_K3L, TPB_K3L), TPB_K3L, smem_of<SolS<NCL>>(LPE_K3L, TPB_K3L), s_, h->prio_low, state, work, n, T, Q, O);
  };
  auto light_b = [&](cudaStream_t s_) {
    launch_p(phase_solve_light_queue<LPE_K3L>, k3lb_grid(h, n), TPB_K3L, smem_of<SolS<NCL>>(LPE_K3L, TPB_K3L), s_, h->prio_high, state, work, T, Q, O);
  };
  const bool split = h->dag != 0;
  if (h->slow_on && !hull_done) {
    // slow lane: the rare classes and the over-budget envs are suspended by the kernels below; the chain is K2b -> budgeted light grid
    launch_hull(h, G, st);
    mark(h, st, CLS_SOLVE, true);
    light_a(st);
    mark(h, st, CLS_SOLVE, false);
    return;
  }
  if (h->timing && h->dag == 3 && !hull_done) {
    mark(h, st, CLS_HULL, true);
    launch_p(phase_hull_solve<LPE_K3L>, k2b_grid(h, n), BLOCK, 4 * HULL_SOLVE_SMEM, st, h->prio_high, state, h->work + (size_t)G.off * WORK_WORDS, T, Q, O);
    mark(h, st, CLS_HULL, false);
    mark(h, st, CLS_HEAVY, true);
    medium(st, 0); heavy(st);
    mark(h, st, CLS_HEAVY, false);
    mark(h, st, CLS_SOLVE, true);
    light_a(st);
    mark(h, st, CLS_SOLVE, false);
    return;
  }
  if (h->timing || hull_done) {
    // timing mode (one stream, so that every event pair brackets its kernels alone) and so100_forward (K2b already ran)
    if (!hull_done) launch_hull(h, G, st);
    mark(h, st, CLS_HEAVY, true);
    medium(st, 0);
    if (split) medium(st, 1);
    heavy(st);
    mark(h, st, CLS_HEAVY, false);
    mark(h, st, CLS_SOLVE, true);
    light_a(st);
    if (split) light_b(st);
    mark(h, st, CLS_SOLVE, false);
    return;
  }
  if (h->dag == 0) {
    // K2b on the chain, then the three solve classes side by side
    launch_hull(h, G, st);
    cudaEventRecord(G.fork, st);
    cudaStreamWaitEvent(G.side, G.fork, 0);
    cudaStreamWaitEvent(G.side2, G.fork, 0);
    medium(G.side, 0);
    heavy(G.side2);
    cudaEventRecord(G.join, G.side);
    cudaEventRecord(G.join2, G.side2);
    light_a(st);
    cudaStreamWaitEvent(st, G.join, 0);
    cudaStreamWaitEvent(st, G.join2, 0);
    return;
  }
  if (h->dag == 3) {
    // K2b fused with the solves of the envs it completes, beside the light grid of the other envs and the medium queue a; the
    // (rare) heavy class after it
    cudaEventRecord(G.fork, st);
    cudaStreamWaitEvent(G.side, G.fork, 0);
    cudaStreamWaitEvent(G.hull, G.fork, 0);
    medium(G.side, 0);
    cudaEventRecord(G.join, G.side);
    launch_p(phase_hull_solve<LPE_K3L>, k2b_grid(h, n), BLOCK, 4 * HULL_SOLVE_SMEM, G.hull, h->prio_high, state, h->work + (size_t)G.off * WORK_WORDS,
             T, Q, O);
    heavy(G.hull);
    cudaEventRecord(G.join3, G.hull);
    light_a(st);
    cudaStreamWaitEvent(st, G.join, 0);
    cudaStreamWaitEvent(st, G.join3, 0);
    return;
  }
  if (h->dag == 2) {
    // only K2b beside the light grid; every queue kernel after K2b
    cudaEventRecord(G.fork, st);
    cudaStreamWaitEvent(G.hull, G.fork, 0);
    launch_hull(h, G, G.hull);
    cudaEventRecord(G.fork2, G.hull);
    light_b(G.hull);
    cudaEventRecord(G.join3, G.hull);
    cudaStreamWaitEvent(G.side2, G.fork2, 0);
    medium(G.side2, 0); medium(G.side2, 1); heavy(G.side2);
    cudaEventRecord(G.join2, G.side2);
    light_a(st);
    cudaStreamWaitEvent(st, G.join2, 0);
    cudaStreamWaitEvent(st, G.join3, 0);
    return;
  }
  cudaEventRecord(G.fork, st);
  cudaStreamWaitEvent(G.side, G.fork, 0);
  cudaStreamWaitEvent(G.hull, G.fork, 0);
  medium(G.side, 0);
  cudaEventRecord(G.join, G.side);
  launch_hull(h, G, G.hull);
  cudaEventRecord(G.fork2, G.hull);
  light_b(G.hull);
  cudaEventRecord(G.join3, G.hull);
  cudaStreamWaitEvent(G.side2, G.fork2, 0);
  medium(G.side2, 1);
  heavy(G.side2);
  cudaEventRecord(G.join2, G.side2);
  light_a(st);
  cudaStreamWaitEvent(st, G.join, 0);
  cudaStreamWaitEvent(st, G.join2, 0);
  cudaStreamWaitEvent(st, G.join3, 0);
}

// Slow-lane kernel of stage `stage` (after that stage's light kernel): beside the regular pipeline on its own stream, joined by
// join_slow_lanes before anything reads the whole group's state (the task kernel, the end of so100_substeps).
static void launch_slow_lane(so100_ctx* h, EnvGroup& G, cudaStream_t st, int stage, int nsub, int trailing) {
  if (!h->slow_on) return;
  float* state = h->state + (size_t)G.off * STATE_WORDS;
  float* work = h->work + (size_t)G.off * WORK_WORDS;
  const int grid = std::max(32, std::min(4 * h->sm_count, G.n / 16));
  const int k = stage % EnvGroup::NSLOW;
  cudaStream_t s_ = st;
  if (!h->timing) {
    cudaEventRecord(G.slow_fork[k], st);
    cudaStreamWaitEvent(G.slow[k], G.slow_fork[k], 0);
    s_ = G.slow[k];
  }
  mark(h, s_, CLS_HEAVY, true);
  launch_p(phase_slow_lane<LPE_K1, LPE_K2A, LPE_K3L>, grid, 32, SLOW_SMEM, s_, h->prio_high, state, work, h->tables(), h->queues(G), stage, nsub, trailing);
  mark(h, s_, CLS_HEAVY, false);
  if (!h->timing) cudaEventRecord(G.slow_done[k], G.slow[k]);
}
static void join_slow_lanes(so100_ctx* h, EnvGroup& G, cudaStream_t st, int nstages) {
  if (!h->slow_on || h->timing) return;
  for (int k = 0; k < std::min(nstages, (int)EnvGroup::NSLOW); k++) cudaStreamWaitEvent(st, G.slow_done[k], 0);
}

static int make_group(so100_ctx* h, EnvGroup& G, int off, int n, int index, bool own_stream) {
  G.off = off; G.n = n; G.ctl = h->qmem + Q_STRIDE * index; G.index = index;
  G.order = h->order + (own_stream ? 0 : 2 * (size_t)h->n) + off;
  {
    std::vector<int> ident(n);
    for (int i = 0; i < n; i++) ident[i] = i;
    CUDA_OK(cudaMemcpy(G.order, ident.data(), n * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(G.order + h->n, ident.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  }
  if (own_stream) CUDA_OK(cudaStreamCreateWithFlags(&G.st, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&G.side, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&G.side2, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&G.hull, cudaStreamNonBlocking));
  for (int k = 0; k < EnvGroup::NSLOW; k++) {
    CUDA_OK(cudaStreamCreateWithFlags(&G.slow[k], cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&G.slow_fork[k], cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&G.slow_done[k], cudaEventDisableTiming));
  }
  CUDA_OK(cudaEventCreateWithFlags(&G.join2, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.join3, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.fork2, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.join, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.done, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreate(&G.t_done));
  CUDA_OK(cudaEventCreateWithFlags(&G.staged, cudaEventDisableTiming));
  return SO100_OK;
}
static void free_group(EnvGroup& G) {
  if (G.st) cudaStreamDestroy(G.st);
  if (G.side) cudaStreamDestroy(G.side);
  if (G.side2) cudaStreamDestroy(G.side2);
  if (G.hull) cudaStreamDestroy(G.hull);
  for (int k = 0; k < EnvGroup::NSLOW; k++) {
    if (G.slow[k]) cudaStreamDestroy(G.slow[k]);
    if (G.slow_fork[k]) cudaEventDestroy(G.slow_fork[k]);
    if (G.slow_done[k]) cudaEventDestroy(G.slow_done[k]);
  }
  if (G.join2) cudaEventDestroy(G.join2);
  if (G.join3) cudaEventDestroy(G.join3);
  if (G.fork2) cudaEventDestroy(G.fork2);
  if (G.fork) cudaEventDestroy(G.fork);
  if (G.join) cudaEventDestroy(G.join);
  if (G.done) cudaEventDestroy(G.done);
  if (G.t_done) cudaEventDestroy(G.t_done);
  if (G.staged) cudaEventDestroy(G.staged);
}

// Runs `body(group, stream)` for every env group: on the groups' own streams, forked from and joined back into `st`,
// or (one group / timing mode) directly on `st`.
template <class F> static void for_each_group(so100_ctx* h, cudaStream_t st, bool allow_groups, F body) {
  // both copies of a group's stage words start a pipeline call at zero (afterwards every stage re-arms the next one's copy)
  auto arm = [](EnvGroup& G, cudaStream_t s_) { cudaMemsetAsync(G.ctl, 0, 2 * Q_PARITY * sizeof(int), s_); };
  if (!allow_groups || h->timing || h->groups.size() < 2) { arm(h->whole, st); body(h->whole, st); return; }
  cudaEventRecord(h->ev_start, st);
  if (h->group_times) cudaEventRecordWithFlags(h->t_start, st, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
  for (size_t gi = 0; gi < h->groups.size(); gi++) {
    EnvGroup& G = h->groups[gi];
    cudaStreamWaitEvent(G.st, h->ev_start, 0);
    if (gi > 0 && (h->stagger == 2 || (h->stagger == 1 && (gi & 1)))) cudaStreamWaitEvent(G.st, h->groups[gi - 1].staged, 0);
    arm(G, G.st);
    body(G, G.st);
    if (h->group_times) cudaEventRecordWithFlags(G.t_done, G.st, h->capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
    cudaEventRecord(G.done, G.st);
    cudaStreamWaitEvent(st, G.done, 0);
  }
}

// device allocations, streams, events and tables of a new handle (any failure leaves a partially built context behind)
static int create_device_side(so100_ctx* h, const DevModel& dm, const std::vector<DevGeom>& geoms, const std::vector<DevPair>& pairs,
                              const std::vector<float4>& verts) {
  const int num_envs = h->n, device = h->device;
  int rc = SO100_OK;
  CUDA_OK(cudaMemcpyToSymbol(c_m, &dm, sizeof(dm)));
  CUDA_OK(cudaMalloc(&h->state, (size_t)num_envs * STATE_WORDS * sizeof(float)));
  CUDA_OK(cudaMalloc(&h->geom, geoms.size() * sizeof(DevGeom)));
  CUDA_OK(cudaMalloc(&h->pair, pairs.size() * sizeof(DevPair)));
  CUDA_OK(cudaMalloc(&h->vert, std::max<size_t>(verts.size(), 1) * sizeof(float4)));
  CUDA_OK(cudaMalloc(&h->diag, SO100_NDIAG * sizeof(unsigned long long)));
  CUDA_OK(cudaMalloc(&h->work, (size_t)num_envs * WORK_WORDS * sizeof(float)));
  CUDA_OK(cudaMemset(h->work, 0, (size_t)num_envs * WORK_WORDS * sizeof(float)));
  CUDA_OK(cudaMalloc(&h->qmem, so100_ctx::qmem_words(num_envs) * sizeof(int)));
  CUDA_OK(cudaMemset(h->qmem, 0, so100_ctx::qmem_words(num_envs) * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->order, 4 * (size_t)num_envs * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->qstat, 2 * sizeof(int)));
  CUDA_OK(cudaMemset(h->qstat, 0, 2 * sizeof(int)));
  {
    int* p = nullptr;
    CUDA_OK(cudaHostAlloc(&p, 2 * sizeof(int), cudaHostAllocDefault));
    p[0] = p[1] = 0;
    h->qstat_host = p;
  }
  if (const char* e = getenv("SO100_ADAPTIVE_GRIDS")) h->adaptive_grids = atoi(e) != 0;
  CUDA_OK(cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreate(&h->t_start));
  if (const char* e = getenv("SO100_GROUP_TIMES")) h->group_times = atoi(e) != 0;
  if (const char* e = getenv("SO100_STAGGER")) h->stagger = atoi(e);
  if (const char* e = getenv("SO100_DAG")) h->dag = atoi(e);
  if (const char* e = getenv("SO100_SLOWLANE")) h->slowlane_enabled = atoi(e) != 0;
  // K1 + K2a as one kernel saves a kernel boundary per stage where the step is latency-bound (4096 envs +2 %, 16384 +1 %) and costs
  // K1 its occupancy (80 instead of 40 registers) where it is throughput-bound (131072 envs -5 %)
  h->fuse_k12 = h->n < 49152;
  if (const char* e = getenv("SO100_FUSE_K12")) h->fuse_k12 = atoi(e) != 0;
  h->pair_far = false;      // measured: no gain at 16384 envs (the dense queue kernel ends the stage), -2 % at 131072
  if (const char* e = getenv("SO100_PAIR_FAR")) h->pair_far = atoi(e) != 0;
  if (const char* e = getenv("SO100_BUDGET_NEWTON")) h->budget_newton = std::max(1, atoi(e));
  if (const char* e = getenv("SO100_BUDGET_GJK")) h->budget_gjk = std::max(1, atoi(e));
  if (const char* e = getenv("SO100_BUDGET_EPA")) h->budget_epa = std::max(1, atoi(e));
  {
    int least = 0, greatest = 0, mode = 1;
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&least, &greatest));     // numerically lower = higher priority
    if (const char* e = getenv("SO100_PRIO")) mode = atoi(e);
    if (mode == 1) { h->prio_low = least; h->prio_high = greatest; h->prio_mid = (least + greatest) / 2; }
    else if (mode == 2) { h->prio_low = least; h->prio_high = greatest; h->prio_mid = greatest; }
    else if (mode == 3) { h->prio_low = least; h->prio_high = greatest; h->prio_mid = least; }
  }
  CUDA_OK(cudaStreamCreateWithFlags(&h->cap, cudaStreamNonBlocking));
  CUDA_OK(cudaMalloc(&h->act_stage, (size_t)num_envs * 6 * sizeof(float)));
  if (const char* e = getenv("SO100_GRAPH")) h->use_graph = atoi(e) != 0;
  if (const char* e = getenv("SO100_REUSE")) h->reuse_enabled = atoi(e) != 0;
  if (const char* e = getenv("SO100_K3H_BLOCKS")) h->k3h_blocks = std::max(1, atoi(e));
  if (const char* e = getenv("SO100_K3M_BLOCKS")) h->k3m_blocks = std::max(1, atoi(e));
  if (const char* e = getenv("SO100_K2B_BLOCKS")) h->k2b_blocks = std::max(1, atoi(e));
  if (const char* e = getenv("SO100_K2B_DIV")) h->k2b_div = std::max(1, atoi(e));
  {
    // env groups: SO100_GROUPS overrides.  Round 1 (one env per warp in the solve kernel): one group per 1024 envs, at most 6 (16384 envs:
    // 5 groups 2.53 ms, 6 2.54, 7 2.54, 8 2.56, 10 2.62 per step).
    // round 2, after the two-envs-per-warp kernels (16384 envs: 1 group 2.58 ms, 2 2.46, 3 2.37, 4 2.38, 6 2.41, 8 2.52; 65536 envs: 2 groups
    // 5.95 ms, 3 5.98, 4 6.12, 6 6.44; 131072 envs: 2 10.9, 4 11.0, 6 11.6; 4096 envs: 3 or 4): three groups, two for large batches
    int ng = num_envs >= 49152 ? 2 : std::min(3, num_envs / 1024);
    if (const char* e = getenv("SO100_GROUPS")) ng = atoi(e);
    ng = std::max(1, std::min(ng, 32));
    rc = make_group(h, h->whole, 0, num_envs, 32, false);
    if (rc) return rc;
    if (ng > 1) {
      const int per = ((num_envs + ng - 1) / ng + 7) / 8 * 8;
      for (int g = 0, off = 0; g < ng && off < num_envs; g++, off += per) {
        h->groups.emplace_back();
        rc = make_group(h, h->groups.back(), off, std::min(per, num_envs - off), g, true);
        if (rc) return rc;
      }
    }
  }
  CUDA_OK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
  CUDA_OK(cudaMemcpy(h->geom, geoms.data(), geoms.size() * sizeof(DevGeom), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->pair, pairs.data(), pairs.size() * sizeof(DevPair), cudaMemcpyHostToDevice));
  if (!verts.empty()) CUDA_OK(cudaMemcpy(h->vert, verts.data(), verts.size() * sizeof(float4), cudaMemcpyHostToDevice));
  {
    std::vector<uchar4> bp(pairs.size());
    for (size_t p = 0; p < pairs.size(); p++) {
      // bit 7 of the mode byte: a contact of this pair couples an arm link with the cube (dense 12x12 Hessian)
      const int l1 = pairs[p].l1, l2 = pairs[p].l2;
      const bool couples = (l1 >= 0 && l1 < NL && l2 == NL) || (l2 >= 0 && l2 < NL && l1 == NL);
      bp[p] = make_uchar4((unsigned char)pairs[p].g1, (unsigned char)pairs[p].g2, (unsigned char)(pairs[p].mode | (couples ? PAIR_COUPLES : 0)),
                          (unsigned char)pairs[p].dim);
    }
    CUDA_OK(cudaMalloc(&h->bpair, bp.size() * sizeof(uchar4)));
    CUDA_OK(cudaMemcpy(h->bpair, bp.data(), bp.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
  }
  const int total = num_envs * STATE_WORDS;
  init_state_kernel<<<(total + 255) / 256, 256>>>(h->state, num_envs);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaDeviceSynchronize());
  return SO100_OK;
}

extern "C" {

const char* so100_last_error(void) { return g_err.c_str(); }

int so100_create(const void* model_blob, size_t nbytes, int num_envs, int device, int task, uint64_t seed,
                 int64_t env_offset, so100_handle* out) {
  if (!model_blob || !out || num_envs <= 0) return fail(SO100_ERR_ARG, "so100_create: bad argument");
  if (nbytes != sizeof(so100_model)) return fail(SO100_ERR_ARG, "so100_create: model blob has the wrong size");
  if (task < SO100_TASK_CUBE_TO_BIN || task > SO100_TASK_TOUCH_CUBE_SPARSE) return fail(SO100_ERR_ARG, "so100_create: unknown task");
  so100_model m;
  memcpy(&m, model_blob, sizeof(m));
  if (m.magic != SO100_MODEL_MAGIC || m.version != SO100_MODEL_VERSION) return fail(SO100_ERR_ARG, "so100_create: model magic/version mismatch");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev <= 0)
    return fail(SO100_ERR_CUDA, "so100_create: no CUDA device available (this library has no CPU path)");
  if (device < 0 || device >= ndev) return fail(SO100_ERR_ARG, "so100_create: bad device index");
  if (device >= MAX_DEVICES) return fail(SO100_ERR_ARG, "so100_create: device index beyond the library's table");
  DeviceGuard guard(device);
  DevModel dm;
  std::vector<DevGeom> geoms;
  std::vector<DevPair> pairs;
  std::vector<float4> verts;
  int rc = build_dev_model(m, dm, geoms, pairs, verts);
  if (rc) return rc;
  rc = configure_kernels(device);
  if (rc) return rc;
  // the uniform model constants live in __constant__ memory shared by every handle of this process on a device:
  // live handles on one device must agree on the model
  so100_ctx* h = new so100_ctx();
  h->n = num_envs; h->device = device; h->task = task; h->seed = seed; h->env_offset = env_offset;
  h->nsub = m.nsubstep;
  {
    std::lock_guard<std::mutex> lock(g_model_mutex);
    if (g_live_handles[device] > 0 && memcmp(&g_live_model[device], &dm, sizeof(dm)) != 0) {
      delete h;
      return fail(SO100_ERR_MODEL, "so100_create: a live handle of this process uses a different model on this device (one model per process and device)");
    }
    g_live_model[device] = dm;
    g_live_handles[device]++;
    h->registered = true;
  }
  // everything below may fail half-way: so100_destroy releases whatever exists by then (and the live-handle count)
  rc = create_device_side(h, dm, geoms, pairs, verts);
  if (rc) {
    const std::string msg = g_err;
    so100_destroy(h);
    g_err = msg;
    return rc;
  }
  *out = h;
  return SO100_OK;
}

int so100_destroy(so100_handle h) {
  if (!h) return SO100_OK;
  if (h->registered) {
    std::lock_guard<std::mutex> lock(g_model_mutex);
    if (g_live_handles[h->device] > 0) g_live_handles[h->device]--;
  }
  DeviceGuard guard(h->device);
  cudaFree(h->qstat);
  if (h->qstat_host) cudaFreeHost((void*)h->qstat_host);
  cudaFree(h->state); cudaFree(h->geom); cudaFree(h->pair); cudaFree(h->vert); cudaFree(h->bpair); cudaFree(h->diag); cudaFree(h->work); cudaFree(h->qmem); cudaFree(h->order);
  free_group(h->whole);
  for (EnvGroup& G : h->groups) free_group(G);
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  if (h->t_start) cudaEventDestroy(h->t_start);
  for (auto& e : h->events) cudaEventDestroy(e.first);
  for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
  cudaFree(h->s_obs); cudaFree(h->s_ag); cudaFree(h->s_dg); cudaFree(h->s_rew); cudaFree(h->s_fin);
  cudaFree(h->s_term); cudaFree(h->s_trunc); cudaFree(h->s_succ); cudaFree(h->ep_stats);
  cudaFree(h->r_planes); cudaFree(h->r_adr); cudaFree(h->r_num); cudaFree(h->r_rgb);
  if (h->cap) cudaStreamDestroy(h->cap);
  cudaFree(h->act_stage);
  cudaFree(h->h_action); cudaFree(h->h_obs); cudaFree(h->h_ag); cudaFree(h->h_dg); cudaFree(h->h_rew); cudaFree(h->h_fin);
  cudaFree(h->h_term); cudaFree(h->h_trunc); cudaFree(h->h_succ);
  delete h;
  return SO100_OK;
}

int so100_num_envs(so100_handle h) { return h ? h->n : SO100_ERR_ARG; }

int so100_launches_per_step(so100_handle h) {
  if (!h) return SO100_ERR_ARG;
  // nsub x (K1, K2a, K2b, K3l, K3m, K3h [+ K3l-b, K3m-b with the a / b work classes]) + trailing (K1, K2a, K2b) + K4
  // per stage: K1 + K2a (one kernel when fused), K2b, then K3l, K3m, K3h (with the slow lane: K3l and the lane kernel; schedules 1 and
  // 2: five solve kernels); trailing position stage; K4
  const bool slow = h->slowlane_enabled && h->grid_class == 0 && h->nsub >= 1 && h->nsub <= 11;
  const int front = (h->fuse_k12 && LPE_K1 == LPE_K2A) ? 2 : 3;
  const int per_group = h->nsub * (front + (slow ? 2 : ((h->dag == 1 || h->dag == 2) ? 5 : 3))) + front + 1;
  return per_group * (int)std::max<size_t>(h->groups.size(), 1);
}

int so100_reset(so100_handle h, const uint8_t* mask, const float* box_pose, float* obs, float* achieved, float* desired,
                void* stream) {
  if (!h) return fail(SO100_ERR_ARG, "so100_reset: null handle");
  DeviceGuard guard(h->device);
  h->work_fresh = false;
  cudaStream_t st = (cudaStream_t)stream;
  reset_kernel<LPE_K4><<<grid_of(h->n, LPE_K4), BLOCK, smem_of<TaskS>(LPE_K4), st>>>(h->state, mask, box_pose, obs, achieved, desired, h->n,
                                                                                      h->task, (uint32_t)h->seed, (uint32_t)(h->seed >> 32),
                                                                                      h->env_offset);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

// Page-locked host destinations of so100_step_host: every env group copies its slice of the results out on its own stream
// as soon as its task kernel is done, so the device -> host traffic of the early groups hides behind the late groups' compute
// (and, under graph replay, costs no launch calls).
struct HostOut {
  float *obs = nullptr, *achieved = nullptr, *desired = nullptr, *reward = nullptr, *final_obs = nullptr;
  uint8_t *terminated = nullptr, *truncated = nullptr, *success = nullptr;
};

// the launch sequence of one env step on `stream` (directly, or while `stream` is being captured into a graph)
static void enqueue_step(so100_ctx* h, const StepArgs& A, cudaStream_t stream, int reuse, const HostOut* ho = nullptr) {
  const float* action = A.action;
  // the slow lane belongs to the small grid class: when a policy holds thousands of cubes at once (large class) the rare
  // classes are not rare and the medium / heavy queue kernels serve them better
  h->slow_on = h->slowlane_enabled && h->grid_class == 0 && h->nsub >= 1 && h->nsub <= 11;
  for_each_group(h, stream, true, [&](EnvGroup& G, cudaStream_t st) {
    const SolveOut O{nullptr, nullptr, 0};
    for (int s = 0; s < h->nsub; s++) {
      G.stage = s;
      launch_kin_box(h, G, st, s == 0 ? action : nullptr, 1, s == 0 ? reuse : 0);
      if (s == 0) cudaEventRecord(G.staged, st);
      launch_solve_stage(h, G, st, O);
      launch_slow_lane(h, G, st, s, h->nsub, 1);
    }
    // trailing mj_step1 (dm_control's legacy step): positions + contacts of the new state, then the task layer
    G.stage = std::min(h->nsub, 11);
    launch_position_stage(h, G, st, h->nsub == 0 ? action : nullptr, 0);
    join_slow_lanes(h, G, st, h->nsub);
    StepArgs B = A;
    B.trace = h->queues(G).trace;
    const size_t o = (size_t)G.off;
    B.state += o * STATE_WORDS; B.n = G.n; B.env_offset += G.off;
    if (B.obs) B.obs += o * 15;
    if (B.achieved) B.achieved += o * 3;
    if (B.desired) B.desired += o * 3;
    if (B.reward) B.reward += o;
    if (B.final_obs) B.final_obs += o * 15;
    if (B.terminated) B.terminated += o;
    if (B.truncated) B.truncated += o;
    if (B.success) B.success += o;
    if (B.ep_return) B.ep_return += o;
    if (B.ep_length) B.ep_length += o;
    mark(h, st, CLS_TASK, true);
    launch_p(phase_task<LPE_K4>, grid_of(G.n, LPE_K4), BLOCK, smem_of<TaskS>(LPE_K4), st, h->prio_mid, B, h->work + o * WORK_WORDS, h->tables());
    mark(h, st, CLS_TASK, false);
    if (ho) {
      auto out = [&](void* dst, const void* src, size_t per_env) {
        if (dst && src) cudaMemcpyAsync((char*)dst + o * per_env, src, (size_t)G.n * per_env, cudaMemcpyDeviceToHost, st);
      };
      out(ho->obs, B.obs, 60); out(ho->final_obs, B.final_obs, 60); out(ho->achieved, B.achieved, 12); out(ho->desired, B.desired, 12);
      out(ho->reward, B.reward, 4); out(ho->terminated, B.terminated, 1); out(ho->truncated, B.truncated, 1); out(ho->success, B.success, 1);
    }
  });
  // queue statistics of this step for the next steps' grid class (all groups have joined `stream` here)
  cudaMemcpyAsync((void*)h->qstat_host, h->qstat, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream);
  cudaMemsetAsync(h->qstat, 0, 2 * sizeof(int), stream);
}

static int step_impl(so100_handle h, const float* action, int autoreset, float* obs, float* achieved, float* desired, float* reward,
                     uint8_t* terminated, uint8_t* truncated, uint8_t* success, float* final_obs, void* stream, const HostOut* ho) {
  StepArgs A;
  A.state = h->state; A.action = action; A.obs = obs; A.achieved = achieved; A.desired = desired; A.reward = reward;
  A.final_obs = final_obs; A.terminated = terminated; A.truncated = truncated; A.success = success;
  A.n = h->n; A.autoreset = autoreset; A.task = h->task;
  A.ep_return = h->ep_return; A.ep_length = h->ep_length;
  A.seed_lo = (uint32_t)h->seed; A.seed_hi = (uint32_t)(h->seed >> 32); A.env_offset = h->env_offset;
  cudaStream_t st = (cudaStream_t)stream;
  const int reuse = (h->reuse_enabled && h->work_fresh && h->nsub > 0) ? 1 : 0;
  h->work_fresh = true;
  // more than 1/16 of a group in the medium queue, or more than 1/2 hull pair per env: large grids (with hysteresis)
  if (h->adaptive_grids) {
    const int med = h->qstat_host[0], hull = h->qstat_host[1];
    if (h->grid_class == 0 && (med > 64 || hull > 512)) h->grid_class = 1;
    else if (h->grid_class == 1 && med < 32 && hull < 384) h->grid_class = 0;
  }
  if (!h->use_graph || h->timing) {
    enqueue_step(h, A, st, reuse, ho);
    CUDA_OK(cudaGetLastError());
    return SO100_OK;
  }
  CUDA_OK(cudaMemcpyAsync(h->act_stage, action, (size_t)h->n * 6 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  A.action = h->act_stage;
  constexpr int NKEY = so100_ctx::NKEY;
  const void* key[NKEY];
  auto make_key = [&]() {
    const void* k[NKEY] = {A.obs, A.achieved, A.desired, A.reward, A.terminated, A.truncated, A.success, A.final_obs,
                           reinterpret_cast<const void*>((size_t)(autoreset != 0)), nullptr};
    if (ho) {
      const void* hk[8] = {ho->obs, ho->achieved, ho->desired, ho->reward, ho->final_obs, ho->terminated, ho->truncated, ho->success};
      memcpy(k + 10, hk, sizeof(hk));
    }
    k[18] = A.ep_return; k[19] = A.ep_length;
    memcpy(key, k, sizeof(k));
  };
  // same pointer set, any variant (key[9] = reuse | grid class << 1)
  auto same_set = [&](const so100_ctx::StepGraph& g) {
    return memcmp(g.key, key, 9 * sizeof(void*)) == 0 && memcmp(g.key + 10, key + 10, (NKEY - 10) * sizeof(void*)) == 0;
  };
  auto count_sets = [&]() {
    int nset = 0;
    for (size_t i = 0; i < h->graphs.size(); i++) {
      bool first = true;
      for (size_t j = 0; j < i; j++)
        if (memcmp(h->graphs[j].key, h->graphs[i].key, 9 * sizeof(void*)) == 0 &&
            memcmp(h->graphs[j].key + 10, h->graphs[i].key + 10, (NKEY - 10) * sizeof(void*)) == 0) { first = false; break; }
      nset += first;
    }
    return nset;
  };
  // caller's pointers, kept for the copies out of the staging outputs
  float *c_obs = obs, *c_ag = achieved, *c_dg = desired, *c_rew = reward, *c_fin = final_obs;
  uint8_t *c_term = terminated, *c_trunc = truncated, *c_succ = success;
  auto redirect = [&]() -> int {
    const size_t n = (size_t)h->n;
    if (!h->s_obs) {
      CUDA_OK(cudaMalloc(&h->s_obs, n * 15 * sizeof(float)));
      CUDA_OK(cudaMalloc(&h->s_fin, n * 15 * sizeof(float)));
      CUDA_OK(cudaMalloc(&h->s_ag, n * 3 * sizeof(float)));
      CUDA_OK(cudaMalloc(&h->s_dg, n * 3 * sizeof(float)));
      CUDA_OK(cudaMalloc(&h->s_rew, n * sizeof(float)));
      CUDA_OK(cudaMalloc(&h->s_term, n));
      CUDA_OK(cudaMalloc(&h->s_trunc, n));
      CUDA_OK(cudaMalloc(&h->s_succ, n));
    }
    A.obs = c_obs ? h->s_obs : nullptr; A.achieved = c_ag ? h->s_ag : nullptr; A.desired = c_dg ? h->s_dg : nullptr; A.reward = c_rew ? h->s_rew : nullptr;
    A.final_obs = c_fin ? h->s_fin : nullptr; A.terminated = c_term ? h->s_term : nullptr; A.truncated = c_trunc ? h->s_trunc : nullptr;
    A.success = c_succ ? h->s_succ : nullptr;
    return SO100_OK;
  };
  bool staged = h->stage_outputs && !ho;
  if (staged) { const int rc = redirect(); if (rc) return rc; }
  make_key();
  bool known = false;
  for (auto& g : h->graphs) known = known || same_set(g);
  if (!known && !staged && (int)count_sets() >= so100_ctx::GRAPH_SETS) {
    if (!ho) {
      // one pointer set too many: this caller rotates its output buffers.  From now on the graph runs on stable staging
      // outputs (captured once) and the results are copied to wherever the call points.
      h->stage_outputs = staged = true;
      const int rc = redirect();
      if (rc) return rc;
      make_key();
    } else {
      // page-locked host destinations are part of the graph: evict the least recently used pointer set
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); i++) if (h->graphs[i].used < h->graphs[lru].used) lru = i;
      const so100_ctx::StepGraph victim = h->graphs[lru];
      for (size_t i = h->graphs.size(); i-- > 0;) {
        so100_ctx::StepGraph& g = h->graphs[i];
        if (memcmp(g.key, victim.key, 9 * sizeof(void*)) == 0 && memcmp(g.key + 10, victim.key + 10, (NKEY - 10) * sizeof(void*)) == 0) {
          cudaGraphExecDestroy(g.exec);
          h->graphs.erase(h->graphs.begin() + i);
        }
      }
    }
  }
  auto find = [&](int cls) -> so100_ctx::StepGraph* {
    key[9] = reinterpret_cast<const void*>((size_t)(reuse | (cls << 1)));
    for (auto& g : h->graphs)
      if (memcmp(g.key, key, sizeof(key)) == 0) return &g;
    return nullptr;
  };
  const int chosen = h->grid_class;
  so100_ctx::StepGraph* hit = find(chosen);
  if (!hit) {
    // first use of this set of pointers: capture the graphs of both grid classes now, so that a later class change is a
    // cache hit and not a capture in the middle of a rollout
    for (int cls = 0; cls < (h->adaptive_grids ? 2 : 1); cls++) {
      if (find(cls)) continue;
      h->grid_class = cls;
      cudaGraph_t graph = nullptr;
      cudaGraphExec_t ge = nullptr;
      CUDA_OK(cudaStreamBeginCapture(h->cap, cudaStreamCaptureModeThreadLocal));
      h->capturing = true;
      enqueue_step(h, A, h->cap, reuse, ho);
      h->capturing = false;
      const cudaError_t launch_err = cudaGetLastError();
      cudaError_t ce = cudaStreamEndCapture(h->cap, &graph);     // always closes the capture, also after a failed launch
      if (ce == cudaSuccess && launch_err != cudaSuccess) ce = launch_err;
      if (ce == cudaSuccess) ce = cudaGraphInstantiateWithFlags(&ge, graph, cudaGraphInstantiateFlagUseNodePriority);   // per-node priorities of launch_p
      if (graph) cudaGraphDestroy(graph);
      if (ce != cudaSuccess) {
        h->grid_class = chosen;
        return fail(SO100_ERR_CUDA, std::string("so100_step: graph capture failed: ") + cudaGetErrorString(ce));
      }
      so100_ctx::StepGraph g;
      find(cls);                       // leaves this class's key in `key`
      memcpy(g.key, key, sizeof(key));
      g.exec = ge;
      g.used = 0;
      h->graphs.push_back(g);
      h->graph_captures++;
    }
    h->grid_class = chosen;
    hit = find(chosen);
    if (!hit) return fail(SO100_ERR_CUDA, "so100_step: graph capture failed");
  }
  hit->used = ++h->graph_clock;
  CUDA_OK(cudaGraphLaunch(hit->exec, st));
  if (staged) {
    const size_t n = (size_t)h->n;
    auto out = [&](void* dst, const void* src, size_t bytes) {
      return (dst && dst != src) ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) : cudaSuccess;
    };
    CUDA_OK(out(c_obs, h->s_obs, n * 60)); CUDA_OK(out(c_fin, h->s_fin, n * 60)); CUDA_OK(out(c_ag, h->s_ag, n * 12));
    CUDA_OK(out(c_dg, h->s_dg, n * 12)); CUDA_OK(out(c_rew, h->s_rew, n * 4)); CUDA_OK(out(c_term, h->s_term, n));
    CUDA_OK(out(c_trunc, h->s_trunc, n)); CUDA_OK(out(c_succ, h->s_succ, n));
  }
  return SO100_OK;
}

int so100_step(so100_handle h, const float* action, int autoreset, float* obs, float* achieved, float* desired, float* reward,
               uint8_t* terminated, uint8_t* truncated, uint8_t* success, float* final_obs, void* stream) {
  if (!h || !action) return fail(SO100_ERR_ARG, "so100_step: null handle or action");
  DeviceGuard guard(h->device);
  return step_impl(h, action, autoreset, obs, achieved, desired, reward, terminated, truncated, success, final_obs, stream, nullptr);
}

int so100_step_host(so100_handle h, const float* action, int autoreset, float* obs, float* achieved, float* desired,
                    float* reward, uint8_t* terminated, uint8_t* truncated, uint8_t* success, float* final_obs, void* stream) {
  if (!h || !action) return fail(SO100_ERR_ARG, "so100_step_host: null handle or action");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = h->n;
  if (!h->h_action) {
    CUDA_OK(cudaMalloc(&h->h_action, n * 6 * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_obs, n * 15 * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_fin, n * 15 * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_ag, n * 3 * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_dg, n * 3 * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_rew, n * sizeof(float)));
    CUDA_OK(cudaMalloc(&h->h_term, n));
    CUDA_OK(cudaMalloc(&h->h_trunc, n));
    CUDA_OK(cudaMalloc(&h->h_succ, n));
  }
  CUDA_OK(cudaMemcpyAsync(h->h_action, action, n * 6 * sizeof(float), cudaMemcpyHostToDevice, st));
  // page-locked destinations: the copies out ride in the step's own pipeline (per env group, see HostOut); pageable ones are
  // copied after the step (an asynchronous copy to pageable memory is staged by the driver and cannot be captured)
  auto pinned = [](const void* p) {
    if (!p) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  bool in_pipeline = pinned(obs) && pinned(achieved) && pinned(desired) && pinned(reward) && pinned(terminated) &&
                     pinned(truncated) && pinned(success) && pinned(final_obs);
  if (in_pipeline) {
    // page-locked destinations are baked into the step graph: only the first few distinct sets ride in the pipeline, a
    // caller that rotates more of them is served by copies after the step (stable device staging buffers, no re-capture)
    const std::array<const void*, 8> set = {obs, achieved, desired, reward, terminated, truncated, success, final_obs};
    bool seen = false;
    for (const auto& k : h->host_sets) seen = seen || k == set;
    if (!seen) {
      if ((int)h->host_sets.size() < so100_ctx::GRAPH_SETS / 2) h->host_sets.push_back(set);
      else in_pipeline = false;
    }
  }
  HostOut ho;
  ho.obs = obs; ho.achieved = achieved; ho.desired = desired; ho.reward = reward; ho.final_obs = final_obs;
  ho.terminated = terminated; ho.truncated = truncated; ho.success = success;
  int rc = step_impl(h, h->h_action, autoreset, h->h_obs, achieved ? h->h_ag : nullptr, desired ? h->h_dg : nullptr,
                     reward ? h->h_rew : nullptr, terminated ? h->h_term : nullptr, truncated ? h->h_trunc : nullptr,
                     success ? h->h_succ : nullptr, final_obs ? h->h_fin : nullptr, stream, in_pipeline ? &ho : nullptr);
  if (rc) return rc;
  if (!in_pipeline) {
    if (obs) CUDA_OK(cudaMemcpyAsync(obs, h->h_obs, n * 15 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (final_obs) CUDA_OK(cudaMemcpyAsync(final_obs, h->h_fin, n * 15 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (achieved) CUDA_OK(cudaMemcpyAsync(achieved, h->h_ag, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (desired) CUDA_OK(cudaMemcpyAsync(desired, h->h_dg, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (reward) CUDA_OK(cudaMemcpyAsync(reward, h->h_rew, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (terminated) CUDA_OK(cudaMemcpyAsync(terminated, h->h_term, n, cudaMemcpyDeviceToHost, st));
    if (truncated) CUDA_OK(cudaMemcpyAsync(truncated, h->h_trunc, n, cudaMemcpyDeviceToHost, st));
    if (success) CUDA_OK(cudaMemcpyAsync(success, h->h_succ, n, cudaMemcpyDeviceToHost, st));
  }
  CUDA_OK(cudaStreamSynchronize(st));
  return SO100_OK;
}

int so100_compute_reward(const float* achieved, const float* desired, int64_t n, float threshold, float* reward, void* stream) {
  if (!achieved || !desired || !reward || n < 0) return fail(SO100_ERR_ARG, "so100_compute_reward: bad argument");
  if (n == 0) return SO100_OK;
  compute_reward_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(achieved, desired, n, threshold, reward);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

static int state_io(so100_handle h, int dir, float* qpos, float* qvel, float* ctrl, float* warm, float* goal, int32_t* sc,
                    int32_t* ts, uint32_t* ep, void* stream) {
  if (!h) return fail(SO100_ERR_ARG, "state io: null handle");
  DeviceGuard guard(h->device);
  if (dir == 1) h->work_fresh = false;
  const int total = h->n * STATE_WORDS;
  state_io_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->state, h->n, dir, qpos, qvel, ctrl, warm, goal, sc, ts, ep);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}
int so100_get_state(so100_handle h, float* qpos, float* qvel, float* ctrl, float* warm, void* stream) {
  return state_io(h, 0, qpos, qvel, ctrl, warm, nullptr, nullptr, nullptr, nullptr, stream);
}
int so100_set_state(so100_handle h, const float* qpos, const float* qvel, const float* ctrl, const float* warm, void* stream) {
  return state_io(h, 1, (float*)qpos, (float*)qvel, (float*)ctrl, (float*)warm, nullptr, nullptr, nullptr, nullptr, stream);
}
int so100_get_aux(so100_handle h, float* goal, int32_t* step_count, int32_t* total_steps, uint32_t* episode, void* stream) {
  return state_io(h, 0, nullptr, nullptr, nullptr, nullptr, goal, step_count, total_steps, episode, stream);
}
int so100_set_aux(so100_handle h, const float* goal, const int32_t* step_count, const int32_t* total_steps,
                  const uint32_t* episode, void* stream) {
  return state_io(h, 1, nullptr, nullptr, nullptr, nullptr, (float*)goal, (int32_t*)step_count, (int32_t*)total_steps,
                  (uint32_t*)episode, stream);
}

int so100_substeps(so100_handle h, int nsub, void* stream) {
  if (!h || nsub < 0) return fail(SO100_ERR_ARG, "so100_substeps: bad argument");
  DeviceGuard guard(h->device);
  h->work_fresh = false;
  h->slow_on = h->slowlane_enabled && h->grid_class == 0;
  // at most 12 stages per pass: the slow lane keeps one queue cursor per stage of a pass
  for (int base = 0; base < nsub; base += 12) {
    const int chunk = std::min(12, nsub - base);
    for_each_group(h, (cudaStream_t)stream, true, [&](EnvGroup& G, cudaStream_t st) {
      const SolveOut O{nullptr, nullptr, 0};
      for (int s = 0; s < chunk; s++) {
        G.stage = s;
        launch_kin_box(h, G, st, nullptr, 1);
        if (s == 0) cudaEventRecord(G.staged, st);
        launch_solve_stage(h, G, st, O);
        launch_slow_lane(h, G, st, s, chunk, 0);
      }
      join_slow_lanes(h, G, st, chunk);
    });
  }
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

int so100_forward(so100_handle h, float* qacc, int32_t* ncon, int32_t* con_geom, float* con_data, float* sites, void* stream) {
  if (!h) return fail(SO100_ERR_ARG, "so100_forward: null handle");
  DeviceGuard guard(h->device);
  h->work_fresh = false;
  cudaStream_t st = (cudaStream_t)stream;
  // the whole batch as one group on the caller's stream, every kernel in sequence: K1, K2a, K2b, the contact export, then the
  // solve classes in forward mode (nothing is integrated)
  h->slow_on = false;        // mj_forward solves every env in the regular kernels (no budgets, medium / heavy queue kernels)
  h->whole.stage = 0;
  CUDA_OK(cudaMemsetAsync(h->whole.ctl, 0, 2 * Q_PARITY * sizeof(int), st));
  launch_kin_box(h, h->whole, st, nullptr, 1);
  launch_hull(h, h->whole, st);
  const int threads = h->n * 32;
  export_forward_kernel<<<(threads + 255) / 256, 256, 0, st>>>(h->work, h->n, ncon, con_geom, con_data, sites, h->tables());
  launch_solve_stage(h, h->whole, st, SolveOut{qacc, con_data, 1}, true);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

#ifdef SO100_TRACE
// development build: copy out (host [TRACE_RECORDS][2] uint64 ns) and re-arm the kernel trace
int so100_trace_read(unsigned long long* out) {
  CUDA_OK(cudaDeviceSynchronize());
  if (out) CUDA_OK(cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 2 * TRACE_RECORDS));
  std::vector<unsigned long long> init(2 * TRACE_RECORDS);
  for (int i = 0; i < TRACE_RECORDS; i++) { init[2 * i] = ~0ull; init[2 * i + 1] = 0ull; }
  CUDA_OK(cudaMemcpyToSymbol(g_trace, init.data(), sizeof(unsigned long long) * 2 * TRACE_RECORDS));
  return TRACE_RECORDS;
}
#endif

#ifdef SO100_HULL_CLOCK
// development build: copy out and clear the GJK/EPA item statistics; returns the number of items recorded
int so100_hull_stats(int32_t* out /* host [65536][4] */) {
  int n = 0;
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(&n, g_hull_stat_n, sizeof(int)));
  CUDA_OK(cudaMemcpyFromSymbol(out, g_hull_stat, sizeof(int) * 4 * 65536));
  const int zero = 0;
  CUDA_OK(cudaMemcpyToSymbol(g_hull_stat_n, &zero, sizeof(int)));
  return n;
}
int so100_hull_phases(int32_t* out /* host [65536][6] */) {
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(out, g_hull_phase, sizeof(int) * 6 * 65536));
  return 0;
}
#endif

#ifdef SO100_SOLVE_TRACE
// development build: per-iteration solver trace of envs 0..63 from the last so100_forward
int so100_solve_trace(float* out /* host [64][104][8] */) {
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpyFromSymbol(out, g_solve_trace, sizeof(float) * 64 * 104 * 8));
  return 0;
}
#endif

int so100_measure_fp32_peak(int device, float* tflops) {
  if (!tflops) return fail(SO100_ERR_ARG, "so100_measure_fp32_peak: null output");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(SO100_ERR_CUDA, "so100_measure_fp32_peak: no such CUDA device");
  DeviceGuard guard(device);
  int sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 1 << 15;
  float* sink = nullptr;
  CUDA_OK(cudaMalloc(&sink, (size_t)blocks * threads * sizeof(float)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  float best = 0.0f;
  for (int rep = 0; rep < 4; rep++) {     // first repetition warms the clocks up
    cudaEventRecord(e0);
    fma_peak_kernel<<<blocks, threads>>>(sink, iters, 1.0001f);
    cudaEventRecord(e1);
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    if (rep > 0) best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  *tflops = best;
  return SO100_OK;
}

int so100_group_times(so100_handle h, float* ms, int32_t* ngroups, void* stream) {
  if (!h || !ms || !ngroups) return fail(SO100_ERR_ARG, "so100_group_times: bad argument");
  DeviceGuard guard(h->device);
  *ngroups = 0;
  if (!h->group_times || h->groups.size() < 2) return SO100_OK;
  CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  for (EnvGroup& G : h->groups) {
    float e = 0;
    if (cudaEventElapsedTime(&e, h->t_start, G.t_done) != cudaSuccess) { cudaGetLastError(); e = -1.0f; }
    ms[(*ngroups)++] = e;
  }
  return SO100_OK;
}

int so100_phase_timing(so100_handle h, int enable, float* ms6, int32_t* launches6, void* stream) {
  if (!h) return fail(SO100_ERR_ARG, "so100_phase_timing: null handle");
  DeviceGuard guard(h->device);
  if (ms6 || launches6) {
    CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
    float ms[CLS_N] = {0};
    int cnt[CLS_N] = {0};
    for (size_t i = 0; i + 1 < h->events.size(); i += 2) {
      const int cls = h->events[i].second;
      float e = 0;
      if (cls >= 0 && cls < CLS_N && cudaEventElapsedTime(&e, h->events[i].first, h->events[i + 1].first) == cudaSuccess) { ms[cls] += e; cnt[cls]++; }
    }
    for (int k = 0; k < CLS_N; k++) { if (ms6) ms6[k] = ms[k]; if (launches6) launches6[k] = cnt[k]; }
  }
  for (auto& e : h->events) cudaEventDestroy(e.first);
  h->events.clear();
  h->timing = enable != 0;
  return SO100_OK;
}

int so100_debug_read(so100_handle h, int what, float* out, int64_t* words_per_env, void* stream) {
  if (!h || (what != 0 && what != 1)) return fail(SO100_ERR_ARG, "so100_debug_read: bad argument");
  DeviceGuard guard(h->device);
  const size_t words = what == 0 ? STATE_WORDS : WORK_WORDS;
  if (words_per_env) *words_per_env = (int64_t)words;
  if (out)
    CUDA_OK(cudaMemcpyAsync(out, what == 0 ? h->state : h->work, (size_t)h->n * words * sizeof(float), cudaMemcpyDeviceToDevice,
                            (cudaStream_t)stream));
  return SO100_OK;
}

int so100_set_episode_outputs(so100_handle h, float* ep_return, int32_t* ep_length) {
  if (!h) return fail(SO100_ERR_ARG, "so100_set_episode_outputs: null handle");
  h->ep_return = ep_return; h->ep_length = ep_length;     // part of the step graph's key: a change captures a new graph
  return SO100_OK;
}

int so100_episode_stats(so100_handle h, double* out4, void* stream) {
  if (!h || !out4) return fail(SO100_ERR_ARG, "so100_episode_stats: bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (!h->ep_stats) CUDA_OK(cudaMalloc(&h->ep_stats, 4 * sizeof(double)));
  CUDA_OK(cudaMemsetAsync(h->ep_stats, 0, 4 * sizeof(double), st));
  episode_reduce_kernel<<<148, 256, 0, st>>>(h->state, h->n, h->ep_stats);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(out4, h->ep_stats, 4 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  return SO100_OK;
}

// ---- HER ring (so100_her.cuh)
static int her_ring(const so100_her_ring* r, HerRing& R) {
  if (!r || r->capacity <= 0 || r->num_envs <= 0) return fail(SO100_ERR_ARG, "so100_her: bad ring geometry");
  const void* ptrs[] = {r->obs, r->next_obs, r->achieved, r->next_achieved, r->desired, r->action, r->reward, r->done,
                        r->ep_start, r->ep_length, r->cur_start, r->cur_length};
  for (const void* p : ptrs) if (!p) return fail(SO100_ERR_ARG, "so100_her: null ring array");
  R = HerRing{r->capacity, r->num_envs, r->obs, r->next_obs, r->achieved, r->next_achieved, r->desired, r->action, r->reward, r->done,
              r->ep_start, r->ep_length, r->cur_start, r->cur_length};
  return SO100_OK;
}

int so100_her_begin(const so100_her_ring* ring, int32_t pos, const float* obs, const float* achieved, const float* desired,
                    const float* action, void* stream) {
  HerRing R;
  if (int rc = her_ring(ring, R)) return rc;
  if (pos < 0 || pos >= R.capacity || !obs || !achieved || !desired || !action) return fail(SO100_ERR_ARG, "so100_her_begin: bad argument");
  her_begin_kernel<<<(R.num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(R, pos);
  her_begin_copy_kernel<<<(unsigned)(((long long)R.num_envs * 27 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, pos, obs, achieved, desired, action);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

int so100_her_commit(const so100_her_ring* ring, int32_t pos, const float* obs, const float* achieved, const float* final_obs,
                     const float* reward, const uint8_t* terminated, const uint8_t* truncated, void* stream) {
  HerRing R;
  if (int rc = her_ring(ring, R)) return rc;
  if (pos < 0 || pos >= R.capacity || !obs || !achieved || !final_obs || !reward || !terminated || !truncated)
    return fail(SO100_ERR_ARG, "so100_her_commit: bad argument");
  her_commit_copy_kernel<<<(unsigned)(((long long)R.num_envs * 18 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, pos, obs, achieved, final_obs,
                                                                                                                  terminated, truncated);
  her_commit_kernel<<<(R.num_envs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(R, pos, reward, terminated, truncated);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

int so100_her_sample(const so100_her_ring* ring, int64_t batch, int32_t n_sampled_goal, float threshold, uint64_t seed, uint32_t call,
                     float* obs, float* action, float* next_obs, float* achieved, float* next_achieved, float* desired, float* reward,
                     uint8_t* done, int32_t* index, void* stream) {
  HerRing R;
  if (int rc = her_ring(ring, R)) return rc;
  if (batch < 0 || n_sampled_goal < 0 || !obs || !action || !next_obs || !achieved || !next_achieved || !desired || !reward || !done || !index)
    return fail(SO100_ERR_ARG, "so100_her_sample: bad argument");
  if (batch == 0) return SO100_OK;
  her_pick_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, (cudaStream_t)stream>>>(R, batch, n_sampled_goal, threshold, (uint32_t)seed,
                                                                                     (uint32_t)(seed >> 32), call, reward, done, index);
  her_gather_kernel<<<(unsigned)((batch * 45 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(R, batch, index, obs, action, next_obs, achieved,
                                                                                            next_achieved, desired);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

int so100_render_config(so100_handle h, const float* planes, int32_t nplanes, const int32_t* geom_plane_adr, const int32_t* geom_plane_num,
                        const float* geom_rgb, const float* camera13, const float* lights, int32_t nlights, float ambient, float head_diffuse,
                        int32_t width, int32_t height) {
  if (!h || !geom_plane_adr || !geom_plane_num || !geom_rgb || !camera13 || nplanes < 0 || (nplanes > 0 && !planes) || width <= 0 || height <= 0 ||
      nlights < 0 || nlights > 4 || (nlights > 0 && !lights))
    return fail(SO100_ERR_ARG, "so100_render_config: bad argument");
  for (int g = 0; g < NGEOM; g++)
    if (geom_plane_num[g] > 0 && (geom_plane_adr[g] < 0 || geom_plane_adr[g] + geom_plane_num[g] > nplanes))
      return fail(SO100_ERR_ARG, "so100_render_config: plane range outside the table");
  DeviceGuard guard(h->device);
  cudaFree(h->r_planes); cudaFree(h->r_adr); cudaFree(h->r_num); cudaFree(h->r_rgb);
  h->r_planes = nullptr; h->r_adr = nullptr; h->r_num = nullptr; h->r_rgb = nullptr; h->render_ready = false;
  CUDA_OK(cudaMalloc(&h->r_planes, std::max(nplanes, 1) * sizeof(float4)));
  CUDA_OK(cudaMalloc(&h->r_adr, NGEOM * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->r_num, NGEOM * sizeof(int)));
  CUDA_OK(cudaMalloc(&h->r_rgb, NGEOM * 3 * sizeof(float)));
  if (nplanes > 0) CUDA_OK(cudaMemcpy(h->r_planes, planes, (size_t)nplanes * sizeof(float4), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->r_adr, geom_plane_adr, NGEOM * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->r_num, geom_plane_num, NGEOM * sizeof(int), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(h->r_rgb, geom_rgb, NGEOM * 3 * sizeof(float), cudaMemcpyHostToDevice));
  RenderCfg& C = h->rcfg;
  memset(&C, 0, sizeof(C));
  C.width = width; C.height = height;
  for (int k = 0; k < 3; k++) { C.cam_pos[k] = camera13[k]; C.cam_x[k] = camera13[3 + k]; C.cam_y[k] = camera13[6 + k]; C.cam_z[k] = camera13[9 + k]; }
  C.tan_half_fovy = tanf(0.5f * camera13[12] * 3.14159265358979f / 180.0f);
  C.ambient = ambient; C.head_diffuse = head_diffuse; C.nlight = nlights;
  for (int l = 0; l < nlights; l++) {
    const float* L = lights + 4 * l;
    const float len = std::sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    for (int k = 0; k < 3; k++) C.light_dir[l][k] = len > 0 ? L[k] / len : 0.0f;
    C.light_diffuse[l] = L[3];
  }
  h->render_ready = true;
  return SO100_OK;
}

int so100_render(so100_handle h, uint8_t* pixels, void* stream) {
  if (!h || !pixels) return fail(SO100_ERR_ARG, "so100_render: bad argument");
  if (!h->render_ready) return fail(SO100_ERR_ARG, "so100_render: call so100_render_config first");
  DeviceGuard guard(h->device);
  const RenderTables R{h->r_planes, h->r_adr, h->r_num, h->r_rgb};
  render_kernel<<<h->n, 128, 0, (cudaStream_t)stream>>>(h->state, h->n, h->rcfg, h->tables(), R, pixels);
  CUDA_OK(cudaGetLastError());
  return SO100_OK;
}

int so100_graph_stats(so100_handle h, int32_t* captures, int32_t* cached, int32_t* staged) {
  if (!h) return fail(SO100_ERR_ARG, "so100_graph_stats: null handle");
  if (captures) *captures = h->graph_captures;
  if (cached) *cached = (int32_t)h->graphs.size();
  if (staged) *staged = h->stage_outputs ? 1 : 0;
  return SO100_OK;
}

int so100_diagnostics(so100_handle h, int64_t* out8, void* stream) {
  if (!h || !out8) return fail(SO100_ERR_ARG, "so100_diagnostics: bad argument");
  DeviceGuard guard(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaMemsetAsync(h->diag, 0, SO100_NDIAG * sizeof(unsigned long long), st));
  diag_reduce_kernel<<<148, 256, 0, st>>>(h->state, h->n, h->diag);
  CUDA_OK(cudaGetLastError());
  unsigned long long tmp[SO100_NDIAG];
  CUDA_OK(cudaMemcpyAsync(tmp, h->diag, sizeof(tmp), cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  for (int k = 0; k < SO100_NDIAG; k++) out8[k] = (int64_t)tmp[k];
  return SO100_OK;
}

}  // extern "C"
