// Per-phase shared-memory layouts, the L2-resident workspace record that links the phase kernels,
// and the work-class queues.
//
// Every phase kernel owns one cooperative tile of LPE lanes per env and a scratch struct that holds
// only what that phase touches (1.7 KB for kinematics/dynamics, 1.2 KB for the box collision stage,
// 4 KB for a hull (GJK/EPA) env, 5.7 KB for a solve with <= 8 contacts, 14.5 KB for the rare solve with
// up to 24), so that shared memory never limits the number of resident warps.  The device functions
// are templates over the scratch type and only name the members they use.
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include "so100_dev.cuh"

namespace so100 {
namespace cg = cooperative_groups;

__constant__ DevModel c_m;

struct DevTables {
  const DevGeom* geom;
  const DevPair* pair;
  const float4* vert;
  const uchar4* bpair;   // broad-phase view of the pair table: (g1, g2, mode | PAIR_COUPLES, condim)
};

template <unsigned LPE> using Tile = cg::thread_block_tile<LPE>;

constexpr int PAIR_COUPLES = 0x80; // DevTables::bpair mode byte, bit 7: the pair joins an arm link and the cube
constexpr int HDR_COUPLED = 1 << 30;   // workspace header word 2: the env has such a contact this substep
constexpr int HDR_OVERFLOW = 1 << 29;  // workspace header word 2: more penetrating box pairs / hull pairs / contacts than the lists hold
constexpr int HDR_STALE = -1;          // workspace header word 3 (hull pairs pending, 0 after a complete collision stage): the
                                       // env was reset after its last collision stage
constexpr int NCL = 8;             // contact capacity of the light solve kernel (99.7 % of all solves)
constexpr int NHP = 16;            // hull pairs per env that may reach GJK/EPA after the oriented-box cull

// ---------------------------------------------------------------- workspace record (float words)
// 16-byte aligned sections so that tiles move them with 128-bit loads.
constexpr int W_FRAMES = 0;        // lpos[7][3] lmat[7][9] axis[6][3] (+2 pad)
constexpr int W_FRAMES_N = 104;
constexpr int W_DYN = 104;         // Mfull[6][6] qfs[12] qas[12]
constexpr int W_DYN_N = 60;
constexpr int W_HDR = 164;         // ncon, nhullpairs, stats (nbox | npen << 8 | nhull << 16), hull pairs still pending
constexpr int W_HULLP = 168;       // NHP pair ids, one byte each
constexpr int W_CON = 172;         // contact c: pos[3] nrm[3] dist pair
constexpr int CON_WORDS = 8;
constexpr int W_HSTAGE = W_CON + CON_WORDS * NC;     // GJK/EPA results by hull-pair slot (pair = -1: no contact), merged in pair order
constexpr int WORK_WORDS = W_HSTAGE + CON_WORDS * NHP + 4;   // 496 words = 1984 B = 62 sectors (the staging area is touched by hull envs only)
static_assert(WORK_WORDS % 8 == 0 && W_HDR % 4 == 0 && W_CON % 4 == 0, "workspace records stay 32-byte aligned, sections 16-byte aligned");
static_assert(NHP <= 16, "hull pair list is 4 words");

// queue control words (device ints)
enum { Q_HULL_COUNT = 0, Q_HULL_NEXT = 1, Q_HEAVY_COUNT = 2, Q_HEAVY_NEXT = 3, Q_SLOW = 4, Q_FAST = 5, Q_MEDA_COUNT = 6, Q_MEDA_NEXT = 7,
       Q_MEDB_COUNT = 8, Q_MEDB_NEXT = 9, Q_LB_COUNT = 10, Q_LB_NEXT = 11, Q_WORDS = 12,
       // The stage words above exist twice per group (Q_PARITY words apart): stage s works on copy s & 1 and its first kernel re-arms
       // the other copy for stage s + 1 (nobody touches that copy during stage s), so the re-arming needs no kernel boundary of its
       // own and K1 / K2a can be one kernel.  A pipeline call starts with a memset of both copies.
       Q_PARITY = 16,
       Q_LANE_BASE = 32,         // slow-lane words, outside the two stage copies (Queues::lanectl)
       Q_LANE_COUNT = 0,         // slow-lane queue length: grows over the stages of a step, re-armed by the first stage only
       Q_LANE_CURSOR = 4,        // [12] per stage: the next slow-lane queue entry of that stage's slow-lane kernel
       Q_STRIDE = 64 };

// Work classes of a substep.  The collision stage ends in two steps: the box stage (K2a) completes every env without a hull
// pair, the GJK/EPA queue kernel (K2b) the other ~14 %.  Everything that only needs K2a starts right after it and runs BESIDE
// K2b (the "a" classes: the regular grid of the light solve kernel, the medium queue a); the envs K2b completes are solved
// after it (the "b" classes: light queue b, medium queue b) together with the rare heavy class.  K2b's latency (its slowest
// EPA item, 35-55 us) is thereby off the critical path of a substep whenever the a-side is the longer one.
#ifdef SO100_TRACE
// development build: first-warp-start / last-warp-end (%globaltimer, ns) of every kernel launch of a step, keyed by
// (env group, stage, kernel kind); read with so100_trace_read (tools/gpu_trace.py draws the per-group timeline)
constexpr int TRACE_RECORDS = 8 * 12 * 10;
__device__ unsigned long long g_trace[TRACE_RECORDS][2];
struct TraceScope {
  int id;
  __device__ __forceinline__ explicit TraceScope(int id_) : id(id_) {
    if ((threadIdx.x & 31) == 0 && id >= 0 && id < TRACE_RECORDS) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMin(&g_trace[id][0], t);
    }
  }
  __device__ __forceinline__ ~TraceScope() {
    if ((threadIdx.x & 31) == 0 && id >= 0 && id < TRACE_RECORDS) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMax(&g_trace[id][1], t);
    }
  }
};
#define SO100_TRACE_SCOPE(id_) TraceScope trace_scope_(id_)
#else
#define SO100_TRACE_SCOPE(id_) do { } while (0)
#endif
enum { TR_KIN = 0, TR_BOX = 1, TR_HULL = 2, TR_LIGHT_A = 3, TR_LIGHT_B = 4, TR_MED_A = 5, TR_MED_B = 6, TR_HEAVY = 7, TR_TASK = 8, TR_SLOW = 9 };

struct Queues {
  int* ctl;      // [Q_WORDS] this stage's copy of the stage words
  int* ctl_next; // the other copy: re-armed by this stage's first kernel for the next stage
  int* lanectl;  // slow-lane words
  int* hull;     // [N * NHP] hull pairs for GJK/EPA this substep, one item = env * NHP + slot
  int* heavy;    // [N] envs with more than NCL contacts this substep (solved after K2b)
  int* medium_a; // [N] envs with at most NCL contacts, one of which couples the arm and the cube (dense Hessian); complete after K2a
  int* medium_b; // [N] the same, complete after K2b
  int* light_b;  // [N] envs of the light class (<= NCL contacts, no arm-cube contact) that had hull pairs: complete after K2b
  // Longest-first order of the light solve kernel: block b solves env order_in[b].  Envs that needed >= 3 Newton
  // iterations (or went to the heavy kernel) are written to the front of order_out, the rest to the back, so that the
  // next substep starts its likely stragglers first (the iteration count of an env is strongly correlated in time).
  const int* order_in;
  int* order_out;
  int* stat;     // [2] largest medium-queue / hull-pair-queue length (relative: count * 1024 / group size) seen since the host last
                 // cleared it; the host sizes the queue kernels' grids for the next step from it (so100_b200.cu: grid class)
  int scale;     // 1024 * 1024 / group size
  int trace;     // development builds (-DSO100_TRACE): base record id of this (group, stage)
  int split;     // 1: a / b work classes as described above; 0: K2b runs before every solve kernel, so there are no b classes
                 // (every env is solved by the regular light grid or the medium a / heavy queue)
  // ---- slow lane (so100_phases.cuh: phase_slow_lane).  The duration of every kernel of a stage is set by its slowest env, and a
  // stage's chain is the SUM of those maxima.  With the slow lane an env whose work exceeds a budget (Newton iterations of the
  // light solve, GJK / EPA iterations of one hull pair) or that belongs to a rare class (arm-cube contact, > NCL contacts) is
  // SUSPENDED: it leaves the regular kernels for the rest of the step and a slow-lane tile takes it through all remaining stages
  // on its own (kinematics -> collision -> solve -> integrate, same device functions and tile widths, hence the same bits),
  // beside the regular pipeline.  The regular chain then only carries budgeted work.
  int* slow;     // [N] slow-lane queue: suspended envs in order of suspension
  int* lane;     // [N] 1: the env is suspended (owned by the slow lane) for the rest of this step
  int slowlane;  // 0: classic schedule (medium / heavy queue kernels, no budgets)
  int budget_newton, budget_gjk, budget_epa;
  int pair_far;  // light solve kernel, two envs per warp: the warp's second tile takes its env from the far end of the longest-first order, so
                 // a likely straggler shares its warp with a likely one-iteration solve instead of with another straggler
  __device__ __forceinline__ bool suspended(int env) const { return slowlane && lane[env] != 0; }
  __device__ __forceinline__ void suspend(int env) const {
    if (atomicExch(&lane[env], 1) == 0) slow[atomicAdd(&lanectl[Q_LANE_COUNT], 1)] = env;
  }
  __device__ __forceinline__ void note(int which, int count) const { atomicMax(&stat[which], (count * scale) >> 10); }
  // work class of an env whose collision stage is complete (`after_hull`: completed by K2b, or had hull pairs when the
  // contact lists of the previous step's trailing stage are reused): > NCL contacts (or list overflow) -> heavy queue; an
  // arm-cube contact among <= NCL -> medium queue a / b; light envs completed by K2a are solved by the regular grid of the
  // light kernel, those completed by K2b go to light queue b
  __device__ __forceinline__ void route(int env, int ncon, bool coupled, bool after_hull) const {
    after_hull = after_hull && split;
    if (slowlane) { if (ncon > NCL || coupled) suspend(env); return; }
    if (ncon > NCL) heavy[atomicAdd(&ctl[Q_HEAVY_COUNT], 1)] = env;
    else if (coupled) {
      if (after_hull) medium_b[atomicAdd(&ctl[Q_MEDB_COUNT], 1)] = env;
      else medium_a[atomicAdd(&ctl[Q_MEDA_COUNT], 1)] = env;
    } else if (after_hull) light_b[atomicAdd(&ctl[Q_LB_COUNT], 1)] = env;
  }
};

// ---------------------------------------------------------------- scratch structs
struct FrameBlock {                // image of W_FRAMES
  float lpos[7][3];                // link origins: 6 arm links + cube
  float lmat[7][9];                // link axes (row-major)
  float axis[NL][3];               // hinge axes, world
  float fpad[2];
};
static_assert(sizeof(FrameBlock) == W_FRAMES_N * 4, "frame block layout");

struct DynBlock {                  // image of W_DYN
  float Mfull[NL][NL];             // arm mass matrix (symmetric, both triangles stored: row d feeds M a without index math)
  float qfs[NV];                   // qfrc_smooth
  float qas[NV];                   // unconstrained acceleration M^-1 qfrc_smooth
};
static_assert(sizeof(DynBlock) == W_DYN_N * 4, "dyn block layout");

// K1: kinematics + dynamics
struct __align__(16) KinS {
  float st[32];                    // qpos[13] qvel[12] ctrl[6]
  FrameBlock f;
  DynBlock d;
  float com[NL][3], Iw[NL][6], U[21][3], Y[21][3], FN[NL][6];
  float bank_pad[20];              // stride of 16 banks (mod 32) between the two 16-lane tiles of a warp
};
static_assert((sizeof(KinS) / 4) % 32 == 16, "KinS stride: half the banks");

// K2a: broad phase + box-like pairs
constexpr int NPEN = 32;           // penetrating box pairs per env that reach contact generation
struct __align__(16) BoxS {
  FrameBlock f;
  float4 gbox[NGEOM];              // world OBB centre + bounding radius of the collidable geoms
  float4 gext[NGEOM];              // world AABB half extents
  unsigned char qc[NPAIR_MAX];     // broad-phase survivors (pair ids)
  unsigned char q1[NPEN], qcode[NPEN];
  float qsep[NPEN];
};
static_assert((sizeof(BoxS) / 4) % 32 == 16, "BoxS stride: half the banks");

// K2b: GJK / EPA for hull pairs
struct __align__(16) HullS {
  FrameBlock f;
#ifdef SO100_HULL_CLOCK
  float epa[908];      // development build: + per-phase cycle counters
#else
  float epa[900];
#endif
};

// K3: constraint rows + Newton solve, NCAP contacts
template <int NCAP_> struct __align__(16) SolS {
  static constexpr int NCAP = NCAP_;
  double ad[NV];                   // qacc iterate in fp64 (jar cancellation, see so100_solve.cuh)
  float st[48];                    // qpos[13] qvel[12] ctrl[6] warm[12] (image of the state record head)
  FrameBlock f;
  DynBlock d;
  float a[NV];                     // qacc iterate
  float hdiag[NV];                 // active diagonal curvature of friction / limit rows
  float vec[NV];                   // gradient -> search direction
  float H[80];                     // packed Hessian (block diagonal: 42 entries, dense: 78)
  float con[NCAP_][CON_WORDS];     // image of the workspace contact list
  float cD[NCAP_][4], caref[NCAP_][4], cmu[NCAP_], cDm[NCAP_];
  float cfrc[NCAP_][4];
  float cH[NCAP_][10];
  unsigned char czone[NCAP_];
  unsigned char ckind[NCAP_];      // bit 0: contact touches an arm link, bit 1: touches the cube
  int ncon, coupled;
#ifdef SO100_SOLVE_CLOCK
  int clk[4], clk2[4];
#endif
  float J[NCAP_ * 4][JS];
  float T[NCAP_ * 4][JS];          // H_c J_c rows of the current Newton iteration
  float bank_pad[NCAP_ == NCL ? 8 : 4];   // the two 16-lane tiles of a warp sit in consecutive structs: a stride of 16 banks (mod 32) keeps
                                          // their lane-contiguous accesses on disjoint banks (ncu counted 0.56 M two-way conflicts per launch)
};
#ifndef SO100_SOLVE_CLOCK      // the development build adds eight counters to the struct
static_assert((sizeof(SolS<NCL>) / 4) % 32 == 16, "SolS<NCL> stride: half the banks");
#endif

// K4 / reset: task layer
struct __align__(16) TaskS {
  float st[STATE_WORDS];
  FrameBlock f;
};

// ---------------------------------------------------------------- tile helpers
template <unsigned LPE> __device__ __forceinline__ V3 shfl_up3(const Tile<LPE>& t, V3 v, int d) {
  return mk(t.shfl_up(v.x, d), t.shfl_up(v.y, d), t.shfl_up(v.z, d));
}
template <unsigned LPE> __device__ __forceinline__ float tsum(const Tile<LPE>& t, float v) {
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) v += t.shfl_xor(v, off);
  return v;
}
template <unsigned LPE> __device__ __forceinline__ void tsum3(const Tile<LPE>& t, float& a, float& b, float& c) {
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) {
    a += t.shfl_xor(a, off);
    b += t.shfl_xor(b, off);
    c += t.shfl_xor(c, off);
  }
}
template <unsigned LPE> __device__ __forceinline__ void tsum2(const Tile<LPE>& t, float& a, float& b) {
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) {
    a += t.shfl_xor(a, off);
    b += t.shfl_xor(b, off);
  }
}

// true on every lane of the WARP when any lane's predicate holds (all tiles of the warp must call it together)
template <unsigned LPE> __device__ __forceinline__ bool warp_any(const Tile<LPE>& t, bool p) {
  if (LPE == 32) return t.any(p);
  return __any_sync(0xffffffffu, p);
}

// Copy NW words (a multiple of 4, both sides 16-byte aligned) with one 128-bit load per lane and round;
// all loads are issued before the first store so their latencies overlap.
template <unsigned LPE, int NW> __device__ __forceinline__ void copy_vec(const Tile<LPE>& t, float* dst, const float* src) {
  static_assert(NW % 4 == 0, "copy_vec moves whole float4s");
  constexpr int NV4 = NW / 4, R = (NV4 + LPE - 1) / LPE;
  const float4* s4 = reinterpret_cast<const float4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  float4 v[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int k = t.thread_rank() + r * LPE;
    if (k < NV4) v[r] = s4[k];
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const int k = t.thread_rank() + r * LPE;
    if (k < NV4) d4[k] = v[r];
  }
}

// tile and env bookkeeping shared by all phase kernels (blockDim.x = TPB_)
#define SO100_TILE_PROLOGUE(LPE_, TPB_, STYPE_)                                \
  extern __shared__ __align__(16) unsigned char smem_raw[];                    \
  constexpr int EPB = TPB_ / LPE_;                                             \
  cg::thread_block blk = cg::this_thread_block();                              \
  Tile<LPE_> t = cg::tiled_partition<LPE_>(blk);                               \
  STYPE_* S = reinterpret_cast<STYPE_*>(smem_raw) + t.meta_group_rank();       \
  const int lane = t.thread_rank();                                            \
  (void)lane; (void)EPB

}  // namespace so100
