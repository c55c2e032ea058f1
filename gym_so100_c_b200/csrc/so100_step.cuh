// Fused bin-a-cube env step for sm_100a: one cooperative tile of LPE lanes per environment,
// all per-env working data in shared memory, warp shuffles for the small dense algebra.
//
// Per `step` launch and env:
//   action -> ctrl (constants.py:44-47,78-86)
//   10 x [ kinematics -> mass matrix -> RNE bias -> position actuators -> collision ->
//          constraint rows -> Newton solve (elliptic cones) -> semi-implicit Euler ]
//   trailing position stage (dm_control legacy step: mj_step1) -> reward / flags / obs
//   (single_arm.py:322-380, env.py:137-145, 372-406) -> optional same-call auto-reset.
// HBM is touched once in and once out per env and launch (a 256 B state record + I/O rows).
#pragma once
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include "so100_dev.cuh"

namespace so100 {
namespace cg = cooperative_groups;

__constant__ DevModel c_m;

#ifdef SO100_PROFILE
// per-stage SM-cycle counters (development builds only: -DSO100_PROFILE); slots:
// 0 kinematics+M, 1 bias/actuation, 2 collide total, 3 contact rows, 4 solve, 5 integrate, 6 broad phase, 7 box stage
__device__ unsigned long long g_prof[8];
#define PROF_MARK(k) do { long long now_ = clock64(); if (t.thread_rank() == 0) atomicAdd(&g_prof[k], (unsigned long long)(now_ - tp_)); tp_ = now_; } while (0)
#define PROF_BEGIN() long long tp_ = clock64()
#else
#define PROF_MARK(k) do { } while (0)
#define PROF_BEGIN() do { } while (0)
#endif

struct DevTables {
  const DevGeom* geom;
  const DevPair* pair;
  const float4* vert;
};

struct __align__(16) EnvS {
  double ad[NV];           // qacc iterate in fp64 (see so100_solve.cuh: jar cancellation)
  float st[STATE_WORDS];   // image of the HBM record: qpos qvel ctrl warm goal counters
  float lpos[7][3];        // link origins: 6 arm links + cube
  float lmat[7][9];        // link axes (row-major)
  float axis[NL][3];       // hinge axes, world
  float Marm[21];          // arm mass matrix, packed lower triangle
  float qfs[NV];           // qfrc_smooth
  float a[NV];             // qacc iterate
  float hdiag[NV];         // active diagonal curvature of friction / limit rows
  float vec[NV];           // gradient -> search direction
  float gcen[NGEOM][3];    // world OBB centres of the collidable geoms
  int ncon, nq1, nbox, nhull;
  float cpos[NC][3], cnrm[NC][3], cdist[NC];
  float cD[NC][4], caref[NC][4], cmu[NC];
  float cfrc[NC][4];
  float cH[NC][10];
  unsigned char cpair[NC];
  unsigned char czone[NC];
  unsigned char ckind[NC];   // bit 0: contact touches an arm link, bit 1: touches the cube
  union {
    struct { float com[NL][3], Iw[NL][6], U[21][3], Y[21][3], FN[NL][6]; } dyn;
    struct { float H[80]; } sol;
  } u;
  union {
    float J[NC * 4][JS];
    struct { unsigned char q1[NPAIR_MAX], qbox[NPAIR_MAX], qhull[NPAIR_MAX], qcode[64]; float qsep[64]; float epa[900]; } col;
  } w;
};

template <unsigned LPE> using Tile = cg::thread_block_tile<LPE>;

template <unsigned LPE> __device__ __forceinline__ V3 shfl_up3(const Tile<LPE>& t, V3 v, int d) {
  return mk(t.shfl_up(v.x, d), t.shfl_up(v.y, d), t.shfl_up(v.z, d));
}
template <unsigned LPE> __device__ __forceinline__ float tsum(const Tile<LPE>& t, float v) {
  return cg::reduce(t, v, cg::plus<float>());
}
// inclusive prefix sum over lanes 0..5 (other lanes carry garbage that never flows down)
template <unsigned LPE> __device__ __forceinline__ V3 scan6(const Tile<LPE>& t, V3 v, int lane) {
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    V3 o = shfl_up3(t, v, d);
    if (lane >= d) v = v + o;
  }
  return v;
}

// =====================================================================================
// position stage part 1: kinematics, mass matrix (App. A steps 1-2)
// =====================================================================================
template <unsigned LPE> __device__ void kinematics(const Tile<LPE>& t, EnvS* S) {
  const int lane = t.thread_rank();
  V3 p = mk(0, 0, 0);
  Q4 q = {1, 0, 0, 0};
  if (lane < NL) {
    float ang = S->st[S_QPOS + lane], sn, cs;
    sincosf(0.5f * ang, &sn, &cs);
    Q4 ql = {cs, c_m.link_axis[lane][0] * sn, c_m.link_axis[lane][1] * sn, c_m.link_axis[lane][2] * sn};
    Q4 qb = {c_m.link_quat[lane][0], c_m.link_quat[lane][1], c_m.link_quat[lane][2], c_m.link_quat[lane][3]};
    q = qmul(qb, ql);
    p = ld3(c_m.link_pos[lane]);
  }
  // prefix composition T_0 o ... o T_l over the serial chain (3 shuffle rounds instead of 6 serial links)
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    V3 po = shfl_up3(t, p, d);
    Q4 qo = {t.shfl_up(q.w, d), t.shfl_up(q.x, d), t.shfl_up(q.y, d), t.shfl_up(q.z, d)};
    if (lane >= d && lane < NL) {
      p = po + qrot(qo, p);
      q = qmul(qo, q);
    }
  }
  if (lane < NL) {
    Q4 qb = {c_m.base_quat[0], c_m.base_quat[1], c_m.base_quat[2], c_m.base_quat[3]};
    p = ld3(c_m.base_pos) + qrot(qb, p);
    q = qnormalize(qmul(qb, q));
  } else if (lane == NL) {
    p = ld3(&S->st[S_QPOS + 6]);
    Q4 qc = {S->st[S_QPOS + 9], S->st[S_QPOS + 10], S->st[S_QPOS + 11], S->st[S_QPOS + 12]};
    q = qnormalize(qc);
  }
  if (lane <= NL) {
    float R[9];
    q2mat(q, R);
    st3(S->lpos[lane], p);
#pragma unroll
    for (int k = 0; k < 9; k++) S->lmat[lane][k] = R[k];
    if (lane < NL) {
      st3(S->axis[lane], mulmv(R, ld3(c_m.link_axis[lane])));
      st3(S->u.dyn.com[lane], p + mulmv(R, ld3(c_m.link_ipos[lane])));
      // Iw = R Ib R^T
      const float* I = c_m.link_Ib[lane];
      float T[9];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        T[r * 3 + 0] = R[r * 3] * I[0] + R[r * 3 + 1] * I[1] + R[r * 3 + 2] * I[2];
        T[r * 3 + 1] = R[r * 3] * I[1] + R[r * 3 + 1] * I[3] + R[r * 3 + 2] * I[4];
        T[r * 3 + 2] = R[r * 3] * I[2] + R[r * 3 + 1] * I[4] + R[r * 3 + 2] * I[5];
      }
      float* Iw = S->u.dyn.Iw[lane];
      Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
      Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
      Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
      Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
      Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
      Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
    }
  }
  t.sync();
}

// u_il = a_i x (c_l - o_i), y_il = I_l a_i for l >= i; then M_ij = sum_{l>=i} m_l u_il.u_jl + a_j.y_il
template <unsigned LPE> __device__ void mass_matrix(const Tile<LPE>& t, EnvS* S) {
  const int lane = t.thread_rank();
  for (int e = lane; e < 21; e += LPE) {
    int l = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    if (tri(l + 1, 0) <= e) l++;
    if (tri(l, 0) > e) l--;
    int i = e - tri(l, 0);
    V3 ai = ld3(S->axis[i]);
    st3(S->u.dyn.U[e], cross(ai, ld3(S->u.dyn.com[l]) - ld3(S->lpos[i])));
    st3(S->u.dyn.Y[e], symv(S->u.dyn.Iw[l], ai));
  }
  t.sync();
  for (int e = lane; e < 21; e += LPE) {
    int i = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    if (tri(i + 1, 0) <= e) i++;
    if (tri(i, 0) > e) i--;
    int j = e - tri(i, 0);
    V3 aj = ld3(S->axis[j]);
    float s = (i == j) ? c_m.armature[i] : 0.0f;
    for (int l = i; l < NL; l++)
      s += c_m.link_mass[l] * dot(ld3(S->u.dyn.U[tri(l, i)]), ld3(S->u.dyn.U[tri(l, j)])) + dot(aj, ld3(S->u.dyn.Y[tri(l, i)]));
    S->Marm[e] = s;
  }
  // no sync: consumers sync before reading Marm
}

// =====================================================================================
// velocity stage: RNE bias via prefix scans, actuators, qfrc_smooth (App. A steps 4-6)
// =====================================================================================
template <unsigned LPE> __device__ void smooth_forces(const Tile<LPE>& t, EnvS* S) {
  const int lane = t.thread_rank();
  V3 ax = mk(0, 0, 0), o = mk(0, 0, 0);
  float qd = 0;
  if (lane < NL) { ax = ld3(S->axis[lane]); o = ld3(S->lpos[lane]); qd = S->st[S_QVEL + lane]; }
  V3 w = scan6(t, ax * qd, lane);                       // omega_l
  V3 wp = shfl_up3(t, w, 1);
  if (lane == 0) wp = mk(0, 0, 0);
  V3 al = scan6(t, cross(wp, ax) * qd, lane);           // alpha_l (qacc = 0)
  V3 alp = shfl_up3(t, al, 1);
  V3 op = shfl_up3(t, o, 1);
  if (lane == 0) { alp = mk(0, 0, 0); op = o; }
  V3 r = o - op;
  V3 ao = scan6(t, cross(alp, r) + cross(wp, cross(wp, r)), lane);   // origin acceleration
  if (lane < NL) {
    ao = ao - mk(c_m.gx, c_m.gy, c_m.gz);
    V3 c = ld3(S->u.dyn.com[lane]) - o;
    V3 ac = ao + cross(al, c) + cross(w, cross(w, c));
    V3 F = ac * c_m.link_mass[lane];
    V3 N = symv(S->u.dyn.Iw[lane], al) + cross(w, symv(S->u.dyn.Iw[lane], w));
    st3(&S->u.dyn.FN[lane][0], F);
    st3(&S->u.dyn.FN[lane][3], N);
  }
  t.sync();
  if (lane < NL) {
    float bias = 0;
    V3 ai = ld3(S->axis[lane]);
    for (int l = lane; l < NL; l++)
      bias += dot(ld3(S->u.dyn.U[tri(l, lane)]), ld3(&S->u.dyn.FN[l][0])) + dot(ai, ld3(&S->u.dyn.FN[l][3]));
    // position actuator: clip(kp (clip(ctrl) - q) - kv qd)
    float u = fminf(fmaxf(S->st[S_CTRL + lane], c_m.ctrl_lo[lane]), c_m.ctrl_hi[lane]);
    float f = c_m.kp[lane] * u - c_m.kp[lane] * S->st[S_QPOS + lane] - c_m.kv[lane] * qd;
    f = fminf(fmaxf(f, c_m.frc_lo[lane]), c_m.frc_hi[lane]);
    S->qfs[lane] = f - bias;
  } else if (lane < NV) {
    int k = lane - NL;
    float b;
    if (k < 3) {
      b = -c_m.cube_mass * (k == 0 ? c_m.gx : (k == 1 ? c_m.gy : c_m.gz));
    } else {   // gyroscopic torque in the body frame (zero for the isotropic cube)
      V3 wl = ld3(&S->st[S_QVEL + 9]);
      V3 Iw = mk(c_m.cube_I[0] * wl.x, c_m.cube_I[1] * wl.y, c_m.cube_I[2] * wl.z);
      b = comp(cross(wl, Iw), k - 3);
    }
    S->qfs[lane] = -b;
  }
  t.sync();
}

// M a for dof d (arm block dense, cube block diagonal)
__device__ __forceinline__ float mul_M(const EnvS* S, const float* a, int d) {
  if (d < NL) {
    float s = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) s = fmaf(S->Marm[d >= j ? tri(d, j) : tri(j, d)], a[j], s);
    return s;
  }
  return (d < 9 ? c_m.cube_mass : c_m.cube_I[d - 9]) * a[d];
}

// qacc_smooth = M^-1 qfrc_smooth; every lane factors the 6x6 arm block redundantly in registers
template <unsigned LPE> __device__ float smooth_acc(const Tile<LPE>& t, const EnvS* S) {
  const int lane = t.thread_rank();
  float L[21], x[NL];
#pragma unroll
  for (int e = 0; e < 21; e++) L[e] = S->Marm[e];
#pragma unroll
  for (int i = 0; i < NL; i++) x[i] = S->qfs[i];
#pragma unroll
  for (int j = 0; j < NL; j++) {
    float d = L[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
    d = rsqrtf(fmaxf(d, 1e-20f));
    L[tri(j, j)] = d;   // stores 1/L_jj
#pragma unroll
    for (int i = j + 1; i < NL; i++) {
      float s = L[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[tri(i, k)] * L[tri(j, k)];
      L[tri(i, j)] = s * d;
    }
  }
#pragma unroll
  for (int i = 0; i < NL; i++) {
    float s = x[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[tri(i, k)] * x[k];
    x[i] = s * L[tri(i, i)];
  }
#pragma unroll
  for (int i = NL - 1; i >= 0; i--) {
    float s = x[i];
#pragma unroll
    for (int k = i + 1; k < NL; k++) s -= L[tri(k, i)] * x[k];
    x[i] = s * L[tri(i, i)];
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) if (lane == i) r = x[i];
  if (lane >= NL && lane < NV) r = S->qfs[lane] / (lane < 9 ? c_m.cube_mass : c_m.cube_I[lane - 9]);
  return r;   // lane d < 12 holds qacc_smooth[d]
}

// =====================================================================================
// collision (App. A step 3)
// =====================================================================================
struct Obb { V3 c, ax[3]; float h[3]; };
// Edge-edge axes A_i x B_j: skipped below sin^2 = 1e-6 and penalised by 2e-6 m / sin so that
// round-off on nearly parallel edges can never beat a face axis (same rule in the oracle).
constexpr float EDGE_MIN_SIN2 = 1e-6f;
constexpr float EDGE_BIAS = 2e-6f;

__device__ __forceinline__ void load_obb(const EnvS* S, const DevGeom& g, int gi, Obb& b) {
  b.c = ld3(S->gcen[gi]);
  b.h[0] = g.half[0]; b.h[1] = g.half[1]; b.h[2] = g.half[2];
  const float* m = g.link >= 0 ? S->lmat[g.link] : g.wmat;
  b.ax[0] = mcol(m, 0); b.ax[1] = mcol(m, 1); b.ax[2] = mcol(m, 2);
}

// distance^2 from point p to an oriented box
__device__ __forceinline__ float point_obb_d2(V3 p, const Obb& b) {
  V3 d = p - b.c;
  float s = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float x = dot(d, b.ax[k]);
    float e = fmaxf(fabsf(x) - b.h[k], 0.0f);
    s = fmaf(e, e, s);
  }
  return s;
}

// Box-box narrow phase, one lane per pair.  Separating axes (15) with the classic R-matrix form;
// face contact: incident face clipped in the 2-D frame of the reference face, up to 8 points.
// Returns the number of points; normal points from A to B; dist < 0.
__device__ int box_box(const Obb& A, const Obb& B, float (*cp)[3], float* cd, V3& normal) {
  float R[3][3], aR[3][3], tA[3], tB[3];
  V3 t = B.c - A.c;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    tA[i] = dot(t, A.ax[i]);
    tB[i] = dot(t, B.ax[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) { R[i][j] = dot(A.ax[i], B.ax[j]); aR[i][j] = fabsf(R[i][j]); }
  }
  float best_face = -1e30f; int code = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    float sep = fabsf(tA[i]) - (A.h[i] + B.h[0] * aR[i][0] + B.h[1] * aR[i][1] + B.h[2] * aR[i][2]);
    if (sep > 0) return 0;
    if (sep > best_face) { best_face = sep; code = i; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float sep = fabsf(tB[j]) - (B.h[j] + A.h[0] * aR[0][j] + A.h[1] * aR[1][j] + A.h[2] * aR[2][j]);
    if (sep > 0) return 0;
    if (sep > best_face) { best_face = sep; code = 3 + j; }
  }
  float best_edge = -1e30f, best_sel = -1e30f; int ei = -1, ej = -1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      float l2 = 1.0f - R[i][j] * R[i][j];
      if (l2 < EDGE_MIN_SIN2) continue;      // (near-)parallel edges: the face axes cover this direction
      float inv = rsqrtf(l2);
      float ra = A.h[i1] * aR[i2][j] + A.h[i2] * aR[i1][j];
      float rb = B.h[j1] * aR[i][j2] + B.h[j2] * aR[i][j1];
      float sep = (fabsf(tA[i2] * R[i1][j] - tA[i1] * R[i2][j]) - (ra + rb)) * inv;
      if (sep > 0) return 0;
      float sel = sep - EDGE_BIAS * inv;
      if (sel > best_sel) { best_sel = sel; best_edge = sep; ei = i; ej = j; }
    }
  }
  if (ei >= 0 && best_sel * 1.05f > best_face) {
    // edge-edge: one point midway between the closest points of the two edges
    V3 n = normalized(cross(A.ax[ei], B.ax[ej]));
    if (dot(n, t) < 0) n = -n;
    V3 pA = A.c, pB = B.c;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k != ei) pA = pA + A.ax[k] * (dot(n, A.ax[k]) > 0 ? A.h[k] : -A.h[k]);
      if (k != ej) pB = pB + B.ax[k] * (dot(n, B.ax[k]) > 0 ? -B.h[k] : B.h[k]);
    }
    V3 r = pA - pB;
    float b = dot(A.ax[ei], B.ax[ej]), d = dot(A.ax[ei], r), e = dot(B.ax[ej], r);
    float den = 1.0f - b * b;
    float s = (b * e - d) / den, u = (e - b * d) / den;
    V3 qa = pA + A.ax[ei] * s, qb = pB + B.ax[ej] * u;
    st3(cp[0], (qa + qb) * 0.5f);
    cd[0] = best_edge;
    normal = n;
    return 1;
  }
  // face contact
  const bool refA = code < 3;
  const Obb& Rf = refA ? A : B;
  const Obb& If = refA ? B : A;
  const int ax = refA ? code : code - 3;
  const int ua = (ax + 1) % 3, va = (ax + 2) % 3;
  V3 tri_ = If.c - Rf.c;
  float sgn = dot(Rf.ax[ax], tri_) < 0 ? -1.0f : 1.0f;       // outward normal = sgn * axis
  V3 nref = Rf.ax[ax] * sgn;
  int iax = 0; float bd = -1;
#pragma unroll
  for (int k = 0; k < 3; k++) { float d = fabsf(dot(If.ax[k], nref)); if (d > bd) { bd = d; iax = k; } }
  float isg = dot(If.ax[iax], nref) > 0 ? -1.0f : 1.0f;
  V3 fc = If.c + If.ax[iax] * (isg * If.h[iax]) - Rf.c;       // incident face centre rel. reference centre
  const int iu = (iax + 1) % 3, iv = (iax + 2) % 3;
  // polygon in reference coordinates (u, v, n)
  float P[2][8][3];
  int np = 4, cur = 0;
  const float su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    V3 p = fc + If.ax[iu] * (su[q] * If.h[iu]) + If.ax[iv] * (sv[q] * If.h[iv]);
    P[0][q][0] = dot(p, Rf.ax[ua]); P[0][q][1] = dot(p, Rf.ax[va]); P[0][q][2] = dot(p, nref);
  }
  // clip against +-h_u, +-h_v (Sutherland-Hodgman in 2-D, depth interpolated)
  for (int pl = 0; pl < 4 && np > 0; pl++) {
    const int cdim = pl >> 1;
    const float s = (pl & 1) ? -1.0f : 1.0f;
    const float lim = cdim == 0 ? Rf.h[ua] : Rf.h[va];
    int no = 0;
    for (int i = 0; i < np; i++) {
      const float* a = P[cur][i];
      const float* b = P[cur][(i + 1 == np) ? 0 : i + 1];
      float da = s * a[cdim] - lim, db = s * b[cdim] - lim;
      if (da <= 0 && no < 8) { P[cur ^ 1][no][0] = a[0]; P[cur ^ 1][no][1] = a[1]; P[cur ^ 1][no][2] = a[2]; no++; }
      if (((da < 0 && db > 0) || (da > 0 && db < 0)) && no < 8) {
        float tt = da / (da - db);
        P[cur ^ 1][no][0] = a[0] + tt * (b[0] - a[0]);
        P[cur ^ 1][no][1] = a[1] + tt * (b[1] - a[1]);
        P[cur ^ 1][no][2] = a[2] + tt * (b[2] - a[2]);
        no++;
      }
    }
    np = no; cur ^= 1;
  }
  int nc = 0;
  for (int q = 0; q < np; q++) {
    float depth = Rf.h[ax] - P[cur][q][2];
    if (depth <= 0) continue;
    V3 p = Rf.c + Rf.ax[ua] * P[cur][q][0] + Rf.ax[va] * P[cur][q][1] + nref * (P[cur][q][2] + 0.5f * depth);
    st3(cp[nc], p);
    cd[nc] = -depth;
    nc++;
  }
  normal = refA ? nref : -nref;
  return nc;
}

template <unsigned LPE> __device__ void hull_stage(const Tile<LPE>& t, EnvS* S, const DevTables& T, int nhull);

template <unsigned LPE> __device__ void collide(const Tile<LPE>& t, EnvS* S, const DevTables& T) {
  const int lane = t.thread_rank();
  PROF_BEGIN();
  // world OBB centres
  for (int g = lane; g < c_m.ngeom; g += LPE) {
    const DevGeom& G = T.geom[g];
    V3 c = ld3(G.center);
    if (G.link >= 0) c = ld3(S->lpos[G.link]) + mulmv(S->lmat[G.link], c);
    st3(S->gcen[g], c);
  }
  if (lane == 0) { S->ncon = 0; }
  t.sync();
  // stage 1: bounding sphere vs sphere, sphere vs oriented box (both ways); compact survivors by mode
  int nbox = 0, nhull = 0;
  for (int base = 0; base < c_m.npair; base += LPE) {
    const int p = base + lane;
    int pass = 0, mode = 0;
    if (p < c_m.npair) {
      const DevPair& P = T.pair[p];
      const DevGeom& G1 = T.geom[P.g1];
      const DevGeom& G2 = T.geom[P.g2];
      V3 c1 = ld3(S->gcen[P.g1]), c2 = ld3(S->gcen[P.g2]);
      V3 d = c2 - c1;
      float rr = G1.rbound + G2.rbound;
      if (dot(d, d) <= rr * rr) {
        Obb b1, b2;
        load_obb(S, G1, P.g1, b1);
        load_obb(S, G2, P.g2, b2);
        if (point_obb_d2(c1, b2) <= G1.rbound * G1.rbound && point_obb_d2(c2, b1) <= G2.rbound * G2.rbound) {
          pass = 1; mode = P.mode;
        }
      }
    }
    unsigned mb = t.ballot(pass && mode != MODE_HULL), mh = t.ballot(pass && mode == MODE_HULL);
    unsigned lt = (1u << lane) - 1u;
    if (pass) {
      if (mode != MODE_HULL) S->w.col.qbox[nbox + __popc(mb & lt)] = (unsigned char)p;
      else S->w.col.qhull[nhull + __popc(mh & lt)] = (unsigned char)p;
    }
    nbox += __popc(mb); nhull += __popc(mh);
  }
  t.sync();
  PROF_MARK(6);
  // stage 2a: box-like pairs, one lane per pair, deterministic append order (pair order)
  for (int base = 0; base < nbox; base += LPE) {
    const int k = base + lane;
    float cp[8][3], cd[8];
    V3 n = mk(0, 0, 1);
    int nc = 0, p = 0;
    if (k < nbox) {
      p = S->w.col.qbox[k];
      const DevPair& P = T.pair[p];
      Obb A, B;
      load_obb(S, T.geom[P.g1], P.g1, A);
      load_obb(S, T.geom[P.g2], P.g2, B);
      nc = box_box(A, B, cp, cd, n);
      if (nc > 1 && P.mode == MODE_BOX_SINGLE) {
        // mjc_Convex semantics (one contact per pair): deepest feature, centroid if not unique
        float dmin = cd[0];
        for (int i = 1; i < nc; i++) dmin = fminf(dmin, cd[i]);
        V3 acc = mk(0, 0, 0); int cnt = 0;
        for (int i = 0; i < nc; i++)
          if (cd[i] <= dmin + 1e-6f) { acc = acc + ld3(cp[i]); cnt++; }
        st3(cp[0], acc * (1.0f / cnt));
        cd[0] = dmin; nc = 1;
      }
    }
    // exclusive prefix of nc over the tile
    int incl = nc;
#pragma unroll
    for (int d = 1; d < LPE; d <<= 1) { int o = t.shfl_up(incl, d); if (lane >= d) incl += o; }
    int off = S->ncon + incl - nc;
    int total = t.shfl(incl, LPE - 1);
    for (int i = 0; i < nc; i++) {
      int c = off + i;
      if (c < NC) {
        st3(S->cpos[c], ld3(cp[i])); st3(S->cnrm[c], n); S->cdist[c] = cd[i]; S->cpair[c] = (unsigned char)p;
      }
    }
    t.sync();
    if (lane == 0) S->ncon = min(S->ncon + total, NC + 1);   // NC+1 marks overflow
    t.sync();
  }
  PROF_MARK(7);
  // stage 2b: pairs that involve a general hull (so100_gjk.cuh)
  hull_stage(t, S, T, nhull);
}

}  // namespace so100
