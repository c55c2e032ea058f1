// Constraint rows + primal Newton solver with elliptic cones + semi-implicit Euler
// (SURVEY.md Appendix A steps 7-9), one tile per env.  Work ownership inside the tile:
//   lane d < 12       dof d: its frictionloss row and (d < 6) its joint-limit row, in registers
//   lane r (strided)  contact row r = 4 c + k: the fp64 dot J_r . qacc; the quad leader (k = 0)
//                     gathers the 4 rows of contact c and owns its cone state / jar / jv
//   lane e (strided)  packed Hessian entry e
// The Hessian is block diagonal (arm 6x6 | cube 6x6) unless a contact joins an arm link and the
// cube; the common uncoupled case factors both blocks in registers (no barriers); the coupled case
// (solve<DENSE = true>, heavy kernel only) eliminates the arm block and factors the 6x6 Schur complement.
//
// The cost/force evaluation and the line-search derivative each have exactly ONE call site (the start-point
// selection, the Newton loop and the final force refresh all run through the same loop body): the kernel is
// instruction-fetch sensitive (ncu: stall_no_instruction), so code size is kept down on purpose.
// The scratch type ES is SolS<8> (light kernel) or SolS<24> (heavy kernel), see so100_scratch.cuh.
#pragma once
#include "so100_dyn.cuh"

namespace so100 {

constexpr int S_DIAG = 49;         // per-env diagnostic counters inside the state record (uint32 words 49..56)
constexpr int S_EPRET = 57;        // float: return of the running episode
constexpr int S_RETSUM = 58;       // float: sum of the returns of this env's finished episodes
constexpr int S_LENSUM = 59;       // uint32: sum of their lengths (the -DSO100_SOLVE_CLOCK development build overwrites 57..63)
#ifndef SO100_NEWTON_MAXIT
#define SO100_NEWTON_MAXIT 100
#endif
#ifndef SO100_LS_MAXIT
#define SO100_LS_MAXIT 50
#endif
// MuJoCo's defaults, which the scene's XML leaves untouched (so_arm100.xml:4 only sets cone / impratio): iterations 100,
// ls_iterations 50.  Both loops are data-dependent (`#pragma unroll 1`), so the caps cost nothing until a solve reaches them:
// warm-started solves need 1-3 Newton iterations and 1-3 line-search iterates.
constexpr int NEWTON_MAXIT = SO100_NEWTON_MAXIT;
constexpr int LS_MAXIT = SO100_LS_MAXIT;
#ifndef SO100_GTOL
#define SO100_GTOL 2e-6f     // gradient tolerance relative to |qfrc_smooth| + |J^T f|
#endif
#ifndef SO100_LS_TOL
#define SO100_LS_TOL 1e-2f   // line search stops at |phi'| <= LS_TOL |phi'(0)|: MuJoCo's default ls_tolerance (0.01)
#endif
#ifndef SO100_ITOL
#define SO100_ITOL 1e-9f     // predicted-improvement tolerance relative to the cost
#endif

__device__ __forceinline__ float impedance(const float* si, float dist) {
  if (si[0] == si[1] || si[2] <= 1e-15f) return 0.5f * (si[0] + si[1]);
  float x = fabsf(dist) / si[2];
  if (x >= 1.0f) return si[1];
  if (x <= 0.0f) return si[0];
  float y, p = si[4], mid = si[3];
  if (p == 1.0f) y = x;
  else if (x <= mid) y = __powf(x, p) / __powf(mid, p - 1.0f);
  else y = 1.0f - __powf(1.0f - x, p) / __powf(1.0f - mid, p - 1.0f);
  return si[0] + y * (si[1] - si[0]);
}

// interpolation weight y of the impedance sigmoid: imp = solimp[0] + y (solimp[1] - solimp[0])
__device__ __forceinline__ float impedance_y(const float* si, float dist) {
  if (si[0] == si[1]) return 0.0f;
  if (si[2] <= 1e-15f) return 0.5f;
  float x = fabsf(dist) / si[2];
  if (x >= 1.0f) return 1.0f;
  if (x <= 0.0f) return 0.0f;
  const float p = si[4], mid = si[3];
  if (p == 1.0f) return x;
  if (p == 2.0f) return x <= mid ? x * x / mid : 1.0f - (1.0f - x) * (1.0f - x) / (1.0f - mid);   // MuJoCo default power
  if (x <= mid) return powf(x, p) / powf(mid, p - 1.0f);
  return 1.0f - powf(1.0f - x, p) / powf(1.0f - mid, p - 1.0f);
}

// mju_makeFrame: tangents of a unit normal
__device__ __forceinline__ void make_frame(V3 n, V3& t1, V3& t2) {
  V3 y = (n.y < 0.5f && n.y > -0.5f) ? mk(0, 1, 0) : mk(0, 0, 1);
  y = y - n * dot(n, y);
  t1 = normalized(y);
  t2 = cross(n, t1);
}

// packed lower-triangle index -> (i, j), e < 78
__device__ __forceinline__ void untri(int e, int& i, int& j) {
  i = (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15) + (e >= 21) + (e >= 28) + (e >= 36) + (e >= 45) + (e >= 55) + (e >= 66);
  j = e - tri(i, 0);
}

// M a for dof d (arm block dense, cube block diagonal)
template <class ES> __device__ __forceinline__ float mul_M(const ES* S, const float* a, int d) {
  if (d < NL) {
    float s = 0;
#pragma unroll
    for (int j = 0; j < NL; j++) s = fmaf(S->d.Mfull[d][j], a[j], s);
    return s;
  }
  return (d < 9 ? c_m.cube_mass : c_m.cube_I[d - 9]) * a[d];
}

// ------------------------------------------------------------------ contact rows
template <unsigned LPE, class ES> __device__ void make_contact_rows(const Tile<LPE>& t, ES* S, const DevTables& T) {
  const int lane = t.thread_rank();
  const int ncon = S->ncon;
  const float rs_imp = rsqrtf(fmaxf(c_m.impratio, 1e-15f));
  int both = 0;
  for (int c = lane; c < ncon; c += LPE) {
    const DevPair& P = T.pair[__float_as_int(S->con[c][7])];
    const int l1 = P.l1, l2 = P.l2;
    const int kind = (((l1 >= 0 && l1 < NL) || (l2 >= 0 && l2 < NL)) ? 1 : 0) | ((l1 == NL || l2 == NL) ? 2 : 0);
    S->ckind[c] = (unsigned char)kind;
    both |= (kind == 3);
    const float dist = S->con[c][6];
    // imp = d0 + y (d1 - d0); 1 - imp is formed from the host's fp64 (1 - d0), not as 1.0f - imp
    const float y = impedance_y(P.solimp, dist);
    const float imp = fmaf(y, P.dd, P.solimp[0]), omi = fmaf(-y, P.dd, P.omd0);
    const float R0 = fmaxf(omi / imp * P.dtran, 1e-15f);
    const float R1 = R0 / fmaxf(c_m.impratio, 1e-15f);
    S->cD[c][0] = 1.0f / R0;
    S->cD[c][1] = 1.0f / R1;
    S->cD[c][2] = 1.0f / R1;
    S->cD[c][3] = (P.f1 * P.f1) / (R1 * P.f0 * P.f0);
    S->cmu[c] = P.f0 * rs_imp;                 // friction[0] * sqrt(R1 / R0)
    {
      const float mu = P.f0 * rs_imp;
      S->cDm[c] = (1.0f / R0) / fmaxf(mu * mu * (1.0f + mu * mu), 1e-15f);   // middle-zone stiffness
    }
    S->caref[c][0] = -P.K * imp * dist;
    S->caref[c][1] = S->caref[c][2] = S->caref[c][3] = 0.0f;
  }
  S->coupled = t.any(both) ? 1 : 0;            // arm and cube blocks of the Hessian are coupled
  for (int it = lane; it < ncon * NV; it += LPE) {
    const int c = it / NV, d = it - c * NV;
    const DevPair& P = T.pair[__float_as_int(S->con[c][7])];
    const int l1 = P.l1, l2 = P.l2;
    V3 n = ld3(&S->con[c][3]), t1, t2, pos = ld3(&S->con[c][0]);
    make_frame(n, t1, t2);
    V3 jp = mk(0, 0, 0), jr = mk(0, 0, 0);
    if (d < NL) {
      float s = ((l2 >= d && l2 < NL) ? 1.0f : 0.0f) - ((l1 >= d && l1 < NL) ? 1.0f : 0.0f);
      if (s != 0.0f) {
        V3 ax = ld3(S->f.axis[d]);
        jr = ax * s;
        jp = cross(ax, pos - ld3(S->f.lpos[d])) * s;
      }
    } else {
      float s = (l2 == NL ? 1.0f : 0.0f) - (l1 == NL ? 1.0f : 0.0f);
      if (s != 0.0f) {
        const int k = d - NL;
        if (k < 3) {
          jp = mk(k == 0 ? s : 0.0f, k == 1 ? s : 0.0f, k == 2 ? s : 0.0f);
        } else {
          V3 col = mcol(S->f.lmat[NL], k - 3);
          jr = col * s;
          jp = cross(col, pos - ld3(S->f.lpos[NL])) * s;
        }
      }
    }
    S->J[c * 4 + 0][d] = dot(n, jp);
    S->J[c * 4 + 1][d] = dot(t1, jp);
    S->J[c * 4 + 2][d] = dot(t2, jp);
    S->J[c * 4 + 3][d] = P.dim > 3 ? dot(n, jr) : 0.0f;
  }
  t.sync();
  for (int it = lane; it < ncon * 4; it += LPE) {
    const int c = it >> 2, k = it & 3;
    const DevPair& P = T.pair[__float_as_int(S->con[c][7])];
    float v = 0;
#pragma unroll 4
    for (int d = 0; d < NV; d++) v = fmaf(S->J[it][d], S->st[S_QVEL + d], v);
    S->caref[c][k] -= P.B * v;
  }
  t.sync();
}

// elliptic cone: cost, force and (when hess) the 4x4 Hessian block, packed lower triangle
__device__ __forceinline__ float cone_eval(bool hess, const float* x, const float* D, float Dm, float mu, float f0, float f1, int dim,
                                           float* force, int& zone, float* Hc) {
  const float fr[3] = {f0, f0, f1};
  float U[4];
  U[0] = x[0] * mu;
  float TT = 0;
#pragma unroll
  for (int j = 1; j < 4; j++) { U[j] = (j < dim) ? x[j] * fr[j - 1] : 0.0f; TT = fmaf(U[j], U[j], TT); }
  const float invT = TT > 0.0f ? rsqrtf(TT) : 0.0f;
  const float Tn = TT * invT, N = U[0];
  if (N >= mu * Tn || (Tn <= 0.0f && N >= 0.0f)) {
    zone = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) force[j] = 0.0f;
    return 0.0f;
  }
  if (mu * N + Tn <= 0.0f || (Tn <= 0.0f && N < 0.0f)) {
    zone = 1;
    float cost = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      float dj = (j < dim) ? D[j] : 0.0f;
      cost = fmaf(0.5f * dj * x[j], x[j], cost);
      force[j] = -dj * x[j];
    }
    if (hess) {
#pragma unroll
      for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b <= a; b++) Hc[tri(a, b)] = (a == b && a < dim) ? D[a] : 0.0f;
    }
    return cost;
  }
  zone = 2;
  const float NmT = N - mu * Tn;
  force[0] = -Dm * NmT * mu;
#pragma unroll
  for (int j = 1; j < 4; j++) force[j] = (j < dim) ? -force[0] * invT * U[j] * fr[j - 1] : 0.0f;
  if (hess) {
    float g[4], wv[4];
    g[0] = mu; wv[0] = 0;
#pragma unroll
    for (int j = 1; j < 4; j++) { wv[j] = (j < dim) ? fr[j - 1] * U[j] : 0.0f; g[j] = -mu * wv[j] * invT; }
    const float k1 = -NmT * mu * invT, k2 = NmT * mu * invT * invT * invT;
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b <= a; b++) {
        float h = g[a] * g[b];
        if (a > 0 && b > 0) {
          h += k2 * wv[a] * wv[b];
          if (a == b && a < dim) h += k1 * fr[a - 1] * fr[a - 1];
        }
        Hc[tri(a, b)] = Dm * h;
      }
  }
  return 0.5f * Dm * NmT * NmT;
}

// Line-search view of one contact: everything in the directional derivatives of the cone cost along x0 + alpha v
// that does not depend on alpha, formed once per Newton iteration by the contact's quad leader.
struct ConeLs {
  float N0, Np;          // normal coordinate mu x0[0] and its slope
  float U0[3], V[3];     // scaled tangential coordinates fr_j x0[j] and their slopes (zero beyond condim)
  float VV;              // |V|^2
  float A1, A2;          // bottom zone: sum D_j x0_j v_j, sum D_j v_j^2
  float Dm, mu;
};

__device__ __forceinline__ void cone_ls_prepare(const float* x0, const float* v, const float* D, float Dm, float mu, float f0, float f1,
                                                int dim, ConeLs& q) {
  const float fr[3] = {f0, f0, f1};
  q.mu = mu;
  q.N0 = x0[0] * mu; q.Np = v[0] * mu;
  q.VV = 0;
  q.A1 = D[0] * x0[0] * v[0]; q.A2 = D[0] * v[0] * v[0];
#pragma unroll
  for (int j = 1; j < 4; j++) {
    const float f = (j < dim) ? fr[j - 1] : 0.0f, dj = (j < dim) ? D[j] : 0.0f;
    q.U0[j - 1] = x0[j] * f; q.V[j - 1] = v[j] * f;
    q.VV = fmaf(q.V[j - 1], q.V[j - 1], q.VV);
    q.A1 = fmaf(dj * x0[j], v[j], q.A1); q.A2 = fmaf(dj * v[j], v[j], q.A2);
  }
  q.Dm = Dm;
}

// first / second directional derivative of the cone cost at x0 + alpha v
__device__ __forceinline__ void cone_ls(const ConeLs& q, float alpha, float& d1, float& d2) {
  const float N = fmaf(alpha, q.Np, q.N0);
  float TT = 0, UV = 0;
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const float U = fmaf(alpha, q.V[j], q.U0[j]);
    TT = fmaf(U, U, TT); UV = fmaf(U, q.V[j], UV);
  }
  // 1 / |U| from one MUFU.RSQ (2 ulp): the line search only steers alpha, the residual is evaluated exactly elsewhere
  const float rT = TT > 0.0f ? rsqrtf(TT) : 0.0f;
  const float Tn = TT * rT, mu = q.mu;
  if (N >= mu * Tn || (Tn <= 0.0f && N >= 0.0f)) return;
  if (mu * N + Tn <= 0.0f || (Tn <= 0.0f && N < 0.0f)) {
    d1 += fmaf(alpha, q.A2, q.A1);
    d2 += q.A2;
    return;
  }
  const float Tp = UV * rT, Tpp = (q.VV - Tp * Tp) * rT, NmT = N - mu * Tn, dp = q.Np - mu * Tp;
  d1 = fmaf(q.Dm * NmT, dp, d1);
  d2 += q.Dm * (dp * dp - NmT * mu * Tpp);
}

// ------------------------------------------------------------------ Newton solver
template <unsigned LPE, int NCAP> struct SolveRegs {
  static constexpr int RPL = (NCAP * 4 + LPE - 1) / LPE;   // contact rows (and, on quad leaders, contacts) per lane
  float qfs, fr_aref, fr_R, fr_D, fr_fl;
  float lim_sgn, lim_D, lim_aref;
  float jar[RPL][4];
  ConeLs ls[RPL];
  int dim[RPL];
  float f0[RPL], f1[RPL];
};

// cost at S->ad (and forces / cone Hessians when hess); returns the tile-wide total.
// Per-lane outputs: Ma (dof lanes), dof_force (friction + limit force on dof d).
template <unsigned LPE, class ES>
__device__ __forceinline__ float eval_cost(const Tile<LPE>& t, ES* S, SolveRegs<LPE, ES::NCAP>& r, bool hess, float& Ma, float& dof_force) {
  constexpr int RPL = SolveRegs<LPE, ES::NCAP>::RPL;
  const int lane = t.thread_rank();
  const int ncon = S->ncon;
  float cost = 0;
  Ma = 0; dof_force = 0;
  if (lane < NV) {
    const float ad = S->a[lane];
    Ma = mul_M(S, S->a, lane);
    cost = 0.5f * ad * Ma - ad * r.qfs;
    // frictionloss row (Huber)
    // selects instead of branches: the 12 dof lanes sit in different Huber zones and would serialise three paths
    const float x = ad - r.fr_aref, rf = r.fr_R * r.fr_fl, ax = fabsf(x);
    const bool quad = ax < rf;
    cost += quad ? 0.5f * r.fr_D * x * x : r.fr_fl * (ax - 0.5f * rf);
    dof_force = quad ? -r.fr_D * x : copysignf(r.fr_fl, -x);
    float hd = quad ? r.fr_D : 0.0f;
    {
      const float xl = r.lim_sgn * ad - r.lim_aref;       // lim_sgn = 0 (no active limit): lim_D = lim_aref = 0, xl = 0
      const float dl = xl < 0.0f ? r.lim_D : 0.0f;
      cost = fmaf(0.5f * dl * xl, xl, cost);
      dof_force = fmaf(-r.lim_sgn * dl, xl, dof_force);
      hd += dl;
    }
    if (hess) S->hdiag[lane] = hd;
  }
  const int nrow = ncon * 4;
#pragma unroll
  for (int s = 0; s < RPL; s++) {
    if (s * (int)LPE >= nrow) break;
    const int row = lane + s * LPE, c = row >> 2;
    float xv = 0;
    if (row < nrow) {
      // jar = J a - aref cancels to ~1e-4 of its terms under stiff contacts: accumulate in fp64
      // two fully unrolled 6-dof blocks (arm / cube), each with two accumulators: all loads in flight at once and
      // half the dependent DFMA chain of a rolled loop
      const int kind = S->ckind[c];
      double v0 = -(double)S->caref[c][row & 3], v1 = 0.0;
      if (kind & 1) {
#pragma unroll
        for (int d = 0; d < NL; d += 2) {
          v0 = fma((double)S->J[row][d], S->ad[d], v0);
          v1 = fma((double)S->J[row][d + 1], S->ad[d + 1], v1);
        }
      }
      if (kind & 2) {
#pragma unroll
        for (int d = NL; d < NV; d += 2) {
          v0 = fma((double)S->J[row][d], S->ad[d], v0);
          v1 = fma((double)S->J[row][d + 1], S->ad[d + 1], v1);
        }
      }
      xv = (float)(v0 + v1);
    }
    const int qb = lane & ~3;
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = t.shfl(xv, qb + k);
    if ((lane & 3) == 0 && row < nrow) {
      float force[4], Hc[10];
      int zone;
#pragma unroll
      for (int k = 0; k < 4; k++) r.jar[s][k] = x[k];
      cost += cone_eval(hess, x, S->cD[c], S->cDm[c], S->cmu[c], r.f0[s], r.f1[s], r.dim[s], force, zone, Hc);
      if (hess) {
#pragma unroll
        for (int k = 0; k < 4; k++) S->cfrc[c][k] = force[k];
        S->czone[c] = (unsigned char)zone;
        if (zone != 0) {
#pragma unroll
          for (int k = 0; k < 10; k++) S->cH[c][k] = Hc[k];
        }
      }
    }
  }
  return tsum(t, cost);
}

// T = H_c J_c for every active contact row (lane per (row, dof)): the Hessian J^T H J then costs 4 multiply-adds per
// (entry, contact) instead of a 4x4 quadratic form per (entry, contact)
template <unsigned LPE, class ES> __device__ __forceinline__ void cone_hess_rows(const Tile<LPE>& t, ES* S) {
  // T_c = H_c J_c (4 rows per contact).  One lane per dof column: it loads the four J entries of its column once and writes the four T
  // entries; groups of W lanes take contacts g, g + G, ...  (Item-per-lane indexing with the zone / kind look-ups inside cost 1.6 k cycles
  // of a Newton iteration; this is one shared-memory round trip per contact.)
  // uncoupled: a contact touches the arm block or the cube block, 6 dofs; coupled: all 12
  const int lane = t.thread_rank();
  const bool cpl = S->coupled != 0;
  const int W = cpl ? NV : NL;
  const int g = cpl ? (lane >= NV ? (lane >= 2 * NV ? 2 : 1) : 0) : (lane * 43) >> 8;      // lane / W for lane < 32
  const int G = cpl ? (int)LPE / NV : (int)LPE / NL;
  const int dl = lane - g * W;
  const int ncon = S->ncon;
  if (g < G) {
    for (int c = g; c < ncon; c += G) {
      const int zone = S->czone[c];
      if (zone == 0) continue;                     // the assembly skips the contact
      const int d = dl + ((!cpl && !(S->ckind[c] & 1)) ? NL : 0);
      const float* Hc = S->cH[c];
      const float j0 = S->J[c * 4 + 0][d], j1 = S->J[c * 4 + 1][d], j2 = S->J[c * 4 + 2][d], j3 = S->J[c * 4 + 3][d];
      float v[4];
      if (zone == 1) {                             // bottom zone: H_c = diag(D)
        v[0] = Hc[tri(0, 0)] * j0; v[1] = Hc[tri(1, 1)] * j1; v[2] = Hc[tri(2, 2)] * j2; v[3] = Hc[tri(3, 3)] * j3;
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          float a = 0;
          a = fmaf(Hc[tri(k > 0 ? k : 0, 0)], j0, a);
          a = fmaf(Hc[1 >= k ? tri(1, k) : tri(k, 1)], j1, a);
          a = fmaf(Hc[2 >= k ? tri(2, k) : tri(k, 2)], j2, a);
          a = fmaf(Hc[3 >= k ? tri(3, k) : tri(k, 3)], j3, a);
          v[k] = a;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; k++) S->T[c * 4 + k][d] = v[k];
    }
  }
  t.sync();
}

// Newton direction -H^-1 g for a Hessian that couples the arm and the cube (a contact between them): dense 12x12.
// Only the heavy solve kernel is compiled with it (solve<DENSE = true>): envs with an arm-cube contact are routed there, so the
// light kernel's hot loop carries neither this code nor its registers.
template <unsigned LPE, class ES> __device__ __forceinline__ float dense_newton_dir(const Tile<LPE>& t, ES* S, float g) {
  const int lane = t.thread_rank();
  float pd;
#ifdef SO100_SOLVE_CLOCK
  long long dc_[5]; dc_[0] = clock64();
#endif
    // the dense 12x12 Hessian (packed lower triangle) is in S->H: assembled by the caller
    // Block elimination H = [[A, B^T], [B, C]] (A arm 6x6, C cube 6x6, B their coupling): factor A in registers,
    // W = L_A^-1 B^T with one column per lane, the Schur complement S = C - W^T W with one entry per lane, factor S in
    // registers.  Two 6-column register factorisations and two barriers instead of a 12-column chain of dependent
    // shuffles (measured: 4.5 k cycles per coupled Newton iteration for the 12-column version).
    float* const Wm = &S->T[0][0];          // scratch in the H_c J_c rows (dead until the next iteration): W[k][r] at 6 k + r
    float* const y1 = Wm + 36;              // L_A^-1 (-g_arm)
    float* const Sm = Wm + 48;              // packed Schur complement
    float* const r2 = Wm + 72;              // its right-hand side
#ifdef SO100_SOLVE_CLOCK
    dc_[1] = clock64();
#endif
    float LA[21];
    chol6_factor<false>(S->H, LA);          // entries 0..20 of the packed 12x12 are the arm block
    {
      float w[NL];
#pragma unroll
      for (int c = 0; c < NL; c++) w[c] = lane < NL ? S->H[tri(NL + (lane < NL ? lane : 0), c)] : -S->vec[c];
      chol6_fwd(LA, w);                     // lanes 0..5: column `lane` of W; every other lane: y1
      if (lane <= NL) {
#pragma unroll
        for (int k = 0; k < NL; k++) {
          if (lane < NL) Wm[NL * k + lane] = w[k];
          else y1[k] = w[k];
        }
      }
    }
    t.sync();
#ifdef SO100_SOLVE_CLOCK
    dc_[2] = clock64();
#endif
    if (lane < 21) {
      const int r = tri_row6(lane), c = lane - tri(r, 0);
      float sacc = S->H[tri(NL + r, NL + c)];
#pragma unroll
      for (int k = 0; k < NL; k++) sacc = fmaf(-Wm[NL * k + r], Wm[NL * k + c], sacc);
      Sm[lane] = sacc;
    } else if (lane < 21 + NL) {
      const int r = lane - 21;
      float racc = -S->vec[NL + r];
#pragma unroll
      for (int k = 0; k < NL; k++) racc = fmaf(-Wm[NL * k + r], y1[k], racc);
      r2[r] = racc;
    }
    t.sync();
    float x2[NL], x1[NL];
    chol6_solve<false>(Sm, r2, 1.0f, x2);
#pragma unroll
    for (int k = 0; k < NL; k++) {
      float acc = y1[k];
#pragma unroll
      for (int r = 0; r < NL; r++) acc = fmaf(-Wm[NL * k + r], x2[r], acc);
      x1[k] = acc;
    }
    chol6_bwd(LA, x1);
    pd = 0.0f;
#pragma unroll
    for (int k = 0; k < NL; k++) {
      pd = (lane == k) ? x1[k] : pd;
      pd = (lane == NL + k) ? x2[k] : pd;
    }
    t.sync();                                // every lane has read the gradient in S->vec
    if (lane < NV) S->vec[lane] = pd;
    t.sync();
#ifdef SO100_SOLVE_CLOCK
  dc_[3] = clock64();
  if (lane == 0) for (int k_ = 0; k_ < 3; k_++) S->clk2[k_] += (int)(dc_[k_ + 1] - dc_[k_]);
#endif
  return pd;
}

#ifdef SO100_SOLVE_CLOCK
// development build: SM-cycle split of the last solve (evaluation, gradient, Hessian + factorisation, line search)
#define SOLVE_CLK_DECL long long clk_[4] = {0, 0, 0, 0}; long long clk_t_ = clock64()
#define SOLVE_CLK(k) do { const long long now_ = clock64(); clk_[k] += now_ - clk_t_; clk_t_ = now_; } while (0)
#define SOLVE_CLK_STORE(S) do { if (t.thread_rank() == 0) for (int k_ = 0; k_ < 4; k_++) (S)->clk[k_] = (int)clk_[k_]; } while (0)
#else
#define SOLVE_CLK_DECL do { } while (0)
#define SOLVE_CLK(k) do { } while (0)
#define SOLVE_CLK_STORE(S) do { } while (0)
#endif

// Solves for qacc (left in S->a / S->ad, contact forces in S->cfrc).  `diag` (nullable): the env's uint32 counters.
// Returns the number of Newton iterations.
// max_it < NEWTON_MAXIT: iteration budget of the regular light kernel; a solve that uses it up without converging returns -1,
// counts nothing in `diag` and leaves an unfinished iterate behind (the env goes to the slow lane, which solves it again in full).
#ifdef SO100_SOLVE_TRACE
// development build: per Newton iteration of env < 64 (so100_forward on a small batch): cost, scaled gradient, its tolerance, phi'(0),
// step length, line-search evaluations, final phi', relative predicted decrease
__device__ float g_solve_trace[64][104][8];
#endif
template <bool DENSE, unsigned LPE, class ES> __device__ int solve(const Tile<LPE>& t, ES* S, const DevTables& T, uint32_t* diag, bool active = true,
                                                                   int max_it = NEWTON_MAXIT, int dbg_env = -1) {
  using Regs = SolveRegs<LPE, ES::NCAP>;
  const int lane = t.thread_rank();
  const int ncon = S->ncon;
  const bool coupled = DENSE && S->coupled != 0;
  Regs r;
  // ---- dof rows
  r.qfs = 0; r.fr_aref = 0; r.fr_R = 1; r.fr_D = 0; r.fr_fl = 0; r.lim_sgn = 0; r.lim_D = 0; r.lim_aref = 0;
  float qas_d = 0, warm_d = 0;
  if (lane < NV) {
    r.qfs = S->d.qfs[lane];
    qas_d = S->d.qas[lane];
    warm_d = S->st[S_WARM + lane];
    r.fr_R = c_m.fr_R[lane]; r.fr_D = c_m.fr_D[lane]; r.fr_fl = c_m.fr_floss[lane];
    const float qd = S->st[S_QVEL + lane];
    r.fr_aref = -c_m.fr_B * qd;
    if (lane < NL) {
      const float q = S->st[S_QPOS + lane];
      float dist = 0;
      if (q < c_m.lim_lo[lane]) { r.lim_sgn = 1.0f; dist = q - c_m.lim_lo[lane]; }
      else if (q > c_m.lim_hi[lane]) { r.lim_sgn = -1.0f; dist = c_m.lim_hi[lane] - q; }
      if (r.lim_sgn != 0.0f) {
        float imp = impedance(c_m.lim_solimp, dist);
        float R = fmaxf((1.0f - imp) / imp * c_m.lim_invw[lane], 1e-15f);
        r.lim_D = 1.0f / R;
        r.lim_aref = -c_m.lim_B * (r.lim_sgn * qd) - c_m.lim_K * imp * dist;
      }
    }
  }
#pragma unroll
  for (int s = 0; s < Regs::RPL; s++) {
    const int c = (lane + s * LPE) >> 2;
    r.dim[s] = 3; r.f0[s] = 1; r.f1[s] = 1;
    if ((lane & 3) == 0 && c < ncon) {
      const DevPair& P = T.pair[__float_as_int(S->con[c][7])];
      r.dim[s] = P.dim; r.f0[s] = P.f0; r.f1[s] = P.f1;
    }
  }
  // ---- one loop body for: the start-point selection (stage 0: cost of the unconstrained acceleration, stage 1:
  //      cost + forces of the warm start), the Newton iterations (stage 2) and the final force refresh
  auto set_a = [&](float v) {
    if (lane < NV) { S->a[lane] = v; S->ad[lane] = (double)v; }
    t.sync();
  };
  set_a(qas_d);
  // ---- Hessian entries of this lane (packed lower triangle: 78 dense, 42 block-diagonal), fixed for the whole solve: indices, the
  // blocks a contact has to touch to contribute, and the mass-matrix term.  (Recomputing them in every Newton iteration, with the
  // contact loop inside the entry loop, made the assembly 3.0 k of the dense iteration's 14 k cycles.)
  constexpr int HR = LPE == 16 ? 3 : (DENSE ? 3 : 2);
  int hidx[HR];
#pragma unroll
  for (int k = 0; k < HR; k++) {
    const int e = lane + k * (int)LPE;
    int gi, gj, need, valid;
    if (coupled) {
      valid = e < 78;
      untri(valid ? e : 0, gi, gj);
      need = (gi < NL ? 1 : 2) | (gj < NL ? 1 : 2);
    } else {
      valid = e < 42;
      const int ee = valid ? e : 0, blk = ee >= 21 ? 1 : 0, rr = ee - 21 * blk;
      const int i = tri_row6(rr), j = rr - tri(i, 0);
      gi = i + NL * blk; gj = j + NL * blk;
      need = 1 << blk;
    }
    hidx[k] = gi | (gj << 4) | (need << 8) | (valid << 10) | ((gi == gj ? 1 : 0) << 11);
  }
  // S->H = M + diag(friction / limit curvature) + sum_c J_c^T (H_c J_c); entry e of lane `lane` is e = lane + k LPE
  auto assemble_hessian = [&]() {
    float h[HR];
#pragma unroll
    for (int k = 0; k < HR; k++) {
      const int gi = hidx[k] & 15, gj = (hidx[k] >> 4) & 15;
      float b = 0.0f;                                 // mass matrix: arm block dense, cube block diagonal, no coupling
      if (gi < NL) b = S->d.Mfull[gi][gj];
      else if ((hidx[k] >> 11) & 1) b = gi < 9 ? c_m.cube_mass : c_m.cube_I[gi - 9];
      h[k] = b + (((hidx[k] >> 11) & 1) ? S->hdiag[gi] : 0.0f);
    }
    for (int c = 0; c < ncon; c++) {
      if (S->czone[c] == 0) continue;
      const int kind = S->ckind[c];
      const float* Jc = &S->J[c * 4][0];
      const float* Tc = &S->T[c * 4][0];
#pragma unroll
      for (int k = 0; k < HR; k++) {
        const int need = (hidx[k] >> 8) & 3;
        if (((hidx[k] >> 10) & 1) && (kind & need) == need) {
          const int gi = hidx[k] & 15, gj = (hidx[k] >> 4) & 15;
          const float a01 = fmaf(Jc[gi], Tc[gj], Jc[JS + gi] * Tc[JS + gj]);
          const float a23 = fmaf(Jc[2 * JS + gi], Tc[2 * JS + gj], Jc[3 * JS + gi] * Tc[3 * JS + gj]);
          h[k] += a01 + a23;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < HR; k++) if ((hidx[k] >> 10) & 1) S->H[lane + k * (int)LPE] = h[k];
    t.sync();
  };
  int stage = 0, it = 0;
  bool done = !active, last = false, converged = false;
  float Ma = 0, dof_force = 0, cost = 0, cost_qas = 0, cost_prev = 3.0e38f;
  int stall = 0;
  float pred_rel = 1.0f;      // predicted decrease of the last Newton step relative to the cost
  // one Newton iteration from the iterate / forces of the last evaluation; sets `done` when the solve is over
  SOLVE_CLK_DECL;
  auto newton_step = [&]() {
    SOLVE_CLK(0);
    // ---- gradient: M a - qfrc_smooth - J^T f
    float g = 0, jtf = 0;
    if (lane < NV) {
      jtf = dof_force;
      const int bit = lane < NL ? 1 : 2;
      for (int c = 0; c < ncon; c++) {
        if (S->czone[c] == 0 || !(S->ckind[c] & bit)) continue;
#pragma unroll
        for (int k = 0; k < 4; k++) jtf = fmaf(S->J[c * 4 + k][lane], S->cfrc[c][k], jtf);
      }
      g = Ma - r.qfs - jtf;
      S->vec[lane] = g;
    }
    // float32 stopping rule: MuJoCo's 1e-8 is below the round-off of the cancelling terms, so the
    // tolerance is 2e-6 relative to their magnitude (|qfrc_smooth| + |J^T f|), scaled like MuJoCo's.
    float gg = g * g, ss = r.qfs * r.qfs + jtf * jtf;
    tsum2(t, gg, ss);
#ifdef SO100_SOLVE_TRACE
    if (lane == 0 && dbg_env >= 0 && dbg_env < 64 && it < 104) {
      float* q_ = g_solve_trace[dbg_env][it];
      q_[0] = cost; q_[1] = sqrtf(gg) * c_m.inv_scale; q_[2] = SO100_GTOL * (1.0f + sqrtf(ss));
    }
#endif
    if (sqrtf(gg) * c_m.inv_scale < SO100_GTOL * (1.0f + sqrtf(ss))) { converged = true; done = true; return; }
    float pd;
    SOLVE_CLK(1);
    cone_hess_rows(t, S);
    assemble_hessian();
    if (!coupled) {
      // ---- block-diagonal Hessian: entries 0..20 arm block, 21..41 cube block
      // both blocks factored + solved in registers: lanes 0..5 (and 12..) the arm block, lanes 6..11 the cube block, so
      // that dof lane d already holds its own component of the direction
      const int hb = (lane >= NL && lane < NV) ? 1 : 0;
      float x[NL];
      chol6_solve<false>(&S->H[21 * hb], &S->vec[NL * hb], -1.0f, x);
      const int kx = lane - NL * hb;
      pd = 0.0f;
#pragma unroll
      for (int k = 0; k < NL; k++) pd = (kx == k) ? x[k] : pd;
      if (lane >= NV) pd = 0.0f;
      t.sync();                                  // every lane has read the gradient in S->vec
      if (lane < NV) S->vec[lane] = pd;
      t.sync();
    } else {
      if constexpr (DENSE) pd = dense_newton_dir(t, S, g);
      else pd = 0.0f;
    }
    SOLVE_CLK(2);
    // ---- line-search set-up
    float Mp = 0;
    if (lane < NV) Mp = mul_M(S, S->vec, lane);
    float pMp = pd * Mp, pg = pd * (Ma - r.qfs), gp = pd * g;    // gp: phi'(0) = g . p
    tsum3(t, pMp, pg, gp);       // one pass of shuffle rounds for the three sums
    const int nrow = ncon * 4;
#pragma unroll
    for (int s = 0; s < Regs::RPL; s++) {
      if (s * (int)LPE >= nrow) break;
      const int row = lane + s * LPE, c = row >> 2;
      float v = 0;
      if (row < nrow) {
        const int kind = S->ckind[c];
        float v1 = 0;
        if (kind & 1) {
#pragma unroll
          for (int d = 0; d < NL; d += 2) { v = fmaf(S->J[row][d], S->vec[d], v); v1 = fmaf(S->J[row][d + 1], S->vec[d + 1], v1); }
        }
        if (kind & 2) {
#pragma unroll
          for (int d = NL; d < NV; d += 2) { v = fmaf(S->J[row][d], S->vec[d], v); v1 = fmaf(S->J[row][d + 1], S->vec[d + 1], v1); }
        }
        v += v1;
      }
      const int qb = lane & ~3;
      float jv[4];
#pragma unroll
      for (int k = 0; k < 4; k++) jv[k] = t.shfl(v, qb + k);
      if ((lane & 3) == 0 && row < nrow) cone_ls_prepare(r.jar[s], jv, S->cD[c], S->cDm[c], S->cmu[c], r.f0[s], r.f1[s], r.dim[s], r.ls[s]);
    }
    const float a0 = S->a[lane < NV ? lane : 0];
    const float xf0 = a0 - r.fr_aref, xl0 = r.lim_sgn * a0 - r.lim_aref;
    // ---- exact line search: safeguarded 1-D Newton on phi'.  phi'(0) = g . p is known from the gradient, and with the exact
    // Hessian phi''(0) = p . H p = -g . p, so the first Newton iterate from alpha = 0 is alpha = 1: the search starts there
    // instead of spending an evaluation on alpha = 0.
    const float d10 = fabsf(gp);
    const bool descent = gp < 0;                 // no descent left at float32 resolution otherwise
    float alpha = 1.0f, d1 = 0, d2 = 1, lo = 0, hi = -1;
    [[maybe_unused]] int ls_n_ = 0;
#pragma unroll 1
    for (int ls = 0; descent && ls < LS_MAXIT; ls++) {
      ls_n_ = ls + 1;
      if (ls > 0) {
        float na = alpha - __fdividef(d1, d2);     // approximate division: a safeguarded iterate, not a result
        if (hi >= 0 && (na <= lo || na >= hi)) na = 0.5f * (lo + hi);
        alpha = na;
      }
      float e1 = 0, e2 = 0;
      if (lane < NV) {
        const float x = fmaf(alpha, pd, xf0), rf = r.fr_R * r.fr_fl;
        const bool quad = fabsf(x) < rf;
        e1 = (quad ? r.fr_D * x : copysignf(r.fr_fl, x)) * pd;
        e2 = quad ? r.fr_D * pd * pd : 0.0f;
        const float v = r.lim_sgn * pd, xl = fmaf(alpha, v, xl0);
        const float dl = xl < 0.0f ? r.lim_D : 0.0f;      // lim_D = 0 without an active limit
        e1 = fmaf(dl * xl, v, e1);
        e2 = fmaf(dl * v, v, e2);
      }
      if ((lane & 3) == 0) {
#pragma unroll
        for (int s = 0; s < Regs::RPL; s++) {
          const int c = (lane + s * LPE) >> 2;
          if (c < ncon) cone_ls(r.ls[s], alpha, e1, e2);
        }
      }
      tsum2(t, e1, e2);
#ifdef SO100_SOLVE_CLOCK
      if (lane == 0) S->clk2[3] += 1;
#endif
      d1 = e1 + pg + alpha * pMp;
      d2 = e2 + pMp;
      if (fabsf(d1) <= SO100_LS_TOL * d10) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      // bracket narrower than any step length matters: phi' has no resolvable root (float32 noise floor, or a kink it jumps across);
      // 1e-4 instead of float32 resolution (2e-7) is ten bisections fewer on such iterations, results unchanged
      if (hi >= 0 && hi - lo <= 1e-4f * hi) break;
    }
#ifdef SO100_SOLVE_TRACE
    if (lane == 0 && dbg_env >= 0 && dbg_env < 64 && it < 104) {
      float* q_ = g_solve_trace[dbg_env][it];
      q_[3] = gp; q_[4] = alpha; q_[5] = (float)ls_n_; q_[6] = d1; q_[7] = 0.5f * alpha * d10 / (1.0f + fabsf(cost));
    }
#endif
    if (!descent) { converged = true; done = true; return; }
    if (lane < NV) {
      const double na = fma((double)alpha, (double)pd, S->ad[lane]);
      S->ad[lane] = na; S->a[lane] = (float)na;
    }
    t.sync();
    it++;
    SOLVE_CLK(3);
    // predicted decrease 1/2 alpha |phi'(0)| below float32 resolution of the cost: stop after refreshing the forces
    // (MuJoCo's "improvement < tolerance" test, made relative because the arithmetic is float32)
    pred_rel = 0.5f * alpha * d10 / (1.0f + fabsf(cost));
    if (pred_rel < SO100_ITOL) last = true;
  };
  // Every trip = one evaluation, then (stage 2) one Newton iteration.  With two envs per warp (LPE = 16) both tiles
  // vote at the top of every trip, so they run the trip in lock step and re-converge there; without the vote, tiles
  // that took different branches once (e.g. the re-evaluation at the unconstrained start point) would never re-converge
  // and the warp would execute both solves serially (ncu: 14.8 active threads per instruction).
#pragma unroll 1
  for (;;) {
    if (!warp_any<LPE>(t, !done)) break;
    if (!done) {
      cost = eval_cost(t, S, r, stage != 0, Ma, dof_force);
      t.sync();
      if (stage == 0) { cost_qas = cost; stage = 1; set_a(warm_d); }
      else if (stage == 1 && cost_qas < cost) { stage = 2; set_a(qas_d); }
      else {
        stage = 2;
        // MuJoCo's "improvement < tolerance" test on the ACTUAL decrease: two consecutive iterations that did not lower the
        // cost by one part in 10^7 (float32 resolution) mean the iterate sits at the float32 optimum; without this the rare
        // solve whose predicted decrease stays above the tolerance (round-off in the cancelling terms) runs to the iteration cap
        // cost_prev is the LOWEST cost seen: at the float32 optimum of a stiff problem the iterate can alternate between two points
        // whose costs differ by a few 1e-6 relative (every second iteration "improves" on the one before, each with a ~25-evaluation
        // line search that ends at bracket resolution), which a test against the previous iteration alone never catches
        // An iteration only counts as stalled when its step did not promise a visible decrease either (predicted < 1e-7 of the cost):
        // a cost of a few hundred resolves 3e-5, and a solve that still gains 5e-5 per iteration on stiff contacts is making progress
        // its float32 cost cannot show (cube in the bin: force error 6e-3 instead of 2e-4 when such iterations were counted).
        if (it > 0 && !(cost < cost_prev - 1e-7f * fabsf(cost_prev))) { if (pred_rel < 1e-7f && ++stall >= 2) last = true; }
        else { stall = 0; cost_prev = cost; }
        if (last || it >= max_it) done = true;
        else newton_step();
      }
    }
  }
  SOLVE_CLK(0);
  SOLVE_CLK_STORE(S);
  if (active && max_it < NEWTON_MAXIT && !(converged || last)) return -1;      // over budget (tile-uniform)
  if (lane == 0 && diag) {
    diag[1] += (converged || last) ? 0u : 1u;
    diag[5] += (uint32_t)it;
    diag[6] += 1u;
    diag[7] += (uint32_t)ncon;
  }
  return it;
}

// ------------------------------------------------------------------ semi-implicit Euler
template <unsigned LPE, class ES> __device__ void integrate(const Tile<LPE>& t, ES* S) {
  const int lane = t.thread_rank();
  const float h = c_m.timestep;
  if (lane < NV) {
    const float a = S->a[lane];
    S->st[S_QVEL + lane] = fmaf(h, a, S->st[S_QVEL + lane]);
    S->st[S_WARM + lane] = a;
  }
  t.sync();
  if (lane < 9) {
    S->st[S_QPOS + lane] = fmaf(h, S->st[S_QVEL + lane], S->st[S_QPOS + lane]);
  } else if (lane == 9) {
    V3 w = ld3(&S->st[S_QVEL + 9]);
    Q4 q = {S->st[S_QPOS + 9], S->st[S_QPOS + 10], S->st[S_QPOS + 11], S->st[S_QPOS + 12]};
    q = qnormalize(q);
    const float wn = sqrtf(dot(w, w)), ang = wn * h;
    if (ang > 0) {
      float sn, cs;
      sincosf(0.5f * ang, &sn, &cs);
      const float k = sn / wn;
      Q4 dq = {cs, w.x * k, w.y * k, w.z * k};
      q = qnormalize(qmul(q, dq));
    }
    S->st[S_QPOS + 9] = q.w; S->st[S_QPOS + 10] = q.x; S->st[S_QPOS + 11] = q.y; S->st[S_QPOS + 12] = q.z;
  }
  t.sync();
}

}  // namespace so100
