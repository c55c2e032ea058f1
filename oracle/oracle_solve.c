/* TEST INFRASTRUCTURE (see oracle_internal.h).  Primal Newton solver with elliptic friction
 * cones (MuJoCo engine_solver.c mj_solNewton + engine_core_constraint.c
 * mj_constraintUpdate restated): minimise over a = qacc
 *     1/2 (a - a_s)^T M (a - a_s) + sum_rows s_i(J a - aref)
 * The optimum is unique; this oracle iterates to a much tighter tolerance than MuJoCo's 1e-8 so
 * its answer is the exact optimum to ~1e-12. */
#include "oracle_internal.h"
#include <stdio.h>
#include <stdlib.h>

typedef struct {
  double cost;
  double Hc[MAXCON][16];   /* cone Hessian blocks (dim x dim, row-major stride 4), middle zone only */
  int zone[MAXCON];        /* 0 top (inactive), 1 bottom (quadratic), 2 middle (cone) */
  int active[MAXEFC];      /* quadratic rows contributing D to the Hessian */
} ctx;

/* evaluates row costs / forces at jar; optionally cone Hessians */
static double update_rows(const oenv* e, const double* jar, double* force, ctx* c, int want_hess) {
  double cost = 0;
  for (int i = 0; i < e->nefc; i++) {
    if (e->etype[i] != EFC_CONTACT_CONT) c->active[i] = 0;
    double x = jar[i];
    switch (e->etype[i]) {
      case EFC_FRICTION: {
        double f = e->efloss[i], R = e->eR[i], D = e->eD[i];
        if (x <= -R * f) { cost += f * (-0.5 * R * f - x); force[i] = f; }
        else if (x >= R * f) { cost += f * (-0.5 * R * f + x); force[i] = -f; }
        else { cost += 0.5 * D * x * x; force[i] = -D * x; c->active[i] = 1; }
      } break;
      case EFC_LIMIT:
        if (x < 0) { cost += 0.5 * e->eD[i] * x * x; force[i] = -e->eD[i] * x; c->active[i] = 1; }
        else force[i] = 0;
        break;
      case EFC_CONTACT: {
        const ocontact* con = &e->con[e->eid[i]];
        int dim = con->dim, ci = e->eid[i];
        double mu = con->mu, fr[3] = {con->friction[0], con->friction[0], con->friction[1]};
        double U[4], N, T = 0;
        U[0] = x * mu;
        for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * fr[j - 1]; T += U[j] * U[j]; }
        T = sqrt(T); N = U[0];
        if (N >= mu * T || (T <= 0 && N >= 0)) {            /* top zone: separated */
          c->zone[ci] = 0;
          for (int j = 0; j < dim; j++) { force[i + j] = 0; c->active[i + j] = 0; }
        } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {  /* bottom zone: inside the polar cone */
          c->zone[ci] = 1;
          for (int j = 0; j < dim; j++) {
            cost += 0.5 * e->eD[i + j] * jar[i + j] * jar[i + j];
            force[i + j] = -e->eD[i + j] * jar[i + j];
            c->active[i + j] = 1;
          }
        } else {                                            /* middle zone: on the cone surface */
          c->zone[ci] = 2;
          for (int j = 0; j < dim; j++) c->active[i + j] = 0;
          double Dm = e->eD[i] / fmax(MJMINVAL, mu * mu * (1 + mu * mu));
          double NmT = N - mu * T;
          cost += 0.5 * Dm * NmT * NmT;
          force[i] = -Dm * NmT * mu;
          for (int j = 1; j < dim; j++) force[i + j] = -force[i] / T * U[j] * fr[j - 1];
          if (want_hess) {
            /* H = Dm [ g g^T - NmT*mu ( diag(fr^2)/T - (fr U)(fr U)^T / T^3 ) ],
               g = d(N - mu T)/dx = (mu, -mu fr_j U_j / T) */
            double g[4], *H = c->Hc[ci];
            g[0] = mu;
            for (int j = 1; j < dim; j++) g[j] = -mu * fr[j - 1] * U[j] / T;
            for (int a = 0; a < dim; a++)
              for (int b = 0; b < dim; b++) {
                double h = g[a] * g[b];
                if (a > 0 && b > 0) {
                  double t2 = -(fr[a - 1] * U[a]) * (fr[b - 1] * U[b]) / (T * T * T);
                  if (a == b) t2 += fr[a - 1] * fr[a - 1] / T;
                  h -= NmT * mu * t2;
                }
                H[a * 4 + b] = Dm * h;
              }
          }
        }
      } break;
      default: break; /* EFC_CONTACT_CONT handled with its first row */
    }
  }
  return cost;
}

/* 1-D derivatives of the constraint cost along jar + alpha*jv */
static void ls_eval(const oenv* e, const double* jar, const double* jv, double alpha, double* d1, double* d2) {
  double g = 0, h = 0;
  for (int i = 0; i < e->nefc; i++) {
    double x = jar[i] + alpha * jv[i], v = jv[i];
    switch (e->etype[i]) {
      case EFC_FRICTION: {
        double f = e->efloss[i], R = e->eR[i], D = e->eD[i];
        if (x <= -R * f) g += -f * v;
        else if (x >= R * f) g += f * v;
        else { g += D * x * v; h += D * v * v; }
      } break;
      case EFC_LIMIT:
        if (x < 0) { g += e->eD[i] * x * v; h += e->eD[i] * v * v; }
        break;
      case EFC_CONTACT: {
        const ocontact* con = &e->con[e->eid[i]];
        int dim = con->dim;
        double mu = con->mu, fr[3] = {con->friction[0], con->friction[0], con->friction[1]};
        double N = x * mu, Np = v * mu, TT = 0, UV = 0, VV = 0;
        for (int j = 1; j < dim; j++) {
          double U = (jar[i + j] + alpha * jv[i + j]) * fr[j - 1], V = jv[i + j] * fr[j - 1];
          TT += U * U; UV += U * V; VV += V * V;
        }
        double T = sqrt(TT);
        if (N >= mu * T || (T <= 0 && N >= 0)) {
        } else if (mu * N + T <= 0 || (T <= 0 && N < 0)) {
          for (int j = 0; j < dim; j++) {
            double xj = jar[i + j] + alpha * jv[i + j];
            g += e->eD[i + j] * xj * jv[i + j];
            h += e->eD[i + j] * jv[i + j] * jv[i + j];
          }
        } else {
          double Dm = e->eD[i] / fmax(MJMINVAL, mu * mu * (1 + mu * mu));
          double Tp = UV / T, Tpp = (VV - Tp * Tp) / T, NmT = N - mu * T, dp = Np - mu * Tp;
          g += Dm * NmT * dp;
          h += Dm * (dp * dp - NmT * mu * Tpp);
        }
      } break;
      default: break;
    }
  }
  *d1 = g; *d2 = h;
}

static double total_cost(const oenv* e, int nv, const double* a, const double* jar, double* force, ctx* c, int hess) {
  /* Gauss term 1/2 (a-a_s)^T M (a-a_s) */
  double gauss = 0;
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int j = 0; j < nv; j++) s += e->M[i * NVMAX + j] * (a[j] - e->qacc_smooth[j]);
    gauss += 0.5 * (a[i] - e->qacc_smooth[i]) * s;
  }
  return gauss + update_rows(e, jar, force, c, hess);
}

static void compute_jar(const oenv* e, int nv, const double* a, double* jar) {
  for (int i = 0; i < e->nefc; i++) {
    double s = -e->earef[i];
    for (int j = 0; j < nv; j++) s += e->J[i][j] * a[j];
    jar[i] = s;
  }
}

/* Solver mode.  0 (default, the CHECKER): iterate to the exact optimum (scaled gradient 1e-11, line search to 1e-14).
 * 1 (the CPU BASELINE of bench.py): MuJoCo's own settings for this scene -- the XML leaves them at their defaults
 * (so_arm100.xml:4 sets only cone / impratio): tolerance 1e-8 on the scaled gradient or the scaled improvement, at most 100
 * Newton iterations, line search of at most 50 iterations stopping at ls_tolerance 0.01 (taken relative to the initial
 * slope, which stops no later than MuJoCo's absolute rule). */
static int g_solver_mode = 0;
void so100o_set_solver_mode(int mode) { g_solver_mode = mode; }

void o_solve(const so100_model* m, oenv* e) {
  int nv = m->nv, nefc = e->nefc;
  const int fast = g_solver_mode == 1;
  const double gtol = fast ? 1e-8 : 1e-11, lstol = fast ? 1e-2 : 1e-14;
  const int maxit = fast ? 100 : 200, lsmax = fast ? 50 : 100;
  double prev_cost = 0;
  static _Thread_local ctx c;
  double a[NVMAX], jar[MAXEFC], force[MAXEFC], jv[MAXEFC];
  /* warm start: previous qacc if it is cheaper than the unconstrained acceleration */
  compute_jar(e, nv, e->warm, jar);
  double cw = total_cost(e, nv, e->warm, jar, force, &c, 0);
  compute_jar(e, nv, e->qacc_smooth, jar);
  double cs = total_cost(e, nv, e->qacc_smooth, jar, force, &c, 0);
  memcpy(a, cw < cs ? e->warm : e->qacc_smooth, sizeof(a));
  double scale = 1.0 / (m->meaninertia * (nv > 1 ? nv : 1));
  e->solver_iter = 0;
  for (int it = 0; it < maxit; it++) {
    compute_jar(e, nv, a, jar);
    double cost = total_cost(e, nv, a, jar, force, &c, 1);
    /* MuJoCo: stop when the last iteration improved the cost by less than the tolerance (scaled) */
    if (fast && it > 0 && (prev_cost - cost) * scale < gtol) { e->solver_iter = it; break; }
    prev_cost = cost;
    /* gradient = M a - qfrc_smooth - J^T force */
    double grad[NVMAX], H[NVMAX * NVMAX], L[NVMAX * NVMAX], p[NVMAX];
    for (int i = 0; i < nv; i++) {
      double s = -e->qfrc_smooth[i];
      for (int j = 0; j < nv; j++) s += e->M[i * NVMAX + j] * a[j];
      for (int r = 0; r < nefc; r++) s -= e->J[r][i] * force[r];
      grad[i] = s;
    }
    double gn = 0;
    for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
    gn = sqrt(gn);
    e->solver_grad = gn * scale;
    e->solver_iter = it;
    /* 1e-11 is ~100x above the fp64 round-off floor of this gradient and 1000x tighter than MuJoCo's 1e-8 */
    if (gn * scale < gtol) break;
    /* Hessian */
    memcpy(H, e->M, sizeof(H));
    for (int r = 0; r < nefc; r++) {
      if (c.active[r]) {
        for (int i = 0; i < nv; i++)
          for (int j = 0; j < nv; j++) H[i * NVMAX + j] += e->eD[r] * e->J[r][i] * e->J[r][j];
      }
      if (e->etype[r] == EFC_CONTACT && c.zone[e->eid[r]] == 2) {
        int dim = e->con[e->eid[r]].dim;
        const double* Hc = c.Hc[e->eid[r]];
        for (int ca = 0; ca < dim; ca++)
          for (int cb = 0; cb < dim; cb++) {
            double h = Hc[ca * 4 + cb];
            for (int i = 0; i < nv; i++)
              for (int j = 0; j < nv; j++) H[i * NVMAX + j] += h * e->J[r + ca][i] * e->J[r + cb][j];
          }
      }
    }
    o_chol(L, H, nv);
    for (int i = 0; i < nv; i++) p[i] = -grad[i];
    o_chol_solve(L, p, nv);
    /* exact line search on phi(alpha) = cost(a + alpha p): safeguarded 1-D Newton on phi' */
    double pMp = 0, pg = 0;
    for (int i = 0; i < nv; i++) {
      double s = 0;
      for (int j = 0; j < nv; j++) s += e->M[i * NVMAX + j] * p[j];
      pMp += p[i] * s;
      double ma = -e->qfrc_smooth[i];
      for (int j = 0; j < nv; j++) ma += e->M[i * NVMAX + j] * a[j];
      pg += p[i] * ma;    /* derivative of the Gauss term at alpha = 0 */
    }
    for (int r = 0; r < nefc; r++) {
      double s = 0;
      for (int j = 0; j < nv; j++) s += e->J[r][j] * p[j];
      jv[r] = s;
    }
    double lo = 0, hi = -1, alpha = 0, d1, d2, d10;
    ls_eval(e, jar, jv, 0, &d1, &d2);
    d1 += pg; d2 += pMp;
    d10 = fabs(d1);
    if (getenv("SO100O_DEBUG")) {
      double gp = 0, pHp = 0;
      for (int i = 0; i < nv; i++) { gp += grad[i] * p[i]; for (int j = 0; j < nv; j++) pHp += p[i] * H[i * NVMAX + j] * p[j]; }
      fprintf(stderr, "   gp %.6e d1(0) %.6e   pHp %.6e d2(0) %.6e\n", gp, d1, pHp, d2);
    }
    if (d1 >= 0) break;                       /* not a descent direction: converged to round-off */
    for (int ls = 0; ls < lsmax; ls++) {
      double step = -d1 / d2, na = alpha + step;
      if (hi >= 0 && (na <= lo || na >= hi)) na = 0.5 * (lo + hi);
      else if (hi < 0 && na <= lo) na = 2 * alpha + 1e-6;
      alpha = na;
      ls_eval(e, jar, jv, alpha, &d1, &d2);
      d1 += pg + alpha * pMp; d2 += pMp;
      if (fabs(d1) < lstol * d10 + 1e-300) break;
      if (d1 < 0) lo = alpha; else hi = alpha;
      if (hi >= 0 && hi - lo < 1e-16 * (1 + hi)) break;
    }
    if (getenv("SO100O_DEBUG")) fprintf(stderr, "it %d cost %.12g grad %.3e alpha %.6g d1 %.3e d10 %.3e\n", it, cost, gn, alpha, d1, d10);
    for (int i = 0; i < nv; i++) a[i] += alpha * p[i];
  }
  compute_jar(e, nv, a, jar);
  total_cost(e, nv, a, jar, force, &c, 0);
  memcpy(e->qacc, a, sizeof(a));
  memcpy(e->eforce, force, nefc * sizeof(double));
  memcpy(e->ejar, jar, nefc * sizeof(double));
  for (int i = 0; i < nv; i++) {
    double s = 0;
    for (int r = 0; r < nefc; r++) s += e->J[r][i] * force[r];
    e->qfrc_constraint[i] = s;
  }
  for (int cidx = 0; cidx < e->ncon; cidx++)
    for (int j = 0; j < e->con[cidx].dim; j++) e->con[cidx].force[j] = force[e->con[cidx].efc + j];
}
