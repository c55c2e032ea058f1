/* TEST INFRASTRUCTURE (see oracle_internal.h).  Smooth dynamics + constraint construction of
 * one MuJoCo substep for the bin-a-cube model, fp64, following SURVEY.md Appendix A steps
 * 1-2, 4-7, 9 (MuJoCo files restated from their published algorithms: engine_core_smooth.c
 * mj_kinematics / mj_crb / mj_rne, engine_forward.c mj_fwdActuation / mj_fwdAcceleration /
 * mj_Euler, engine_core_constraint.c mj_makeConstraint / mj_makeImpedance). */
#include "oracle_internal.h"

/* ------------------------------------------------------------------ kinematics (App. A step 1) */
static void o_kinematics(const so100_model* m, oenv* e) {
  memset(e->xpos[0], 0, sizeof(e->xpos[0]));
  e->xquat[0][0] = 1; e->xquat[0][1] = e->xquat[0][2] = e->xquat[0][3] = 0;
  quat2mat(e->xmat[0], e->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b], jt = m->body_jtype[b];
    if (jt == SO100_JNT_FREE) {
      int a = m->body_qposadr[b];
      copy3(e->xpos[b], e->qpos + a);
      memcpy(e->xquat[b], e->qpos + a + 3, 4 * sizeof(double));
      quat_normalize(e->xquat[b]);
    } else {
      double r[3];
      mulmv(r, e->xmat[p], m->body_pos[b]);
      add3(e->xpos[b], e->xpos[p], r);
      quat_mul(e->xquat[b], e->xquat[p], m->body_quat[b]);
      if (jt == SO100_JNT_HINGE) {
        int a = m->body_qposadr[b];
        double ang = e->qpos[a] - m->qpos0[a], s = sin(0.5 * ang);
        double ql[4] = {cos(0.5 * ang), m->body_jaxis[b][0] * s, m->body_jaxis[b][1] * s, m->body_jaxis[b][2] * s};
        double q[4];
        quat_mul(q, e->xquat[b], ql);
        memcpy(e->xquat[b], q, sizeof(q));
        /* joint anchor is the body origin (jnt_pos = 0), so xpos needs no off-centre correction */
      }
    }
    quat_normalize(e->xquat[b]);
    quat2mat(e->xmat[b], e->xquat[b]);
    double r[3], qi[4];
    mulmv(r, e->xmat[b], m->body_ipos[b]);
    add3(e->xipos[b], e->xpos[b], r);
    quat_mul(qi, e->xquat[b], m->body_iquat[b]);
    quat2mat(e->ximat[b], qi);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_body[g];
    double r[3], q[4];
    mulmv(r, e->xmat[b], m->geom_pos[g]);
    add3(e->gpos[g], e->xpos[b], r);
    quat_mul(q, e->xquat[b], m->geom_quat[g]);
    quat2mat(e->gmat[g], q);
    mulmv(r, e->xmat[b], m->geom_center[g]);
    add3(e->gcen[g], e->xpos[b], r);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_body[s];
    double r[3];
    mulmv(r, e->xmat[b], m->site_pos[s]);
    add3(e->site[s], e->xpos[b], r);
  }
}

/* Jacobian (3 x nv each, row-major with stride NVMAX) of a world point rigidly attached to `body`. */
void o_jac(const so100_model* m, const oenv* e, int body, const double* point, double* jp, double* jr) {
  memset(jp, 0, 3 * NVMAX * sizeof(double));
  memset(jr, 0, 3 * NVMAX * sizeof(double));
  for (int b = body; b > 0; b = m->body_parent[b]) {
    int jt = m->body_jtype[b], d = m->body_dofadr[b];
    if (jt == SO100_JNT_HINGE) {
      double ax[3], r[3], c[3];
      mulmv(ax, e->xmat[b], m->body_jaxis[b]);
      sub3(r, point, e->xpos[b]);
      cross3(c, ax, r);
      for (int k = 0; k < 3; k++) { jr[k * NVMAX + d] = ax[k]; jp[k * NVMAX + d] = c[k]; }
    } else if (jt == SO100_JNT_FREE) {
      double r[3];
      sub3(r, point, e->xpos[b]);
      for (int k = 0; k < 3; k++) {
        jp[k * NVMAX + d + k] = 1;
        double ax[3] = {e->xmat[b][0 + k], e->xmat[b][3 + k], e->xmat[b][6 + k]}, c[3]; /* local axis k in world */
        cross3(c, ax, r);
        for (int r_ = 0; r_ < 3; r_++) { jr[r_ * NVMAX + d + 3 + k] = ax[r_]; jp[r_ * NVMAX + d + 3 + k] = c[r_]; }
      }
    }
  }
}

/* dense Cholesky A = L L^T (lower, row-major n x n with stride NVMAX); returns rank deficiency */
int o_chol(double* L, const double* A, int n) {
  int bad = 0;
  memset(L, 0, NVMAX * NVMAX * sizeof(double));
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double s = A[i * NVMAX + j];
      for (int k = 0; k < j; k++) s -= L[i * NVMAX + k] * L[j * NVMAX + k];
      if (i == j) {
        if (s < MJMINVAL) { s = MJMINVAL; bad++; }
        L[i * NVMAX + i] = sqrt(s);
      } else {
        L[i * NVMAX + j] = s / L[j * NVMAX + j];
      }
    }
  return bad;
}
void o_chol_solve(const double* L, double* x, int n) {
  for (int i = 0; i < n; i++) {
    double s = x[i];
    for (int k = 0; k < i; k++) s -= L[i * NVMAX + k] * x[k];
    x[i] = s / L[i * NVMAX + i];
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k * NVMAX + i] * x[k];
    x[i] = s / L[i * NVMAX + i];
  }
}

/* mass matrix from its definition  M = sum_b m Jp^T Jp + Jr^T I_w Jr  (+ armature); App. A step 2 */
static void o_mass_matrix(const so100_model* m, oenv* e) {
  int nv = m->nv;
  memset(e->M, 0, sizeof(e->M));
  for (int b = 1; b < m->nbody; b++) {
    double mass = m->body_mass[b];
    if (mass <= 0 || m->body_weldid[b] == 0) continue;
    double jp[3 * NVMAX], jr[3 * NVMAX], Iw[9], tmp[9];
    o_jac(m, e, b, e->xipos[b], jp, jr);
    /* Iw = Ri diag Ri^T */
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) tmp[r * 3 + c] = e->ximat[b][r * 3 + c] * m->body_inertia[b][c];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += tmp[r * 3 + k] * e->ximat[b][c * 3 + k];
        Iw[r * 3 + c] = s;
      }
    for (int i = 0; i < nv; i++)
      for (int j = 0; j < nv; j++) {
        double s = 0;
        for (int k = 0; k < 3; k++) s += mass * jp[k * NVMAX + i] * jp[k * NVMAX + j];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) s += jr[r * NVMAX + i] * Iw[r * 3 + c] * jr[c * NVMAX + j];
        e->M[i * NVMAX + j] += s;
      }
  }
  for (int i = 0; i < nv; i++) e->M[i * NVMAX + i] += m->dof_armature[i];
  o_chol(e->Lchol, e->M, nv);
}

void o_position(const so100_model* m, oenv* e) {
  o_kinematics(m, e);
  o_mass_matrix(m, e);
  o_collide(m, e);
}

/* ------------------------------------------------------------------ RNE bias (App. A step 4) */
static void o_rne(const so100_model* m, oenv* e) {
  double w[NB][3], alpha[NB][3], acc[NB][3], f[NB][3], n[NB][3];
  int nv = m->nv;
  memset(w, 0, sizeof(w)); memset(alpha, 0, sizeof(alpha)); memset(acc, 0, sizeof(acc));
  memset(f, 0, sizeof(f)); memset(n, 0, sizeof(n));
  memset(e->bias, 0, sizeof(e->bias));
  /* world "accelerates" with -gravity */
  scl3(acc[0], m->gravity, -1.0);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parent[b], jt = m->body_jtype[b];
    if (m->body_weldid[b] == 0) { copy3(acc[b], acc[0]); continue; }
    if (jt == SO100_JNT_FREE) {
      /* translational dofs: world frame; rotational dofs: body frame.  The scene's free body has
         its centre of mass at the body origin (so100_transfer_cube.xml:9), asserted at create. */
      int d = m->body_dofadr[b];
      double wl[3] = {e->qvel[d + 3], e->qvel[d + 4], e->qvel[d + 5]}, ww[3], Iw_w[3], tmp[3], tq[3], tql[3];
      mulmv(ww, e->xmat[b], wl);
      /* I_w w = R diag R^T w */
      mulmtv(tmp, e->ximat[b], ww);
      for (int k = 0; k < 3; k++) tmp[k] *= m->body_inertia[b][k];
      mulmv(Iw_w, e->ximat[b], tmp);
      cross3(tq, ww, Iw_w);
      mulmtv(tql, e->xmat[b], tq);
      for (int k = 0; k < 3; k++) {
        e->bias[d + k] = m->body_mass[b] * acc[0][k];
        e->bias[d + 3 + k] = tql[k];
      }
      continue;
    }
    double r[3], t1[3], t2[3];
    sub3(r, e->xpos[b], e->xpos[p]);
    /* origin acceleration: a_b = a_p + alpha_p x r + w_p x (w_p x r) */
    cross3(t1, alpha[p], r);
    cross3(t2, w[p], r);
    cross3(t2, w[p], t2);
    add3(acc[b], acc[p], t1);
    add3(acc[b], acc[b], t2);
    copy3(w[b], w[p]);
    copy3(alpha[b], alpha[p]);
    if (jt == SO100_JNT_HINGE) {
      int d = m->body_dofadr[b];
      double ax[3], c[3];
      mulmv(ax, e->xmat[b], m->body_jaxis[b]);
      cross3(c, w[p], ax);
      addscl3(alpha[b], alpha[b], c, e->qvel[d]);   /* d/dt(axis) * qdot, qacc = 0 */
      addscl3(w[b], w[b], ax, e->qvel[d]);
    }
    /* inertial force / torque about the centre of mass */
    double c[3], ac[3], Iw_a[3], Iw_w[3], tmp[3], N[3];
    sub3(c, e->xipos[b], e->xpos[b]);
    cross3(t1, alpha[b], c);
    cross3(t2, w[b], c);
    cross3(t2, w[b], t2);
    add3(ac, acc[b], t1);
    add3(ac, ac, t2);
    scl3(f[b], ac, m->body_mass[b]);
    mulmtv(tmp, e->ximat[b], alpha[b]);
    for (int k = 0; k < 3; k++) tmp[k] *= m->body_inertia[b][k];
    mulmv(Iw_a, e->ximat[b], tmp);
    mulmtv(tmp, e->ximat[b], w[b]);
    for (int k = 0; k < 3; k++) tmp[k] *= m->body_inertia[b][k];
    mulmv(Iw_w, e->ximat[b], tmp);
    cross3(N, w[b], Iw_w);
    add3(N, N, Iw_a);
    cross3(t1, c, f[b]);
    add3(n[b], N, t1);    /* moment about the body origin */
  }
  for (int b = m->nbody - 1; b >= 1; b--) {
    int p = m->body_parent[b], jt = m->body_jtype[b];
    if (m->body_weldid[b] == 0 || jt == SO100_JNT_FREE) continue;
    if (jt == SO100_JNT_HINGE) {
      double ax[3];
      mulmv(ax, e->xmat[b], m->body_jaxis[b]);
      e->bias[m->body_dofadr[b]] = dot3(ax, n[b]);
    }
    if (p > 0 && m->body_weldid[p] != 0) {
      double r[3], t[3];
      sub3(r, e->xpos[b], e->xpos[p]);
      cross3(t, r, f[b]);
      add3(n[p], n[p], n[b]);
      add3(n[p], n[p], t);
      add3(f[p], f[p], f[b]);
    }
  }
  (void)nv;
}

/* bias, position actuators, qacc_smooth (App. A steps 4-6) */
void o_velocity_actuation(const so100_model* m, oenv* e) {
  int nv = m->nv;
  o_rne(m, e);
  memset(e->qfrc_act, 0, sizeof(e->qfrc_act));
  for (int a = 0; a < m->nu; a++) {
    int d = m->act_dof[a];
    int qa = m->body_qposadr[m->dof_body[d]];
    double u = e->ctrl[a];
    if (u < m->act_ctrlrange[a][0]) u = m->act_ctrlrange[a][0];
    if (u > m->act_ctrlrange[a][1]) u = m->act_ctrlrange[a][1];
    double frc = m->act_kp[a] * u - m->act_kp[a] * e->qpos[qa] - m->act_kv[a] * e->qvel[d];
    if (frc < m->act_forcerange[a][0]) frc = m->act_forcerange[a][0];
    if (frc > m->act_forcerange[a][1]) frc = m->act_forcerange[a][1];
    e->qfrc_act[d] += frc;
  }
  for (int i = 0; i < nv; i++) {
    e->qfrc_smooth[i] = e->qfrc_act[i] - e->bias[i]; /* qfrc_passive == 0 for this model */
    e->qacc_smooth[i] = e->qfrc_smooth[i];
  }
  o_chol_solve(e->Lchol, e->qacc_smooth, nv);
}

/* ------------------------------------------------------------------ constraints (App. A step 7) */
static void clamp_solimp(double* s) {
  if (s[0] < MJMINIMP) s[0] = MJMINIMP; if (s[0] > MJMAXIMP) s[0] = MJMAXIMP;
  if (s[1] < MJMINIMP) s[1] = MJMINIMP; if (s[1] > MJMAXIMP) s[1] = MJMAXIMP;
  if (s[2] < 0) s[2] = 0;
  if (s[3] < MJMINIMP) s[3] = MJMINIMP; if (s[3] > MJMAXIMP) s[3] = MJMAXIMP;
  if (s[4] < 1) s[4] = 1;
}
static double impedance(const double* solimp, double dist) {
  if (solimp[0] == solimp[1] || solimp[2] <= MJMINVAL) return 0.5 * (solimp[0] + solimp[1]);
  double x = fabs(dist) / solimp[2];
  if (x >= 1) return solimp[1];
  if (x <= 0) return solimp[0];
  double y, p = solimp[4], mid = solimp[3];
  if (p == 1) y = x;
  else if (x <= mid) y = pow(x, p) / pow(mid, p - 1);
  else y = 1 - pow(1 - x, p) / pow(1 - mid, p - 1);
  return solimp[0] + y * (solimp[1] - solimp[0]);
}
/* fills R, D, aref for row i given pos/vel/diagApprox; returns imp */
static void finish_row(const so100_model* m, oenv* e, int i, const double* solref_in, const double* solimp_in,
                       int is_friction) {
  double solimp[5], solref[2] = {solref_in[0], solref_in[1]};
  memcpy(solimp, solimp_in, sizeof(solimp));
  clamp_solimp(solimp);
  if (solref[0] > 0 && solref[0] < 2 * m->timestep) solref[0] = 2 * m->timestep; /* refsafe */
  double imp = impedance(solimp, e->epos[i]);
  double R = (1 - imp) / imp * e->ediag[i];
  if (R < MJMINVAL) R = MJMINVAL;
  double dmax = solimp[1];
  double K = 1.0 / fmax(MJMINVAL, dmax * dmax * solref[0] * solref[0] * solref[1] * solref[1]);
  double B = 2.0 / fmax(MJMINVAL, dmax * solref[0]);
  if (is_friction) K = 0;
  e->eR[i] = R;
  e->eD[i] = 1.0 / R;
  e->earef[i] = -B * e->evel[i] - K * imp * e->epos[i];
}

void o_make_constraints(const so100_model* m, oenv* e) {
  static const double def_solref[2] = {0.02, 1.0}, def_solimp[5] = {0.9, 0.95, 0.001, 0.5, 2.0};
  int nv = m->nv, n = 0;
  /* (i) dof frictionloss rows */
  for (int d = 0; d < nv; d++) {
    if (m->dof_frictionloss[d] <= 0) continue;
    memset(e->J[n], 0, sizeof(e->J[n]));
    e->J[n][d] = 1;
    e->etype[n] = EFC_FRICTION; e->eid[n] = d;
    e->epos[n] = 0; e->evel[n] = e->qvel[d];
    e->ediag[n] = m->dof_invweight0[d];
    e->efloss[n] = m->dof_frictionloss[d];
    finish_row(m, e, n, def_solref, def_solimp, 1);
    n++;
  }
  /* (ii) joint limits, active when violated (margin 0) */
  for (int d = 0; d < nv; d++) {
    if (!m->dof_limited[d]) continue;
    int qa = m->body_qposadr[m->dof_body[d]];
    double q = e->qpos[qa];
    for (int side = 0; side < 2; side++) {
      double dist = side == 0 ? q - m->dof_range[d][0] : m->dof_range[d][1] - q;
      if (dist >= 0) continue;
      memset(e->J[n], 0, sizeof(e->J[n]));
      e->J[n][d] = side == 0 ? 1 : -1;
      e->etype[n] = EFC_LIMIT; e->eid[n] = d;
      e->epos[n] = dist; e->evel[n] = e->J[n][d] * e->qvel[d];
      e->ediag[n] = m->dof_invweight0[d];
      e->efloss[n] = 0;
      finish_row(m, e, n, def_solref, def_solimp, 0);
      n++;
    }
  }
  /* (iii) elliptic contacts */
  for (int c = 0; c < e->ncon; c++) {
    ocontact* con = &e->con[c];
    int b1 = m->geom_body[con->g1], b2 = m->geom_body[con->g2];
    double jp1[3 * NVMAX], jr1[3 * NVMAX], jp2[3 * NVMAX], jr2[3 * NVMAX];
    o_jac(m, e, b1, con->pos, jp1, jr1);
    o_jac(m, e, b2, con->pos, jp2, jr2);
    con->efc = n;
    double tran = m->body_invweight0[b1][0] + m->body_invweight0[b2][0];
    double rot = m->body_invweight0[b1][1] + m->body_invweight0[b2][1];
    for (int r = 0; r < con->dim; r++) {
      int i = n + r;
      const double* ax = con->frame + 3 * (r < 3 ? r : 0);
      double v = 0;
      for (int d = 0; d < nv; d++) {
        double s = 0;
        for (int k = 0; k < 3; k++)
          s += ax[k] * (r < 3 ? (jp2[k * NVMAX + d] - jp1[k * NVMAX + d]) : (jr2[k * NVMAX + d] - jr1[k * NVMAX + d]));
        e->J[i][d] = s;
        v += s * e->qvel[d];
      }
      e->etype[i] = r == 0 ? EFC_CONTACT : EFC_CONTACT_CONT; e->eid[i] = c;
      e->epos[i] = r == 0 ? con->dist : 0; e->evel[i] = v;
      e->ediag[i] = r < 3 ? tran : rot;
      e->efloss[i] = 0;
      finish_row(m, e, i, con->solref, con->solimp, r > 0);
    }
    /* elliptic cone: regularisation of the friction rows follows the normal row */
    int i0 = n;
    if (con->dim > 1) {
      double fr[3] = {con->friction[0], con->friction[0], con->friction[1]}; /* t1, t2, torsion */
      e->eR[i0 + 1] = e->eR[i0] / fmax(MJMINVAL, m->impratio);
      con->mu = con->friction[0] * sqrt(e->eR[i0 + 1] / e->eR[i0]);
      for (int j = 2; j < con->dim; j++) e->eR[i0 + j] = e->eR[i0 + 1] * fr[0] * fr[0] / (fr[j - 1] * fr[j - 1]);
      for (int j = 1; j < con->dim; j++) e->eD[i0 + j] = 1.0 / e->eR[i0 + j];
    }
    n += con->dim;
  }
  e->nefc = n;
}

/* ------------------------------------------------------------------ semi-implicit Euler (App. A step 9) */
void o_integrate(const so100_model* m, oenv* e) {
  double h = m->timestep;
  for (int i = 0; i < m->nv; i++) e->qvel[i] += h * e->qacc[i];
  for (int b = 1; b < m->nbody; b++) {
    int jt = m->body_jtype[b];
    if (jt == SO100_JNT_HINGE) {
      e->qpos[m->body_qposadr[b]] += h * e->qvel[m->body_dofadr[b]];
    } else if (jt == SO100_JNT_FREE) {
      int qa = m->body_qposadr[b], d = m->body_dofadr[b];
      for (int k = 0; k < 3; k++) e->qpos[qa + k] += h * e->qvel[d + k];
      double w[3] = {e->qvel[d + 3], e->qvel[d + 4], e->qvel[d + 5]};
      double ang = norm3(w) * h;
      if (ang > 0) {
        double ax[3] = {w[0], w[1], w[2]};
        normalize3(ax);
        double s = sin(0.5 * ang), dq[4] = {cos(0.5 * ang), ax[0] * s, ax[1] * s, ax[2] * s}, q[4];
        quat_mul(q, e->qpos + qa + 3, dq);
        quat_normalize(q);
        memcpy(e->qpos + qa + 3, q, sizeof(q));
      } else {
        quat_normalize(e->qpos + qa + 3);
      }
    }
  }
  memcpy(e->warm, e->qacc, sizeof(e->warm));
}
