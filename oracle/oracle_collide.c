/* TEST INFRASTRUCTURE (see oracle_internal.h).  Collision detection for the bin-a-cube scene,
 * fp64: static candidate-pair list -> bounding-sphere + oriented-box cull -> narrow phase
 * (SURVEY.md Appendix A step 3).  Box-box pairs use separating axes + face clipping (up to 8
 * points per pair, MuJoCo engine_collision_box.c semantics: common normal from geom1 to
 * geom2, point midway between the surfaces, dist < 0); every pair involving a mesh -- the
 * table included -- goes through GJK + EPA on the hull support maps and yields ONE contact
 * (multiccd off), as MuJoCo's mjc_Convex does.  Where the contact point is not unique
 * (parallel faces) the choice below is this oracle's own; normal and depth are unique. */
#include "oracle_internal.h"

typedef struct {
  int type;            /* SO100_GEOM_BOX / SO100_GEOM_MESH */
  const double* pos;   /* frame origin (world) */
  const double* mat;   /* frame axes, row-major */
  const double* size;  /* box half sizes */
  const double (*vert)[3];
  int nvert;
} shape;

static void make_shape(const so100_model* m, const oenv* e, int g, shape* s) {
  s->type = m->geom_type[g];
  if (s->type == SO100_GEOM_BOX) {
    s->pos = e->gpos[g]; s->mat = e->gmat[g]; s->size = m->geom_size[g]; s->vert = 0; s->nvert = 0;
  } else {
    int b = m->geom_body[g];
    s->pos = e->xpos[b]; s->mat = e->xmat[b]; s->size = 0;
    s->vert = &m->vert[m->geom_vadr[g]]; s->nvert = m->geom_vnum[g];
  }
}

/* support point of a shape in world direction d */
static void support(const shape* s, const double* d, double* out) {
  double dl[3], pl[3];
  mulmtv(dl, s->mat, d);
  if (s->type == SO100_GEOM_BOX) {
    for (int k = 0; k < 3; k++) pl[k] = dl[k] >= 0 ? s->size[k] : -s->size[k];
  } else {
    int best = 0;
    double bv = -1e300;
    for (int i = 0; i < s->nvert; i++) {
      double v = dot3(s->vert[i], dl);
      if (v > bv) { bv = v; best = i; }
    }
    copy3(pl, s->vert[best]);
  }
  mulmv(out, s->mat, pl);
  add3(out, out, s->pos);
}

typedef struct { double w[3], a[3], b[3]; } mvert;   /* Minkowski-difference vertex A - B */

static void msupport(const shape* A, const shape* B, const double* d, mvert* v) {
  double nd[3] = {-d[0], -d[1], -d[2]};
  support(A, d, v->a);
  support(B, nd, v->b);
  sub3(v->w, v->a, v->b);
}

/* ---------------------------------------------------------------- GJK (boolean, keeps simplex) */
/* returns 1 when the simplex (after update) contains the origin; otherwise sets the next search dir */
static int do_simplex(mvert* s, int* n, double* dir) {
  double ao[3], ab[3], ac[3], ad[3], abc[3], t[3];
  if (*n == 2) {
    /* s[1] = A (newest), s[0] = B */
    scl3(ao, s[1].w, -1); sub3(ab, s[0].w, s[1].w);
    if (dot3(ab, ao) > 0) { cross3(t, ab, ao); cross3(dir, t, ab); }
    else { s[0] = s[1]; *n = 1; copy3(dir, ao); }
    return 0;
  }
  if (*n == 3) {
    /* A = s[2], B = s[1], C = s[0] */
    mvert A = s[2], B = s[1], C = s[0];
    scl3(ao, A.w, -1); sub3(ab, B.w, A.w); sub3(ac, C.w, A.w);
    cross3(abc, ab, ac);
    cross3(t, abc, ac);
    if (dot3(t, ao) > 0) {
      if (dot3(ac, ao) > 0) { s[0] = C; s[1] = A; *n = 2; cross3(t, ac, ao); cross3(dir, t, ac); }
      else goto star;
      return 0;
    }
    cross3(t, ab, abc);
    if (dot3(t, ao) > 0) {
    star:
      if (dot3(ab, ao) > 0) { s[0] = B; s[1] = A; *n = 2; cross3(t, ab, ao); cross3(dir, t, ab); }
      else { s[0] = A; *n = 1; copy3(dir, ao); }
      return 0;
    }
    if (dot3(abc, ao) > 0) { copy3(dir, abc); }
    else { s[0] = B; s[1] = C; s[2] = A; scl3(dir, abc, -1); }
    return 0;
  }
  /* tetrahedron: A = s[3], B = s[2], C = s[1], D = s[0] */
  {
    mvert A = s[3], B = s[2], C = s[1], D = s[0];
    double acd[3], adb[3];
    scl3(ao, A.w, -1); sub3(ab, B.w, A.w); sub3(ac, C.w, A.w); sub3(ad, D.w, A.w);
    cross3(abc, ab, ac); cross3(acd, ac, ad); cross3(adb, ad, ab);
    /* make the three face normals point away from the opposite vertex */
    if (dot3(abc, ad) > 0) scl3(abc, abc, -1);
    if (dot3(acd, ab) > 0) scl3(acd, acd, -1);
    if (dot3(adb, ac) > 0) scl3(adb, adb, -1);
    if (dot3(abc, ao) > 0) { s[0] = C; s[1] = B; s[2] = A; *n = 3; return do_simplex(s, n, dir); }
    if (dot3(acd, ao) > 0) { s[0] = D; s[1] = C; s[2] = A; *n = 3; return do_simplex(s, n, dir); }
    if (dot3(adb, ao) > 0) { s[0] = B; s[1] = D; s[2] = A; *n = 3; return do_simplex(s, n, dir); }
    return 1;
  }
}

/* ---------------------------------------------------------------- EPA */
#define EPA_MAXV 192
#define EPA_MAXF 384
typedef struct { int v[3]; double n[3], dplane, dtri; int alive; } eface;

/* closest point on triangle (a,b,c) to the origin: barycentric weights out, returns squared distance */
static double tri_closest(const double* a, const double* b, const double* c, double* lam) {
  double ab[3], ac[3], ap[3];
  sub3(ab, b, a); sub3(ac, c, a); scl3(ap, a, -1);
  double d1 = dot3(ab, ap), d2 = dot3(ac, ap);
  if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; return dot3(a, a); }
  double bp[3]; scl3(bp, b, -1);
  double d3 = dot3(ab, bp), d4 = dot3(ac, bp);
  if (d3 >= 0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; return dot3(b, b); }
  double vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    double v = d1 / (d1 - d3);
    lam[0] = 1 - v; lam[1] = v; lam[2] = 0;
  } else {
    double cp[3]; scl3(cp, c, -1);
    double d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    if (d6 >= 0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; return dot3(c, c); }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
      double w = d2 / (d2 - d6);
      lam[0] = 1 - w; lam[1] = 0; lam[2] = w;
    } else {
      double va = d3 * d6 - d5 * d4;
      if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        double w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        lam[0] = 0; lam[1] = 1 - w; lam[2] = w;
      } else {
        double den = 1.0 / (va + vb + vc);
        lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1 - lam[1] - lam[2];
      }
    }
  }
  double p[3];
  for (int k = 0; k < 3; k++) p[k] = lam[0] * a[k] + lam[1] * b[k] + lam[2] * c[k];
  return dot3(p, p);
}

static int make_face(eface* f, const mvert* V, int i, int j, int k) {
  double ab[3], ac[3], lam[3];
  f->v[0] = i; f->v[1] = j; f->v[2] = k;
  sub3(ab, V[j].w, V[i].w); sub3(ac, V[k].w, V[i].w);
  cross3(f->n, ab, ac);
  double len = norm3(f->n);
  if (len < 1e-20) { f->alive = 0; return 0; }
  scl3(f->n, f->n, 1.0 / len);
  f->dplane = dot3(f->n, V[i].w);
  f->dtri = sqrt(tri_closest(V[i].w, V[j].w, V[k].w, lam));
  f->alive = 1;
  return 1;
}

/* penetration of A into B.  Returns 1 on penetration with: normal (A -> B), depth > 0,
   witness points pa (on A) and pb (on B). */
static int gjk_epa(const shape* A, const shape* B, const double* ca, const double* cb,
                   double* normal, double* depth, double* pa, double* pb) {
  mvert s[4];
  int n = 0;
  double dir[3];
  sub3(dir, cb, ca);
  if (dot3(dir, dir) < 1e-24) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
  msupport(A, B, dir, &s[0]);
  n = 1;
  scl3(dir, s[0].w, -1);
  int hit = 0;
  for (int it = 0; it < 128; it++) {
    double dd = dot3(dir, dir);
    if (dd < 1e-30) return 0;                 /* origin on the simplex boundary: touching, dist >= 0 */
    mvert w;
    msupport(A, B, dir, &w);
    if (dot3(w.w, dir) <= 0) return 0;        /* separating direction found */
    s[n++] = w;
    if (do_simplex(s, &n, dir)) { hit = 1; break; }
  }
  if (!hit) return 0;

  /* EPA from the enclosing tetrahedron */
  static _Thread_local mvert V[EPA_MAXV];
  static _Thread_local eface F[EPA_MAXF];
  int nvtx = 4, nf = 0;
  for (int i = 0; i < 4; i++) V[i] = s[i];
  static const int tf[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
  for (int i = 0; i < 4; i++) {
    int a = tf[i][0], b = tf[i][1], c = tf[i][2], opp = 6 - a - b - c;
    if (!make_face(&F[nf], V, a, b, c)) return 0;   /* flat tetrahedron: treat as touching */
    double t[3];
    sub3(t, V[opp].w, V[a].w);
    if (dot3(F[nf].n, t) > 0) {                      /* orient outward */
      make_face(&F[nf], V, a, c, b);
    }
    nf++;
  }
  int best = -1;
  for (int it = 0; it < 96; it++) {
    best = -1;
    double bd = 1e300;
    for (int i = 0; i < nf; i++)
      if (F[i].alive && F[i].dtri < bd) { bd = F[i].dtri; best = i; }
    if (best < 0) return 0;
    mvert w;
    msupport(A, B, F[best].n, &w);
    double dw = dot3(w.w, F[best].n);
    if (dw - F[best].dplane < 1e-11 || nvtx >= EPA_MAXV || nf + 64 >= EPA_MAXF) break;
    /* horizon of the faces visible from w */
    int edges[EPA_MAXF][2], ne = 0;
    for (int i = 0; i < nf; i++) {
      if (!F[i].alive) continue;
      double t[3];
      sub3(t, w.w, V[F[i].v[0]].w);
      if (dot3(F[i].n, t) > 1e-14) {
        F[i].alive = 0;
        for (int k = 0; k < 3; k++) {
          int ea = F[i].v[k], eb = F[i].v[(k + 1) % 3], found = -1;
          for (int q = 0; q < ne; q++)
            if (edges[q][0] == eb && edges[q][1] == ea) { found = q; break; }
          if (found >= 0) { edges[found][0] = edges[ne - 1][0]; edges[found][1] = edges[ne - 1][1]; ne--; }
          else { edges[ne][0] = ea; edges[ne][1] = eb; ne++; }
        }
      }
    }
    if (ne == 0) break;   /* numerically nothing visible: converged */
    V[nvtx] = w;
    for (int q = 0; q < ne && nf < EPA_MAXF; q++) {
      if (make_face(&F[nf], V, edges[q][0], edges[q][1], nvtx)) nf++;
    }
    nvtx++;
  }
  if (best < 0) return 0;
  const eface* f = &F[best];
  double lam[3];
  tri_closest(V[f->v[0]].w, V[f->v[1]].w, V[f->v[2]].w, lam);
  for (int k = 0; k < 3; k++) {
    pa[k] = lam[0] * V[f->v[0]].a[k] + lam[1] * V[f->v[1]].a[k] + lam[2] * V[f->v[2]].a[k];
    pb[k] = lam[0] * V[f->v[0]].b[k] + lam[1] * V[f->v[1]].b[k] + lam[2] * V[f->v[2]].b[k];
  }
  copy3(normal, f->n);
  *depth = f->dplane;
  return *depth > 0;
}

/* ---------------------------------------------------------------- oriented-box overlap (cull) */
static int obb_separated(const double* c1, const double* R1, const double* h1,
                         const double* c2, const double* R2, const double* h2) {
  double t[3], ax[15][3];
  int na = 0;
  sub3(t, c2, c1);
  for (int k = 0; k < 3; k++) { ax[na][0] = R1[k]; ax[na][1] = R1[3 + k]; ax[na][2] = R1[6 + k]; na++; }
  for (int k = 0; k < 3; k++) { ax[na][0] = R2[k]; ax[na][1] = R2[3 + k]; ax[na][2] = R2[6 + k]; na++; }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      cross3(ax[na], ax[i], ax[3 + j]);
      if (dot3(ax[na], ax[na]) > 1e-12) na++;
    }
  for (int a = 0; a < na; a++) {
    double r1 = 0, r2 = 0;
    for (int k = 0; k < 3; k++) { r1 += h1[k] * fabs(dot3(ax[a], ax[k])); r2 += h2[k] * fabs(dot3(ax[a], ax[3 + k])); }
    if (fabs(dot3(t, ax[a])) > r1 + r2) return 1;
  }
  return 0;
}

/* ---------------------------------------------------------------- box-box */
typedef struct { double p[3]; } pt3;

/* clip polygon against half-space  n.x <= d  (Sutherland-Hodgman) */
static int clip_poly(pt3* in, int nin, const double* n, double d, pt3* out) {
  int no = 0;
  for (int i = 0; i < nin; i++) {
    const double* a = in[i].p;
    const double* b = in[(i + 1) % nin].p;
    double da = dot3(n, a) - d, db = dot3(n, b) - d;
    if (da <= 0) copy3(out[no++].p, a);
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      double t = da / (da - db);
      for (int k = 0; k < 3; k++) out[no].p[k] = a[k] + t * (b[k] - a[k]);
      no++;
    }
  }
  return no;
}

static int box_box(const double* cA, const double* RA, const double* hA,
                   const double* cB, const double* RB, const double* hB,
                   double (*cpos)[3], double* cdist, double* normal) {
  double A[3][3], B[3][3], t[3];
  for (int k = 0; k < 3; k++) {
    A[k][0] = RA[k]; A[k][1] = RA[3 + k]; A[k][2] = RA[6 + k];
    B[k][0] = RB[k]; B[k][1] = RB[3 + k]; B[k][2] = RB[6 + k];
  }
  sub3(t, cB, cA);
  /* face axes */
  double best_face = -1e300; int face_code = -1;
  for (int k = 0; k < 6; k++) {
    const double* L = k < 3 ? A[k] : B[k - 3];
    double rA = 0, rB = 0;
    for (int j = 0; j < 3; j++) { rA += hA[j] * fabs(dot3(A[j], L)); rB += hB[j] * fabs(dot3(B[j], L)); }
    double sep = fabs(dot3(t, L)) - (rA + rB);
    if (sep > 0) return 0;
    if (sep > best_face) { best_face = sep; face_code = k; }
  }
  double best_edge = -1e300, best_sel = -1e300, edge_axis[3] = {0, 0, 0}; int ei = -1, ej = -1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double L[3];
      cross3(L, A[i], B[j]);
      double len = norm3(L);
      if (len * len < 1e-6) continue;   /* (near-)parallel edges: covered by the face axes */
      scl3(L, L, 1.0 / len);
      double rA = 0, rB = 0;
      for (int k = 0; k < 3; k++) { rA += hA[k] * fabs(dot3(A[k], L)); rB += hB[k] * fabs(dot3(B[k], L)); }
      /* edge axes are penalised by 2e-6 m / sin(angle): float32 round-off on nearly parallel edges must
         never beat a face axis in the CUDA path, and the oracle follows the same rule */
      double sep = fabs(dot3(t, L)) - (rA + rB);
      if (sep > 0) return 0;
      double sel = sep - 2e-6 / len;
      if (sel > best_sel) { best_sel = sel; best_edge = sep; copy3(edge_axis, L); ei = i; ej = j; }
    }
  /* an edge axis must beat the best face axis by 5% (faces preferred: stable manifolds) */
  if (ei >= 0 && best_sel * 1.05 > best_face) {
    double n[3];
    copy3(n, edge_axis);
    if (dot3(n, t) < 0) scl3(n, n, -1);
    /* edge of A farthest along +n, edge of B farthest along -n */
    double pA[3], pB[3];
    copy3(pA, cA); copy3(pB, cB);
    for (int k = 0; k < 3; k++) {
      if (k != ei) addscl3(pA, pA, A[k], dot3(n, A[k]) > 0 ? hA[k] : -hA[k]);
      if (k != ej) addscl3(pB, pB, B[k], dot3(n, B[k]) > 0 ? -hB[k] : hB[k]);
    }
    /* closest points of the two lines pA + s*A[ei], pB + u*B[ej] */
    double r[3];
    sub3(r, pA, pB);
    double a = 1, b = dot3(A[ei], B[ej]), c = 1, d = dot3(A[ei], r), e_ = dot3(B[ej], r);
    double den = a * c - b * b;
    double s = (b * e_ - c * d) / den, u = (a * e_ - b * d) / den;
    double qa[3], qb[3];
    addscl3(qa, pA, A[ei], s);
    addscl3(qb, pB, B[ej], u);
    for (int k = 0; k < 3; k++) cpos[0][k] = 0.5 * (qa[k] + qb[k]);
    cdist[0] = best_edge;
    copy3(normal, n);
    return 1;
  }
  /* face contact: reference box owns the axis */
  int refA = face_code < 3;
  const double (*Rf)[3] = refA ? A : B;
  const double (*If)[3] = refA ? B : A;
  const double* cR = refA ? cA : cB; const double* cI = refA ? cB : cA;
  const double* hR = refA ? hA : hB; const double* hI = refA ? hB : hA;
  int ax = refA ? face_code : face_code - 3;
  double tri[3];
  sub3(tri, cI, cR);                         /* reference -> incident */
  double nref[3];
  copy3(nref, Rf[ax]);
  if (dot3(nref, tri) < 0) scl3(nref, nref, -1);   /* outward normal of the reference face */
  /* incident face: most anti-parallel to nref */
  int iax = 0; double bestd = -1;
  for (int k = 0; k < 3; k++) { double d = fabs(dot3(If[k], nref)); if (d > bestd) { bestd = d; iax = k; } }
  double isgn = dot3(If[iax], nref) > 0 ? -1 : 1;
  double fc[3];
  addscl3(fc, cI, If[iax], isgn * hI[iax]);
  int u = (iax + 1) % 3, v = (iax + 2) % 3;
  pt3 poly[16], tmp[16];
  static const int su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
  for (int q = 0; q < 4; q++)
    for (int k = 0; k < 3; k++) poly[q].p[k] = fc[k] + su[q] * hI[u] * If[u][k] + sv[q] * hI[v] * If[v][k];
  int np = 4;
  /* clip against the four side planes of the reference face */
  for (int k = 0; k < 3 && np > 0; k++) {
    if (k == ax) continue;
    double nn[3];
    copy3(nn, Rf[k]);
    np = clip_poly(poly, np, nn, dot3(nn, cR) + hR[k], tmp);
    scl3(nn, nn, -1);
    np = clip_poly(tmp, np, nn, dot3(nn, cR) + hR[k], poly);
  }
  double dref = dot3(nref, cR) + hR[ax];
  int nc = 0;
  for (int q = 0; q < np && nc < 8; q++) {
    double depth = dref - dot3(nref, poly[q].p);
    if (depth <= 0) continue;
    addscl3(cpos[nc], poly[q].p, nref, 0.5 * depth);
    cdist[nc] = -depth;
    nc++;
  }
  copy3(normal, nref);
  if (!refA) scl3(normal, normal, -1);       /* always geom1 -> geom2 */
  return nc;
}

/* mju_makeFrame: complete frame[0:3] (normal) to an orthonormal basis */
static void make_frame(double* fr) {
  normalize3(fr);
  double* y = fr + 3;
  y[0] = 0; y[1] = 0; y[2] = 0;
  if (fr[1] < 0.5 && fr[1] > -0.5) y[1] = 1; else y[2] = 1;
  double d = dot3(fr, y);
  addscl3(y, y, fr, -d);
  normalize3(y);
  cross3(fr + 6, fr, y);
}

static void add_contact(const so100_model* m, oenv* e, int pair, const double* pos, const double* normal, double dist) {
  if (e->ncon >= MAXCON) { e->overflow++; return; }
  ocontact* c = &e->con[e->ncon++];
  c->pair = pair; c->g1 = m->pair_g1[pair]; c->g2 = m->pair_g2[pair];
  c->dim = m->pair_condim[pair];
  c->dist = dist;
  copy3(c->pos, pos);
  copy3(c->frame, normal);
  make_frame(c->frame);
  memcpy(c->friction, m->pair_friction[pair], sizeof(c->friction));
  memcpy(c->solref, m->pair_solref[pair], sizeof(c->solref));
  memcpy(c->solimp, m->pair_solimp[pair], sizeof(c->solimp));
  c->mu = 0; c->efc = -1;
  memset(c->force, 0, sizeof(c->force));
}

/* Vertices of a shape within `tol` of its support plane in direction d: count, centroid (world), radius
   and the world vector from the centroid to the farthest member (the edge direction when count == 2). */
static int support_set(const shape* s, const double* d, double tol, double* centroid, double* radius, double* far) {
  double dl[3], best = -1e300, acc[3] = {0, 0, 0}, fl[3] = {0, 0, 0};
  int n = s->type == SO100_GEOM_BOX ? 8 : s->nvert, cnt = 0;
  mulmtv(dl, s->mat, d);
  for (int pass = 0; pass < 3; pass++) {
    double r2 = -1;
    for (int i = 0; i < n; i++) {
      double v[3];
      if (s->type == SO100_GEOM_BOX) { v[0] = (i & 1 ? 1 : -1) * s->size[0]; v[1] = (i & 2 ? 1 : -1) * s->size[1]; v[2] = (i & 4 ? 1 : -1) * s->size[2]; }
      else copy3(v, s->vert[i]);
      double val = dot3(v, dl);
      if (pass == 0) { if (val > best) best = val; continue; }
      if (val < best - tol) continue;
      if (pass == 1) { add3(acc, acc, v); cnt++; }
      else { double t[3]; sub3(t, v, acc); double q = dot3(t, t); if (q > r2) { r2 = q; copy3(fl, t); } }
    }
    if (pass == 1) scl3(acc, acc, 1.0 / cnt);
    if (pass == 2) *radius = sqrt(r2 > 0 ? r2 : 0);
  }
  mulmv(centroid, s->mat, acc);
  add3(centroid, centroid, s->pos);
  mulmv(far, s->mat, fl);
  return cnt;
}

/* Contact point of a GJK/EPA hit.  The EPA witness midpoint is unique for vertex-face and for crossing
   edge-edge configurations but arbitrary when faces, a face and an edge, or two edges are parallel; there
   the point is defined as the centroid of the deepest feature (vertices within 1e-6 m of the support plane
   along the normal) of the geom whose deepest feature is smaller, moved half the depth towards the other
   geom.  This keeps flat resting contacts torque-free and makes the point independent of the
   floating-point path. */
static void deepest_feature_point(const shape* A, const shape* B, const double* n, double depth, double* pos) {
  double cA[3], cB[3], rA, rB, fA[3], fB[3], nn[3] = {-n[0], -n[1], -n[2]};
  int nA = support_set(A, n, 1e-6, cA, &rA, fA), nB = support_set(B, nn, 1e-6, cB, &rB, fB);
  if (nA == 2 && nB == 2) {
    double x[3];
    cross3(x, fA, fB);
    if (dot3(x, x) > 1e-6 * dot3(fA, fA) * dot3(fB, fB)) return;    /* crossing edges: EPA witness is unique */
  }
  if (nA == 1 || (nB != 1 && rA <= rB)) addscl3(pos, cA, n, -0.5 * depth);
  else addscl3(pos, cB, n, 0.5 * depth);
}

/* A contact normal within 1e-3 rad of a face normal of a cuboid geom in the pair IS that face normal (the
   cuboid's face carries the contact); EPA only resolves it to ~sqrt(tolerance / size).  Snap it and take the
   penetration along the snapped direction, so that the deepest-feature selection below sees exact heights. */
static void snap_normal(const shape* A, int cubA, const shape* B, int cubB, double* n, double* depth) {
  for (int s = 0; s < 2; s++) {
    const shape* X = s == 0 ? A : B;
    if (!(s == 0 ? cubA : cubB)) continue;
    for (int k = 0; k < 3; k++) {
      double ax[3] = {X->mat[k], X->mat[3 + k], X->mat[6 + k]};
      double c = dot3(n, ax);
      if (fabs(c) > 1.0 - 5e-7) {
        double pa[3], pb[3], nn[3];
        scl3(n, ax, c > 0 ? 1.0 : -1.0);
        scl3(nn, n, -1.0);
        support(A, n, pa);
        support(B, nn, pb);
        *depth = dot3(n, pa) - dot3(n, pb);
        return;
      }
    }
  }
}

/* box geom, or mesh whose hull vertices are exactly the 8 corners of its bounding box */
static int is_cuboid(const so100_model* m, int g) {
  if (m->geom_type[g] == SO100_GEOM_BOX) return 1;
  if (m->geom_vnum[g] != 8) return 0;
  for (int v = 0; v < 8; v++)
    for (int k = 0; k < 3; k++)
      if (fabs(fabs(m->vert[m->geom_vadr[g] + v][k] - m->geom_center[g][k]) - m->geom_half[g][k]) > 1e-9) return 0;
  return 1;
}

/* entry used by the tests to cross-check the two narrow phases on box pairs */
int so100o_test_box_pair(const double* cA, const double* RA, const double* hA, const double* cB, const double* RB,
                         const double* hB, double* sat_out /* n[3], depth */, double* epa_out /* n[3], depth */) {
  double cpos[8][3], cdist[8], normal[3];
  int nc = box_box(cA, RA, hA, cB, RB, hB, cpos, cdist, normal);
  double dmin = 0;
  for (int k = 0; k < nc; k++) if (cdist[k] < dmin) dmin = cdist[k];
  copy3(sat_out, normal); sat_out[3] = -dmin;
  shape A = {SO100_GEOM_BOX, cA, RA, hA, 0, 0}, B = {SO100_GEOM_BOX, cB, RB, hB, 0, 0};
  double n2[3] = {0, 0, 0}, depth = 0, pa[3], pb[3];
  int hit = gjk_epa(&A, &B, cA, cB, n2, &depth, pa, pb);
  copy3(epa_out, n2); epa_out[3] = hit ? depth : 0;
  return nc * 16 + hit;
}

/* entry used by the tests to check the box-box manifold (points, depths, normal) against an independent computation */
int so100o_test_box_manifold(const double* cA, const double* RA, const double* hA, const double* cB, const double* RB,
                             const double* hB, double* normal_out, double* pos_out /* [8][3] */, double* dist_out /* [8] */) {
  double cpos[8][3], cdist[8], normal[3] = {0, 0, 0};
  int nc = box_box(cA, RA, hA, cB, RB, hB, cpos, cdist, normal);
  copy3(normal_out, normal);
  for (int k = 0; k < nc; k++) { copy3(pos_out + 3 * k, cpos[k]); dist_out[k] = cdist[k]; }
  return nc;
}

void o_collide(const so100_model* m, oenv* e) {
  e->ncon = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_g1[p], g2 = m->pair_g2[p];
    /* bounding spheres */
    double d[3];
    sub3(d, e->gcen[g2], e->gcen[g1]);
    double rr = m->geom_rbound[g1] + m->geom_rbound[g2];
    if (dot3(d, d) > rr * rr) continue;
    /* oriented bounding boxes (a box geom's OBB is the box itself) */
    const double* R1 = m->geom_type[g1] == SO100_GEOM_BOX ? e->gmat[g1] : e->xmat[m->geom_body[g1]];
    const double* R2 = m->geom_type[g2] == SO100_GEOM_BOX ? e->gmat[g2] : e->xmat[m->geom_body[g2]];
    if (obb_separated(e->gcen[g1], R1, m->geom_half[g1], e->gcen[g2], R2, m->geom_half[g2])) continue;
    if (is_cuboid(m, g1) && is_cuboid(m, g2)) {
      /* Two boxes.  A mesh whose hull is an exact cuboid (the table) has the support map of a box, so
         the separating-axis result equals the GJK/EPA one (tests/test_oracle_collision.py checks this);
         pairs that involve such a mesh keep mjc_Convex's one-contact-per-pair rule: deepest feature,
         centroid of the deepest points (within 1e-6 m) when it is not a single vertex. */
      double cpos[8][3], cdist[8], normal[3];
      int nc = box_box(e->gcen[g1], R1, m->geom_half[g1], e->gcen[g2], R2, m->geom_half[g2], cpos, cdist, normal);
      int single = m->geom_type[g1] != SO100_GEOM_BOX || m->geom_type[g2] != SO100_GEOM_BOX;
      if (single && nc > 1) {
        double dmin = cdist[0], acc[3] = {0, 0, 0};
        int cnt = 0;
        for (int k = 1; k < nc; k++) if (cdist[k] < dmin) dmin = cdist[k];
        for (int k = 0; k < nc; k++)
          if (cdist[k] <= dmin + 1e-6) { add3(acc, acc, cpos[k]); cnt++; }
        scl3(cpos[0], acc, 1.0 / cnt);
        cdist[0] = dmin;
        nc = 1;
      }
      for (int k = 0; k < nc; k++) add_contact(m, e, p, cpos[k], normal, cdist[k]);
    } else {
      shape A, B;
      double normal[3], depth, pa[3], pb[3], pos[3];
      make_shape(m, e, g1, &A);
      make_shape(m, e, g2, &B);
      if (gjk_epa(&A, &B, e->gcen[g1], e->gcen[g2], normal, &depth, pa, pb)) {
        for (int k = 0; k < 3; k++) pos[k] = 0.5 * (pa[k] + pb[k]);
        snap_normal(&A, is_cuboid(m, g1), &B, is_cuboid(m, g2), normal, &depth);
        deepest_feature_point(&A, &B, normal, depth, pos);
        add_contact(m, e, p, pos, normal, -depth);
      }
    }
  }
}
