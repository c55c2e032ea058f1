// Collision, stage A (App. A step 3): static candidate-pair list -> sphere + AABB cull ->
//   box-like pairs: separating-axis test with one lane per pair, then contact generation with the
//                   whole tile per penetrating pair (24 candidate points, one per lane: incident-face
//                   corners, incident-edge x reference-edge intersections, reference corners under the
//                   incident face -- the vertices of the clipped incident polygon);
//   hull pairs:     oriented-box cull only; survivors are listed in the workspace header and the env
//                   is queued for the GJK/EPA kernel (so100_gjk.cuh).
// Contacts are appended in pair order, so the contact list is deterministic.
#pragma once
#include "so100_scratch.cuh"

namespace so100 {

struct Obb { V3 c, ax[3]; float h[3]; };
// Edge-edge axes A_i x B_j: skipped below sin^2 = 1e-6 and penalised by 2e-6 m / sin so that
// round-off on nearly parallel edges can never beat a face axis (same rule in the oracle).
constexpr float EDGE_MIN_SIN2 = 1e-6f;
constexpr float EDGE_BIAS = 2e-6f;

__device__ __forceinline__ V3 geom_center(const FrameBlock& f, const DevGeom& g) {
  V3 c = ld3(g.center);
  if (g.link >= 0) c = ld3(f.lpos[g.link]) + mulmv(f.lmat[g.link], c);
  return c;
}

__device__ __forceinline__ void load_obb(const FrameBlock& f, const DevGeom& g, V3 center, Obb& b) {
  b.c = center;
  b.h[0] = g.half[0]; b.h[1] = g.half[1]; b.h[2] = g.half[2];
  const float* m = g.link >= 0 ? f.lmat[g.link] : g.wmat;
  b.ax[0] = mcol(m, 0); b.ax[1] = mcol(m, 1); b.ax[2] = mcol(m, 2);
}

__device__ __forceinline__ void put_contact(float* con, int c, V3 p, V3 n, float dist, int pair) {
  float4* q = reinterpret_cast<float4*>(con + c * CON_WORDS);
  q[0] = make_float4(p.x, p.y, p.z, n.x);
  q[1] = make_float4(n.y, n.z, dist, __int_as_float(pair));
}

// 15-axis separating-axis test.  Returns false when separated; otherwise `code` = 0..5 (face axis of
// A / B) or 6 + 3 i + j (edge axis A_i x B_j) and `sep` < 0 the signed separation along it.
// cull = true (hull pairs): overlap test only, with a 1e-6 slack on |R|.
__device__ __forceinline__ bool box_sat(const Obb& A, const Obb& B, bool cull, int& code, float& sep_out) {
  float R[3][3], aR[3][3], tA[3], tB[3];
  const V3 t = B.c - A.c;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    tA[i] = dot(t, A.ax[i]);
    tB[i] = dot(t, B.ax[i]);
#pragma unroll
    for (int j = 0; j < 3; j++) { R[i][j] = dot(A.ax[i], B.ax[j]); aR[i][j] = fabsf(R[i][j]) + (cull ? 1e-6f : 0.0f); }
  }
  float best_face = -1e30f;
  code = 0;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float sep = fabsf(tA[i]) - (A.h[i] + B.h[0] * aR[i][0] + B.h[1] * aR[i][1] + B.h[2] * aR[i][2]);
    if (sep > 0) return false;
    if (sep > best_face) { best_face = sep; code = i; }
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const float sep = fabsf(tB[j]) - (B.h[j] + A.h[0] * aR[0][j] + A.h[1] * aR[1][j] + A.h[2] * aR[2][j]);
    if (sep > 0) return false;
    if (sep > best_face) { best_face = sep; code = 3 + j; }
  }
  float best_edge = -1e30f, best_sel = -1e30f;
  int ecode = -1;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const float ra = A.h[i1] * aR[i2][j] + A.h[i2] * aR[i1][j];
      const float rb = B.h[j1] * aR[i][j2] + B.h[j2] * aR[i][j1];
      const float num = fabsf(tA[i2] * R[i1][j] - tA[i1] * R[i2][j]) - (ra + rb);
      if (cull) {
        if (num > 0) return false;
        continue;
      }
      const float l2 = 1.0f - R[i][j] * R[i][j];
      if (l2 < EDGE_MIN_SIN2) continue;      // (near-)parallel edges: the face axes cover this direction
      const float inv = rsqrtf(l2);
      const float sep = num * inv;
      if (sep > 0) return false;
      const float sel = sep - EDGE_BIAS * inv;
      if (sel > best_sel) { best_sel = sel; best_edge = sep; ecode = 6 + 3 * i + j; }
    }
  }
  sep_out = best_face;
  if (ecode >= 0 && best_sel * 1.05f > best_face) { code = ecode; sep_out = best_edge; }
  return true;
}

__device__ __forceinline__ V3 sel_axis(const Obb& b, int k) { return k == 0 ? b.ax[0] : (k == 1 ? b.ax[1] : b.ax[2]); }
__device__ __forceinline__ float sel_h(const Obb& b, int k) { return k == 0 ? b.h[0] : (k == 1 ? b.h[1] : b.h[2]); }

// Contact generation for one penetrating box pair with the whole tile.  Appends up to 8 contacts
// (1 when `single`) to `con` at index `base`; every lane returns the number appended.
template <unsigned LPE>
__device__ int box_contacts(const Tile<LPE>& t, float* con, int base, const Obb& A, const Obb& B, int code, float sep, bool single,
                            int pair) {
  static_assert(LPE == 16 || LPE == 32, "box_contacts: 24 candidate points over one pass of 32 lanes or two passes of 16");
  const int lane = t.thread_rank();
  const V3 tAB = B.c - A.c;
  if (code >= 6) {
    // edge-edge: one point midway between the closest points of the two edges
    const int ei = (code - 6) / 3, ej = (code - 6) % 3;
    const V3 ea = sel_axis(A, ei), eb = sel_axis(B, ej);
    V3 n = normalized(cross(ea, eb));
    if (dot(n, tAB) < 0) n = -n;
    V3 pA = A.c, pB = B.c;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k != ei) pA = pA + A.ax[k] * (dot(n, A.ax[k]) > 0 ? A.h[k] : -A.h[k]);
      if (k != ej) pB = pB + B.ax[k] * (dot(n, B.ax[k]) > 0 ? -B.h[k] : B.h[k]);
    }
    const V3 r = pA - pB;
    const float b = dot(ea, eb), d = dot(ea, r), e = dot(eb, r), den = 1.0f - b * b;
    const float s = (b * e - d) / den, u = (e - b * d) / den;
    const V3 q = ((pA + ea * s) + (pB + eb * u)) * 0.5f;
    if (lane == 0 && base < NC) put_contact(con, base, q, n, sep, pair);
    return 1;
  }
  // face contact: the reference box owns the axis
  const bool refA = code < 3;
  const Obb& Rf = refA ? A : B;
  const Obb& If = refA ? B : A;
  const int ax = refA ? code : code - 3;
  const int ua = (ax + 1) % 3, va = (ax + 2) % 3;
  const V3 axn = sel_axis(Rf, ax), axu = sel_axis(Rf, ua), axv = sel_axis(Rf, va);
  const float hn = sel_h(Rf, ax), hu = sel_h(Rf, ua), hv = sel_h(Rf, va);
  const V3 rel = If.c - Rf.c;
  const V3 nref = dot(axn, rel) < 0 ? -axn : axn;              // outward normal of the reference face
  // incident face: the face of the other box most anti-parallel to nref
  const float d0 = dot(If.ax[0], nref), d1 = dot(If.ax[1], nref), d2 = dot(If.ax[2], nref);
  int iax = 0; float bd = fabsf(d0), ds = d0;
  if (fabsf(d1) > bd) { bd = fabsf(d1); iax = 1; ds = d1; }
  if (fabsf(d2) > bd) { bd = fabsf(d2); iax = 2; ds = d2; }
  const float isg = ds > 0 ? -1.0f : 1.0f;
  const int iu = (iax + 1) % 3, iv = (iax + 2) % 3;
  const V3 fc = rel + sel_axis(If, iax) * (isg * sel_h(If, iax));
  const V3 eu = sel_axis(If, iu) * sel_h(If, iu), ev = sel_axis(If, iv) * sel_h(If, iv);
  // incident corners in reference coordinates (u, v, n), counter-clockwise in the incident face
  float pu[4], pv[4], pn[4];
  {
    const float su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const V3 p = fc + eu * su[q] + ev * sv[q];
      pu[q] = dot(p, axu); pv[q] = dot(p, axv); pn[q] = dot(p, nref);
    }
  }
  // candidate point k of 24 (reference coordinates u, v, n): true when it is a vertex of the clipped incident polygon
  auto candidate = [&](int k, float& cu, float& cv, float& cn) -> bool {
    bool valid = false;
    cu = 0; cv = 0; cn = 0;
    if (k < 4) {
      cu = k == 0 ? pu[0] : (k == 1 ? pu[1] : (k == 2 ? pu[2] : pu[3]));
      cv = k == 0 ? pv[0] : (k == 1 ? pv[1] : (k == 2 ? pv[2] : pv[3]));
      cn = k == 0 ? pn[0] : (k == 1 ? pn[1] : (k == 2 ? pn[2] : pn[3]));
      valid = fabsf(cu) <= hu && fabsf(cv) <= hv;
    } else if (k < 20) {
      const int e = (k - 4) >> 2, l = (k - 4) & 3, e2 = (e + 1) & 3;
      const float au = e == 0 ? pu[0] : (e == 1 ? pu[1] : (e == 2 ? pu[2] : pu[3]));
      const float av = e == 0 ? pv[0] : (e == 1 ? pv[1] : (e == 2 ? pv[2] : pv[3]));
      const float an = e == 0 ? pn[0] : (e == 1 ? pn[1] : (e == 2 ? pn[2] : pn[3]));
      const float bu = e2 == 0 ? pu[0] : (e2 == 1 ? pu[1] : (e2 == 2 ? pu[2] : pu[3]));
      const float bv = e2 == 0 ? pv[0] : (e2 == 1 ? pv[1] : (e2 == 2 ? pv[2] : pv[3]));
      const float bn = e2 == 0 ? pn[0] : (e2 == 1 ? pn[1] : (e2 == 2 ? pn[2] : pn[3]));
      const bool onu = l < 2;                                   // clip line u = +-hu, else v = +-hv
      const float lim = (l & 1) ? -(onu ? hu : hv) : (onu ? hu : hv);
      const float da = (onu ? au : av) - lim, db = (onu ? bu : bv) - lim;
      if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
        const float tt = da / (da - db);
        cu = au + tt * (bu - au); cv = av + tt * (bv - av); cn = an + tt * (bn - an);
        if (onu) { cu = lim; valid = fabsf(cv) <= hv; } else { cv = lim; valid = fabsf(cu) <= hu; }
      }
    } else if (k < 24) {
      const int q = k - 20;
      cu = (q == 0 || q == 3) ? hu : -hu;
      cv = (q < 2) ? hv : -hv;
      // inside the incident quad (strictly: boundary cases are produced by the edge candidates)
      bool pos = true, neg = true;
#pragma unroll
      for (int e = 0; e < 4; e++) {
        const int e2 = (e + 1) & 3;
        const float cr = (pu[e2] - pu[e]) * (cv - pv[e]) - (pv[e2] - pv[e]) * (cu - pu[e]);
        pos = pos && cr > 0; neg = neg && cr < 0;
      }
      if (pos || neg) {
        // height of the incident plane above (cu, cv)
        const float e1u = pu[1] - pu[0], e1v = pv[1] - pv[0], e1n = pn[1] - pn[0];
        const float e2u = pu[3] - pu[0], e2v = pv[3] - pv[0], e2n = pn[3] - pn[0];
        const float Nu = e1v * e2n - e1n * e2v, Nv = e1n * e2u - e1u * e2n, Nn = e1u * e2v - e1v * e2u;
        cn = pn[0] - (Nu * (cu - pu[0]) + Nv * (cv - pv[0])) / Nn;
        valid = fabsf(Nn) > 1e-20f;
      }
    }
    return valid;
  };
  // one candidate per lane and pass: a 32-lane tile needs one pass, a 16-lane tile two (candidates 0..15, then 16..23)
  constexpr int PASSES = (24 + (int)LPE - 1) / (int)LPE;
  const V3 nout = refA ? nref : -nref;                          // always geom1 -> geom2
  bool valid[PASSES];
  float depth[PASSES];
  V3 p[PASSES];
#pragma unroll
  for (int ps = 0; ps < PASSES; ps++) {
    float cu, cv, cn;
    const int k = lane + ps * (int)LPE;
    valid[ps] = k < 24 && candidate(k, cu, cv, cn);
    depth[ps] = hn - cn;
    valid[ps] = valid[ps] && depth[ps] > 0;
    p[ps] = Rf.c + axu * cu + axv * cv + nref * (cn + 0.5f * depth[ps]);
  }
  if (single) {
    // mjc_Convex semantics (one contact per pair): deepest feature, centroid if it is not a single vertex
    float dmax = -1.0f;
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++) dmax = fmaxf(dmax, valid[ps] ? depth[ps] : -1.0f);
#pragma unroll
    for (int off = LPE / 2; off > 0; off >>= 1) dmax = fmaxf(dmax, t.shfl_xor(dmax, off));
    if (dmax <= 0) return 0;
    float sx = 0, sy = 0, sz = 0, cnt = 0;
#pragma unroll
    for (int ps = 0; ps < PASSES; ps++) {
      const bool deep = valid[ps] && depth[ps] >= dmax - 1e-6f;
      sx += deep ? p[ps].x : 0.0f; sy += deep ? p[ps].y : 0.0f; sz += deep ? p[ps].z : 0.0f; cnt += deep ? 1.0f : 0.0f;
    }
    tsum2(t, sx, sy);
    tsum2(t, sz, cnt);
    if (lane == 0 && base < NC) {
      const float ic = 1.0f / cnt;
      put_contact(con, base, mk(sx * ic, sy * ic, sz * ic), nout, -dmax, pair);
    }
    return 1;
  }
  int total = 0;
#pragma unroll
  for (int ps = 0; ps < PASSES; ps++) {
    const unsigned m = t.ballot(valid[ps]);
    const int slot = total + __popc(m & ((1u << lane) - 1u));
    if (valid[ps] && slot < 8 && base + slot < NC) put_contact(con, base + slot, p[ps], nout, -depth[ps], pair);
    total += __popc(m);
  }
  return min(total, 8);
}

// Stage A for one env.  Writes the contact list and the header of workspace record `w`; returns (on every lane)
// the number of hull pairs that need GJK/EPA, with *ncon_out the contact count so far (NC + 1 = list full) and *coupled_out
// whether one of them joins an arm link and the cube (dense Hessian: medium / heavy solve kernel).
//   1. world bounds of the 25 collidable geoms (one lane each): centre + bounding radius, AABB half extents;
//   2. broad phase over the static 191-pair table (one lane per pair): sphere-sphere and AABB-AABB; any
//      conservative filter gives the same contacts, because every survivor goes through the exact test of 3;
//   3. 15-axis separating-axis test on the oriented boxes, one lane per surviving pair (box pairs and hull
//      pairs in the same rounds): penetrating box pairs -> q1 with their axis code, hull pairs whose boxes
//      overlap -> the GJK/EPA list in the workspace header;
//   4. contact points, whole tile per penetrating box pair.
template <unsigned LPE> __device__ int collide_box_env(const Tile<LPE>& t, BoxS* S, float* w, const DevTables& T, int* ncon_out, bool* coupled_out) {
  const int lane = t.thread_rank();
  const unsigned lt = (1u << lane) - 1u;
  float* con = w + W_CON;
  for (int g = lane; g < c_m.ngeom; g += LPE) {
    const DevGeom& G = T.geom[g];
    const V3 c = geom_center(S->f, G);
    const float* m = G.link >= 0 ? S->f.lmat[G.link] : G.wmat;
    S->gbox[g] = make_float4(c.x, c.y, c.z, G.rbound);
    S->gext[g] = make_float4(fabsf(m[0]) * G.half[0] + fabsf(m[1]) * G.half[1] + fabsf(m[2]) * G.half[2],
                             fabsf(m[3]) * G.half[0] + fabsf(m[4]) * G.half[1] + fabsf(m[5]) * G.half[2],
                             fabsf(m[6]) * G.half[0] + fabsf(m[7]) * G.half[1] + fabsf(m[8]) * G.half[2], 0.0f);
  }
  t.sync();
  int ncand = 0;
  for (int base = 0; base < c_m.npair; base += LPE) {
    const int p = base + lane;
    bool pass = false;
    if (p < c_m.npair) {
      const uchar4 bp = T.bpair[p];
      const float4 a = S->gbox[bp.x], b = S->gbox[bp.y], ea = S->gext[bp.x], eb = S->gext[bp.y];
      const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z, rr = a.w + b.w;
      pass = fmaf(dx, dx, fmaf(dy, dy, dz * dz)) <= rr * rr && fabsf(dx) <= ea.x + eb.x && fabsf(dy) <= ea.y + eb.y &&
             fabsf(dz) <= ea.z + eb.z;
    }
    const unsigned m = t.ballot(pass);
    if (pass) S->qc[ncand + __popc(m & lt)] = (unsigned char)p;
    ncand += __popc(m);
  }
  t.sync();
  int npen = 0, nsurv = 0, nbox = 0;
  for (int base = 0; base < ncand; base += LPE) {
    const int k = base + lane;
    bool hit = false, hull = false;
    int p = 0, code = 0;
    float sep = 0;
    if (k < ncand) {
      p = S->qc[k];
      const uchar4 bp = T.bpair[p];
      hull = (bp.z & 0x7f) == MODE_HULL;
      const float4 c1 = S->gbox[bp.x], c2 = S->gbox[bp.y];
      Obb A, B;
      load_obb(S->f, T.geom[bp.x], mk(c1.x, c1.y, c1.z), A);
      load_obb(S->f, T.geom[bp.y], mk(c2.x, c2.y, c2.z), B);
      hit = box_sat(A, B, hull, code, sep);
    }
    const unsigned mb = t.ballot(hit && !hull), mh = t.ballot(hit && hull);
    nbox += __popc(t.ballot(k < ncand && !hull));
    if (hit && !hull) {
      const int slot = npen + __popc(mb & lt);
      if (slot < NPEN) {
        S->q1[slot] = (unsigned char)p;
        S->qcode[slot] = (unsigned char)code;
        S->qsep[slot] = sep;
      }
    }
    if (hit && hull) {
      const int slot = nsurv + __popc(mh & lt);
      if (slot < NHP) reinterpret_cast<unsigned char*>(w + W_HULLP)[slot] = (unsigned char)p;
    }
    npen += __popc(mb);
    nsurv += __popc(mh);
  }
  t.sync();
  int ncon = 0, coupled = 0;
  for (int k = 0; k < min(npen, NPEN); k++) {
    const int p = S->q1[k];
    const uchar4 bp = T.bpair[p];
    const float4 c1 = S->gbox[bp.x], c2 = S->gbox[bp.y];
    Obb A, B;
    load_obb(S->f, T.geom[bp.x], mk(c1.x, c1.y, c1.z), A);
    load_obb(S->f, T.geom[bp.y], mk(c2.x, c2.y, c2.z), B);
    const int nc = box_contacts(t, con, ncon, A, B, (int)S->qcode[k], S->qsep[k], (bp.z & 0x7f) == MODE_BOX_SINGLE, p);
    if (nc > 0 && (bp.z & PAIR_COUPLES)) coupled = HDR_COUPLED;
    ncon = min(ncon + nc, NC + 1);   // NC + 1: the list is full (the NC records written are valid)
  }
  // more penetrating box pairs / hull pairs than the lists hold: the contact count stays the number of valid records; the
  // overflow is flagged in the header and counted by the task kernel (diagnostics[0])
  int overflow = 0;
  if (npen > NPEN || nsurv > NHP) { overflow = HDR_OVERFLOW; nsurv = min(nsurv, NHP); }
  if (lane == 0)
    *reinterpret_cast<int4*>(w + W_HDR) = make_int4(ncon, nsurv, min(nbox, 255) | (min(npen, 255) << 8) | (min(ncand - nbox, 255) << 16) | coupled | overflow, nsurv);
  *ncon_out = ncon;
  *coupled_out = coupled != 0;
  return nsurv;
}

}  // namespace so100
