// Collision, stage B: GJK + EPA for the pairs that involve a general convex hull (arm links, jaw
// hulls, base), one tile per queued env: the support map over the hull vertices is lane-parallel
// (float4 loads + shuffle arg-max), the simplex / polytope logic is tile-uniform, the EPA polytope
// lives in shared memory.  One contact per pair (MuJoCo mjc_Convex with multiccd off): normal and
// depth are the minimum-translation solution, the point is the midpoint of the EPA witness points
// (or the deepest-feature centroid when the witness is not unique).
#pragma once
#include "so100_box.cuh"

namespace so100 {

struct Shape {
  int boxlike;
  V3 base;            // boxlike: world centre; hull: world origin of the vertex frame
  const float* mat;   // row-major axes (shared-memory link frame or static table)
  float h[3];
  int vadr, vnum;
};

__device__ __forceinline__ void load_shape(const FrameBlock& f, const DevGeom& G, V3 center, Shape& s) {
  s.boxlike = G.boxlike;
  s.mat = G.link >= 0 ? f.lmat[G.link] : G.wmat;
  s.h[0] = G.half[0]; s.h[1] = G.half[1]; s.h[2] = G.half[2];
  s.vadr = G.vadr; s.vnum = G.vnum;
  if (G.boxlike) s.base = center;
  else s.base = G.link >= 0 ? ld3(f.lpos[G.link]) : ld3(G.org);
}

// order-preserving map float -> uint (for integer warp reductions); -0 and +0 map to the same key
__device__ __forceinline__ unsigned ordered_key(float f) {
  const unsigned b = __float_as_uint(f + 0.0f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
// index of the largest value over the tile, ties to the lowest index (every lane passes its best value / index):
// two redux.sync instructions on a full warp instead of a 5-round shuffle arg-max
template <unsigned LPE> __device__ __forceinline__ int tile_argmax(const Tile<LPE>& t, float best, int bi) {
  if constexpr (LPE == 32) {
    const unsigned key = ordered_key(best);
    const unsigned kmax = __reduce_max_sync(0xffffffffu, key);
    return (int)__reduce_min_sync(0xffffffffu, key == kmax ? (unsigned)bi : 0x7fffffffu);
  } else {
#pragma unroll
    for (int off = LPE / 2; off > 0; off >>= 1) {
      const float ov = t.shfl_xor(best, off);
      const int oi = t.shfl_xor(bi, off);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    return bi;
  }
}

// support point in world direction d; identical on every lane of the tile
template <unsigned LPE>
__device__ __forceinline__ V3 support(const Tile<LPE>& t, const Shape& s, V3 d, const float4* __restrict__ vert) {
  const V3 dl = mulmtv(s.mat, d);
  V3 pl;
  if (s.boxlike) {
    pl = mk(dl.x >= 0 ? s.h[0] : -s.h[0], dl.y >= 0 ? s.h[1] : -s.h[1], dl.z >= 0 ? s.h[2] : -s.h[2]);
  } else {
    float best = -3.0e38f;
    int bi = 0x7fffffff;
    for (int v = t.thread_rank(); v < s.vnum; v += LPE) {
      const float4 p = __ldg(&vert[s.vadr + v]);
      const float val = fmaf(p.x, dl.x, fmaf(p.y, dl.y, p.z * dl.z));
      if (val > best) { best = val; bi = v; }
    }
    bi = tile_argmax(t, best, bi);
    const float4 p = __ldg(&vert[s.vadr + bi]);
    pl = mk(p.x, p.y, p.z);
  }
  return s.base + mulmv(s.mat, pl);
}

struct MV { V3 w, a; };   // Minkowski-difference vertex w = a - b (b is recovered as a - w)

template <unsigned LPE>
__device__ __forceinline__ MV msupport(const Tile<LPE>& t, const Shape& A, const Shape& B, V3 d, const float4* vert) {
  MV m;
  m.a = support(t, A, d, vert);
  const V3 b = support(t, B, -d, vert);
  m.w = m.a - b;
  return m;
}

// GJK simplex in named registers: a is always the newest point, then b, c, d (n of them are live).  An indexed array
// `MV s[4]` with a run-time count lives in local memory (150 LDL/STL in the hot loop of the queue kernel); the moves below
// are the same permutations the array version made, so every dot / cross product sees the same operands.
struct Simplex { MV a, b, c, d; int n; };

// triangle case of the simplex update on (a, b, c)
__device__ __forceinline__ void simplex_triangle(Simplex& s, V3& dir) {
  const MV A = s.a, B = s.b, C = s.c;
  const V3 ao = -A.w, ab = B.w - A.w, ac = C.w - A.w, abc = cross(ab, ac);
  bool edge_ab = false;
  if (dot(cross(abc, ac), ao) > 0) {
    if (dot(ac, ao) > 0) { s.b = C; s.n = 2; dir = cross(cross(ac, ao), ac); return; }     // {a, c}
    edge_ab = true;
  } else if (dot(cross(ab, abc), ao) > 0) {
    edge_ab = true;
  }
  if (edge_ab) {
    if (dot(ab, ao) > 0) { s.n = 2; dir = cross(cross(ab, ao), ab); }                        // {a, b}
    else { s.n = 1; dir = ao; }                                                              // {a}
    return;
  }
  if (dot(abc, ao) > 0) dir = abc;
  else { s.b = C; s.c = B; dir = -abc; }                                                     // flip the winding
}

// simplex update after a point was pushed; returns true when the tetrahedron encloses the origin
__device__ __forceinline__ bool do_simplex(Simplex& s, V3& dir) {
  if (s.n == 2) {
    const V3 ao = -s.a.w, ab = s.b.w - s.a.w;
    if (dot(ab, ao) > 0) dir = cross(cross(ab, ao), ab);
    else { s.n = 1; dir = ao; }
    return false;
  }
  if (s.n == 3) { simplex_triangle(s, dir); return false; }
  const MV A = s.a, B = s.b, C = s.c, D = s.d;
  const V3 ao = -A.w, ab = B.w - A.w, ac = C.w - A.w, ad = D.w - A.w;
  V3 abc = cross(ab, ac), acd = cross(ac, ad), adb = cross(ad, ab);
  if (dot(abc, ad) > 0) abc = -abc;
  if (dot(acd, ab) > 0) acd = -acd;
  if (dot(adb, ac) > 0) adb = -adb;
  if (dot(abc, ao) > 0) { }                                   // (a, b, c)
  else if (dot(acd, ao) > 0) { s.b = C; s.c = D; }            // (a, c, d)
  else if (dot(adb, ao) > 0) { s.b = D; s.c = B; }            // (a, d, b)
  else return true;
  s.n = 3;
  simplex_triangle(s, dir);
  return false;
}

// closest point of triangle (a,b,c) to the origin: barycentric weights, returns squared distance
__device__ inline float tri_closest(V3 a, V3 b, V3 c, float* lam) {
  const V3 ab = b - a, ac = c - a;
  const float d1 = -dot(ab, a), d2 = -dot(ac, a);
  if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = 0; lam[2] = 0; return dot(a, a); }
  const float d3 = -dot(ab, b), d4 = -dot(ac, b);
  if (d3 >= 0 && d4 <= d3) { lam[0] = 0; lam[1] = 1; lam[2] = 0; return dot(b, b); }
  const float vc = d1 * d4 - d3 * d2;
  if (vc <= 0 && d1 >= 0 && d3 <= 0) {
    const float v = d1 / (d1 - d3);
    lam[0] = 1 - v; lam[1] = v; lam[2] = 0;
  } else {
    const float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0 && d5 <= d6) { lam[0] = 0; lam[1] = 0; lam[2] = 1; return dot(c, c); }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) {
      const float w = d2 / (d2 - d6);
      lam[0] = 1 - w; lam[1] = 0; lam[2] = w;
    } else {
      const float va = d3 * d6 - d5 * d4;
      if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
        const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        lam[0] = 0; lam[1] = 1 - w; lam[2] = w;
      } else {
        const float den = 1.0f / (va + vb + vc);
        lam[1] = vb * den; lam[2] = vc * den; lam[0] = 1 - lam[1] - lam[2];
      }
    }
  }
  const V3 p = a * lam[0] + b * lam[1] + c * lam[2];
  return dot(p, p);
}

// EPA polytope in shared memory (HullS::epa)
constexpr int EPA_MAXV = 32;
constexpr int EPA_MAXF = 64;
struct EpaScratch {
  float vw[EPA_MAXV][3], va[EPA_MAXV][3];
  float fn[EPA_MAXF][3], fdp[EPA_MAXF], fdt[EPA_MAXF];
  unsigned char fv[EPA_MAXF][4];        // vertex ids, [3] = alive flag
  unsigned char vis[EPA_MAXF];
  unsigned char owner[EPA_MAXV][EPA_MAXV];  // face owning the directed edge a -> b
  unsigned char hedge[EPA_MAXF][2];
  unsigned char freed[EPA_MAXF];
#ifdef SO100_HULL_CLOCK
  int dbg[2];
  int ph[6];     // cycles: closest face, support, visibility, horizon, new faces, (spare)
#endif
};
#ifdef SO100_HULL_CLOCK
#define HULL_PH(k) { const long long c_ = clock64(); if (lane == 0) E->ph[k] += (int)(c_ - phc_); phc_ = c_; }
#else
#define HULL_PH(k)
#endif

__device__ __forceinline__ void epa_make_face(EpaScratch* E, int slot, int i, int j, int k) {
  const V3 a = ld3(E->vw[i]), b = ld3(E->vw[j]), c = ld3(E->vw[k]);
  V3 n = cross(b - a, c - a);
  const float len2 = dot(n, n);
  float lam[3];
  E->fv[slot][0] = (unsigned char)i; E->fv[slot][1] = (unsigned char)j; E->fv[slot][2] = (unsigned char)k;
  E->owner[i][j] = (unsigned char)slot; E->owner[j][k] = (unsigned char)slot; E->owner[k][i] = (unsigned char)slot;
  if (len2 < 1e-30f) { E->fv[slot][3] = 0; E->fdt[slot] = 3.0e38f; E->fdp[slot] = 0; st3(E->fn[slot], mk(0, 0, 1)); return; }
  n = n * rsqrtf(len2);
  st3(E->fn[slot], n);
  E->fdp[slot] = dot(n, a);
  E->fdt[slot] = sqrtf(tri_closest(a, b, c, lam));
  E->fv[slot][3] = 1;
}

#ifdef SO100_HULL_CLOCK
// development build: per GJK/EPA item (ns, GJK iterations, EPA iterations, hull vertices), ring buffer of 65536 items
__device__ int g_hull_stat[65536][4];
__device__ int g_hull_phase[65536][6];
__device__ int g_hull_stat_n;
#endif

// Penetration of A into B.  On a hit: normal (A -> B), depth > 0, contact point (midpoint of the
// witness points).  All lanes return the same values.
// max_gjk / max_epa: iteration budgets of the regular (budgeted) GJK/EPA kernel; when one runs out *over_budget is set and the
// result is void (the env goes to the slow lane, which calls this without budgets).
template <unsigned LPE>
__device__ bool gjk_epa(const Tile<LPE>& t, const Shape& A, const Shape& B, V3 ca, V3 cb, const float4* vert,
                        EpaScratch* E, V3& normal, float& depth, V3& pos, int max_gjk = 48, int max_epa = EPA_MAXV - 4,
                        bool* over_budget = nullptr) {
  const int lane = t.thread_rank();
  Simplex s;
  V3 dir = cb - ca;
  if (dot(dir, dir) < 1e-20f) dir = mk(1, 0, 0);
  s.a = msupport(t, A, B, dir, vert);
  s.b = s.a; s.c = s.a; s.d = s.a;
  s.n = 1;
  dir = -s.a.w;
  bool hit = false;
  int gjk_its_ = 0;
  for (int it = 0; it < 48; it++) {
    if (it >= max_gjk) { *over_budget = true; return false; }
    gjk_its_ = it + 1;
    if (dot(dir, dir) < 1e-24f) break;
    const MV w = msupport(t, A, B, dir, vert);
    if (dot(w.w, dir) <= 0) break;
    s.d = s.c; s.c = s.b; s.b = s.a; s.a = w; s.n++;
    if (do_simplex(s, dir)) { hit = true; break; }
  }
#ifdef SO100_HULL_CLOCK
  E->dbg[0] = gjk_its_; E->dbg[1] = 0;
  if (lane < 6) E->ph[lane] = 0;
#endif
  if (!hit) return false;
  // ---- EPA
  t.sync();
  if (lane < 4) {
    // polytope vertices 0..3 = oldest .. newest simplex point (d, c, b, a)
    const MV m = lane == 0 ? s.d : (lane == 1 ? s.c : (lane == 2 ? s.b : s.a));
    st3(E->vw[lane], m.w); st3(E->va[lane], m.a);
  }
  t.sync();
  if (lane < 4) {
    const int tf[4][3] = {{0, 1, 2}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
    int a = tf[lane][0], b = tf[lane][1], c = tf[lane][2];
    const int opp = 6 - a - b - c;
    const V3 va = ld3(E->vw[a]);
    const V3 nn = cross(ld3(E->vw[b]) - va, ld3(E->vw[c]) - va);
    if (dot(nn, ld3(E->vw[opp]) - va) > 0) { const int tmp = b; b = c; c = tmp; }
    epa_make_face(E, lane, a, b, c);
  }
  t.sync();
  int nface = 4, nv = 4, best = 0;
  bool degenerate = false;
  for (int k = 0; k < 4; k++) degenerate |= (E->fv[k][3] == 0);
  if (degenerate) return false;   // flat tetrahedron: touching
  for (int it = 0; it < EPA_MAXV - 4; it++) {
    if (it >= max_epa) { *over_budget = true; t.sync(); return false; }
#ifdef SO100_HULL_CLOCK
    if (lane == 0) E->dbg[1] = it + 1;
    long long phc_ = clock64();
#endif
    // closest face (ties -> lowest slot)
    float bd = 3.0e38f; int bf = 0x7fffffff;
    for (int f = lane; f < nface; f += LPE)
      if (E->fv[f][3] && E->fdt[f] < bd) { bd = E->fdt[f]; bf = f; }
    bf = tile_argmax(t, -bd, bf);
    if (bf == 0x7fffffff) break;
    best = bf;
    const V3 nb = ld3(E->fn[best]);
    HULL_PH(0)
    const MV w = msupport(t, A, B, nb, vert);
    HULL_PH(1)
    if (dot(w.w, nb) - E->fdp[best] < 1e-6f || nface + 2 > EPA_MAXF) break;
    // visibility
    int myvis = 0;
    for (int f = lane; f < nface; f += LPE) {
      const unsigned char v = (E->fv[f][3] && dot(ld3(E->fn[f]), w.w - ld3(E->vw[E->fv[f][0]])) > 0.0f) ? 1 : 0;
      E->vis[f] = v;
      myvis += v;
    }
    t.sync();
    HULL_PH(2)
    if (!E->vis[best]) break;      // round-off: the expanding face does not see the new point
    // freed slots (compaction of visible faces) and horizon edges
    int nfree = 0, nh = 0;
    for (int base = 0; base < nface; base += LPE) {
      const int f = base + lane;
      const bool v = f < nface && E->vis[f];
      const unsigned m = t.ballot(v);
      if (v) E->freed[nfree + __popc(m & ((1u << lane) - 1u))] = (unsigned char)f;
      nfree += __popc(m);
    }
    for (int base = 0; base < nface * 3; base += LPE) {
      const int idx = base + lane, f = idx / 3, e = idx - f * 3;
      bool h = false;
      int ea = 0, eb = 0;
      if (f < nface && E->vis[f]) {
        ea = E->fv[f][e]; eb = E->fv[f][e == 2 ? 0 : e + 1];
        h = !E->vis[E->owner[eb][ea]];
      }
      const unsigned m = t.ballot(h);
      if (h) {
        const int k = nh + __popc(m & ((1u << lane) - 1u));
        if (k < EPA_MAXF) { E->hedge[k][0] = (unsigned char)ea; E->hedge[k][1] = (unsigned char)eb; }
      }
      nh += __popc(m);
    }
    t.sync();
    HULL_PH(3)
    if (nh < 3 || nh > EPA_MAXF || nface + (nh - nfree) > EPA_MAXF) break;
    if (lane == 0) { st3(E->vw[nv], w.w); st3(E->va[nv], w.a); }
    for (int f = lane; f < nface; f += LPE) if (E->vis[f]) E->fv[f][3] = 0;
    t.sync();
    for (int k = lane; k < nh; k += LPE) {
      const int slot = k < nfree ? E->freed[k] : nface + (k - nfree);
      epa_make_face(E, slot, E->hedge[k][0], E->hedge[k][1], nv);
    }
    nface += max(nh - nfree, 0);
    nv++;
    t.sync();
    HULL_PH(4)
  }
  const V3 a0 = ld3(E->vw[E->fv[best][0]]), a1 = ld3(E->vw[E->fv[best][1]]), a2 = ld3(E->vw[E->fv[best][2]]);
  float lam[3];
  tri_closest(a0, a1, a2, lam);
  const V3 pa = ld3(E->va[E->fv[best][0]]) * lam[0] + ld3(E->va[E->fv[best][1]]) * lam[1] + ld3(E->va[E->fv[best][2]]) * lam[2];
  const V3 pw = a0 * lam[0] + a1 * lam[1] + a2 * lam[2];
  normal = ld3(E->fn[best]);
  depth = E->fdp[best];
  pos = pa - pw * 0.5f;
  t.sync();
  return depth > 0;
}

static_assert(sizeof(EpaScratch) <= sizeof(HullS::epa), "EPA scratch does not fit its shared-memory slot");

// vertices of a shape within `tol` of its support plane in world direction d: count, centroid (world), radius and
// the world vector from the centroid to the farthest member (the edge direction when count == 2)
template <unsigned LPE>
__device__ int support_set(const Tile<LPE>& t, const Shape& s, V3 d, float tol, const float4* __restrict__ vert, V3& centroid,
                           float& radius, V3& far) {
  const V3 dl = mulmtv(s.mat, d);
  const int n = s.boxlike ? 8 : s.vnum;
  auto vertex = [&](int i) {
    if (s.boxlike) return mk((i & 1) ? s.h[0] : -s.h[0], (i & 2) ? s.h[1] : -s.h[1], (i & 4) ? s.h[2] : -s.h[2]);
    const float4 p = __ldg(&vert[s.vadr + i]);
    return mk(p.x, p.y, p.z);
  };
  float best = -3.0e38f;
  for (int i = t.thread_rank(); i < n; i += LPE) best = fmaxf(best, dot(vertex(i), dl));
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) best = fmaxf(best, t.shfl_xor(best, off));
  float sx = 0, sy = 0, sz = 0, cnt = 0;
  for (int i = t.thread_rank(); i < n; i += LPE) {
    const V3 v = vertex(i);
    if (dot(v, dl) >= best - tol) { sx += v.x; sy += v.y; sz += v.z; cnt += 1.0f; }
  }
#pragma unroll
  for (int off = LPE / 2; off > 0; off >>= 1) {
    sx += t.shfl_xor(sx, off); sy += t.shfl_xor(sy, off); sz += t.shfl_xor(sz, off); cnt += t.shfl_xor(cnt, off);
  }
  const float ic = 1.0f / cnt;
  const V3 c = mk(sx * ic, sy * ic, sz * ic);
  float r2 = 0.0f;
  V3 fl = mk(0, 0, 0);
  if (cnt > 1.5f) {          // a single member is its own centroid: radius 0, no third sweep
    r2 = -1.0f;
    int bi = 0x7fffffff;
    for (int i = t.thread_rank(); i < n; i += LPE) {
      const V3 v = vertex(i);
      if (dot(v, dl) >= best - tol) {
        const V3 e = v - c;
        const float q = dot(e, e);
        if (q > r2) { r2 = q; bi = i; fl = e; }
      }
    }
#pragma unroll
    for (int off = LPE / 2; off > 0; off >>= 1) {
      const float oq = t.shfl_xor(r2, off);
      const int oi = t.shfl_xor(bi, off);
      const V3 of = mk(t.shfl_xor(fl.x, off), t.shfl_xor(fl.y, off), t.shfl_xor(fl.z, off));
      if (oq > r2 || (oq == r2 && oi < bi)) { r2 = oq; bi = oi; fl = of; }
    }
  }
  radius = sqrtf(fmaxf(r2, 0.0f));
  centroid = s.base + mulmv(s.mat, c);
  far = mulmv(s.mat, fl);
  return (int)(cnt + 0.5f);
}

// A normal within 1e-3 rad of a face normal of a box-like geom in the pair is snapped onto it and the
// penetration is re-measured along the snapped direction (same rule as the oracle).
template <unsigned LPE>
__device__ void snap_normal(const Tile<LPE>& t, const Shape& A, const Shape& B, V3& n, float& depth, const float4* vert) {
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const Shape& X = s == 0 ? A : B;
    if (!X.boxlike) continue;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const V3 ax = mcol(X.mat, k);
      const float c = dot(n, ax);
      if (fabsf(c) > 1.0f - 5e-7f) {
        n = c > 0 ? ax : -ax;
        const V3 pa = support(t, A, n, vert), pb = support(t, B, -n, vert);
        depth = dot(n, pa) - dot(n, pb);
        return;
      }
    }
  }
}

// Contact point of a GJK/EPA hit: the EPA witness midpoint when it is unique (vertex-face, edge-edge),
// otherwise the centroid of the smaller deepest feature moved half the depth towards the other geom
// (same rule as the oracle: flat resting contacts stay torque-free and precision-independent).
template <unsigned LPE>
__device__ V3 deepest_feature_point(const Tile<LPE>& t, const Shape& A, const Shape& B, V3 n, float depth, V3 pos,
                                    const float4* vert) {
  V3 cA, cB, fA, fB;
  float rA, rB;
  const int nA = support_set(t, A, n, 1e-6f, vert, cA, rA, fA);
  if (nA == 1) return cA - n * (0.5f * depth);          // a single deepest vertex of A decides; B's feature is not needed
  const int nB = support_set(t, B, -n, 1e-6f, vert, cB, rB, fB);
  if (nA == 2 && nB == 2) {
    const V3 x = cross(fA, fB);
    if (dot(x, x) > 1e-6f * dot(fA, fA) * dot(fB, fB)) return pos;   // crossing edges: the EPA witness is unique
  }
  if (nB != 1 && rA <= rB) return cA - n * (0.5f * depth);
  return cB + n * (0.5f * depth);
}

// Stage B, one queue item = one hull pair of one env: GJK/EPA with the whole tile; the result goes to the pair's
// staging slot.  The tile that finishes an env's last pending pair merges the staged contacts into the contact list
// in pair order (deterministic contact order, after the box contacts) and returns the final count, with *coupled_out
// whether an arm-cube contact exists; every other tile returns -1.
// With budgets (max_gjk / max_epa < their caps) an item that exceeds them returns -2 and leaves the env's pending counter alone.
template <unsigned LPE> __device__ int collide_hull_item(const Tile<LPE>& t, HullS* S, float* w, int slot, const DevTables& T, bool* coupled_out,
                                                         int max_gjk = 48, int max_epa = EPA_MAXV - 4) {
  const int lane = t.thread_rank();
  const int p = reinterpret_cast<const unsigned char*>(w + W_HULLP)[slot];
  const DevPair& P = T.pair[p];
  const DevGeom& G1 = T.geom[P.g1];
  const DevGeom& G2 = T.geom[P.g2];
  const V3 c1 = geom_center(S->f, G1), c2 = geom_center(S->f, G2);
  Shape A, B;
  load_shape(S->f, G1, c1, A);
  load_shape(S->f, G2, c2, B);
  V3 n = mk(0, 0, 1), pos = mk(0, 0, 0);
  float depth = 0;
  int pid = -1;
#ifdef SO100_HULL_CLOCK
  unsigned long long hc0_, hc1_, hc2_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(hc0_));
  bool hit_ = gjk_epa(t, A, B, c1, c2, T.vert, reinterpret_cast<EpaScratch*>(S->epa), n, depth, pos);
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(hc1_));
  if (hit_) {
    snap_normal(t, A, B, n, depth, T.vert);
    pos = deepest_feature_point(t, A, B, n, depth, pos, T.vert);
    pid = p;
  }
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(hc2_));
  t.sync();
  if (lane == 0) {
    const int k_ = atomicAdd(&g_hull_stat_n, 1) & 65535;
    const EpaScratch* E_ = reinterpret_cast<const EpaScratch*>(S->epa);
    g_hull_stat[k_][0] = (int)(hc1_ - hc0_); g_hull_stat[k_][1] = (int)(hc2_ - hc1_);
    g_hull_stat[k_][2] = E_->dbg[0] | (E_->dbg[1] << 8) | ((int)hit_ << 16); g_hull_stat[k_][3] = A.vnum + B.vnum;
    for (int q_ = 0; q_ < 6; q_++) g_hull_phase[k_][q_] = hit_ ? E_->ph[q_] : 0;
  }
#else
  bool over = false;
  if (gjk_epa(t, A, B, c1, c2, T.vert, reinterpret_cast<EpaScratch*>(S->epa), n, depth, pos, max_gjk, max_epa, &over)) {
    snap_normal(t, A, B, n, depth, T.vert);
    pos = deepest_feature_point(t, A, B, n, depth, pos, T.vert);
    pid = p;
  }
  if (over) return -2;
#endif
  int* hdr = reinterpret_cast<int*>(w + W_HDR);
  const int nbox = hdr[0], nsurv = min(hdr[1], NHP);     // written by K2a (the previous kernel)
  if (nsurv > 1) {
    // several pairs, several tiles: staged results are published with a fence, the tile that takes the counter to zero merges
    int last = 0;
    if (lane == 0) {
      put_contact(w + W_HSTAGE, slot, pos, n, -depth, pid);
      __threadfence();
      last = atomicSub(&hdr[3], 1) == 1;
    }
    last = t.shfl(last, 0);
    if (!last) return -1;
    __threadfence();
  } else if (lane == 0) {
    // the env's only pair (the common case): nobody to synchronise with
    put_contact(w + W_HSTAGE, slot, pos, n, -depth, pid);
    hdr[3] = 0;
  }
  t.sync();
  float4 q0 = make_float4(0, 0, 0, 0), q1 = make_float4(0, 0, 0, __int_as_float(-1));
  if (lane < nsurv) {
    q0 = __ldcg(reinterpret_cast<const float4*>(w + W_HSTAGE + lane * CON_WORDS));
    q1 = __ldcg(reinterpret_cast<const float4*>(w + W_HSTAGE + lane * CON_WORDS) + 1);
  }
  const bool valid = lane < nsurv && __float_as_int(q1.w) >= 0;
  const unsigned m = t.ballot(valid);
  const int c = nbox + __popc(m & ((1u << lane) - 1u));
  if (valid && c < NC) {
    float4* dst = reinterpret_cast<float4*>(w + W_CON + c * CON_WORDS);
    dst[0] = q0; dst[1] = q1;
  }
  const int ncon = min(nbox + __popc(m), NC + 1);
  const bool couples = t.any(valid && (T.bpair[valid ? __float_as_int(q1.w) : 0].z & PAIR_COUPLES)) || (hdr[2] & HDR_COUPLED);
  if (lane == 0) {
    hdr[0] = ncon;
    if (couples) hdr[2] |= HDR_COUPLED;
  }
  *coupled_out = couples;
  return ncon;
}

}  // namespace so100
