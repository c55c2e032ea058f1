// Low-resolution state-to-pixels renderer for obs_type "so100_pixels_agent_pos" (gym_so100/env.py:50-66, 130-136: the
// observation is the image of camera "top" plus the six joint angles; scene_so100.xml:28 places that camera at (0, 0.6, 0.8)
// looking at the table body, fovy 78 degrees).  One block per env: the first warp runs the forward kinematics of the state
// record, then every thread casts the rays of its pixels against the scene's collision geometry -- boxes by the slab test in
// their frame, convex hulls as the intersection of their facets' half-spaces (planes in the body frame, built on the host from
// the hull vertices) behind a bounding-sphere cull -- and shades the nearest hit with MuJoCo's fixed-function terms that matter
// at this resolution (headlight ambient + diffuse, the scene's three directional lights, no shadows, no specular).
// This is NOT MuJoCo's OpenGL image: the arm is drawn by its collision hulls (the visual meshes are the same STL parts without
// the convex-hull closure, plus the motors), there is no anti-aliasing and no shadow map.  DESIGN.md section 11.
#pragma once
#include "so100_task.cuh"

namespace so100 {

struct RenderCfg {
  int width, height;
  float cam_pos[3], cam_x[3], cam_y[3], cam_z[3];   // camera frame in the world: looks along -z, x right, y up
  float tan_half_fovy;
  float ambient, head_diffuse;
  int nlight;
  float light_dir[4][3], light_diffuse[4];         // directional lights (unit direction the light travels)
  float background[3];
};

struct RenderTables {
  const float4* planes;     // [P] body-frame facets: n . x <= w
  const int* adr;           // [NGEOM] first plane of a hull
  const int* num;           // [NGEOM] > 0: hull with that many planes; 0: box (centre / half sizes of DevGeom); < 0: not drawn
  const float* rgb;         // [NGEOM, 3]
};

__global__ void __launch_bounds__(128) render_kernel(const float* state, int n, RenderCfg C, DevTables T, RenderTables R, uint8_t* pixels) {
  __shared__ __align__(16) TaskS S;
  __shared__ float gR[NGEOM][9], gp[NGEOM][3], gc[NGEOM][4];      // frame axes, frame origin, world centre + bounding radius
  cg::thread_block blk = cg::this_thread_block();
  Tile<32> t = cg::tiled_partition<32>(blk);
  const int env = blockIdx.x;
  if (env >= n) return;
  if (threadIdx.x < 32) {
    copy_vec<32, STATE_WORDS>(t, S.st, state + (size_t)env * STATE_WORDS);
    t.sync();
    kinematics<false>(t, &S);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < c_m.ngeom; g += blockDim.x) {
    const DevGeom& G = T.geom[g];
    const float* m = G.link >= 0 ? S.f.lmat[G.link] : G.wmat;
    for (int k = 0; k < 9; k++) gR[g][k] = m[k];
    V3 org, cen;
    if (G.link >= 0) {
      org = ld3(S.f.lpos[G.link]);
      cen = org + mulmv(m, ld3(G.center));
      if (R.num[g] == 0) org = cen;                  // a box is intersected in the frame centred on it
    } else {
      cen = ld3(G.center);                           // static geoms: centre already in the world
      org = R.num[g] == 0 ? cen : ld3(G.org);
    }
    st3(gp[g], org);
    gc[g][0] = cen.x; gc[g][1] = cen.y; gc[g][2] = cen.z; gc[g][3] = G.rbound;
  }
  __syncthreads();
  const V3 o = ld3(C.cam_pos), cx = ld3(C.cam_x), cy = ld3(C.cam_y), cz = ld3(C.cam_z);
  const float aspect = (float)C.width / (float)C.height;
  const int npix = C.width * C.height;
  for (int pix = threadIdx.x; pix < npix; pix += blockDim.x) {
    const int r = pix / C.width, c = pix - r * C.width;
    const float px = C.tan_half_fovy * aspect * (2.0f * (c + 0.5f) / C.width - 1.0f);
    const float py = C.tan_half_fovy * (1.0f - 2.0f * (r + 0.5f) / C.height);
    const V3 d = normalized(cx * px + cy * py - cz);
    float best = 3.0e38f;
    int hit = -1;
    V3 nrm = mk(0, 0, 1);
    for (int g = 0; g < c_m.ngeom; g++) {
      const int np = R.num[g];
      if (np < 0) continue;
      // bounding sphere
      const V3 oc = mk(gc[g][0], gc[g][1], gc[g][2]) - o;
      const float b = dot(oc, d), q = dot(oc, oc) - b * b;
      if (q > gc[g][3] * gc[g][3] || b + gc[g][3] < 0.0f || b - gc[g][3] > best) continue;
      const V3 ol = mulmtv(gR[g], o - ld3(gp[g])), dl = mulmtv(gR[g], d);
      float tn = 0.0f, tf = best;
      V3 nl = mk(0, 0, 0);
      bool miss = false;
      if (np == 0) {
        const DevGeom& G = T.geom[g];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float ok = comp(ol, k), dk = comp(dl, k), h = G.half[k];
          if (fabsf(dk) < 1e-12f) { if (fabsf(ok) > h) miss = true; continue; }
          const float inv = 1.0f / dk;
          float t1 = (-h - ok) * inv, t2 = (h - ok) * inv;
          const float sgn = t1 <= t2 ? -1.0f : 1.0f;         // entering through the -h face when moving along +k
          if (t1 > t2) { const float tmp = t1; t1 = t2; t2 = tmp; }
          if (t1 > tn) { tn = t1; nl = mk(k == 0 ? sgn : 0.0f, k == 1 ? sgn : 0.0f, k == 2 ? sgn : 0.0f); }
          tf = fminf(tf, t2);
        }
      } else {
        const float4* P = R.planes + R.adr[g];
        for (int k = 0; k < np; k++) {
          const float4 pl = __ldg(&P[k]);
          const float den = pl.x * dl.x + pl.y * dl.y + pl.z * dl.z;
          const float dist = pl.w - (pl.x * ol.x + pl.y * ol.y + pl.z * ol.z);
          if (den < 0.0f) {
            const float tt = dist / den;
            if (tt > tn) { tn = tt; nl = mk(pl.x, pl.y, pl.z); }
          } else if (den > 0.0f) {
            tf = fminf(tf, dist / den);
          } else if (dist < 0.0f) {
            miss = true;
          }
          if (tn > tf) break;
        }
      }
      if (miss || tn > tf || tn <= 0.0f || tn >= best) continue;
      best = tn; hit = g;
      nrm = mulmv(gR[g], nl);
    }
    float rgb[3] = {C.background[0], C.background[1], C.background[2]};
    if (hit >= 0) {
      float inten = C.ambient + C.head_diffuse * fmaxf(0.0f, -dot(nrm, d));
      for (int l = 0; l < C.nlight; l++) inten += C.light_diffuse[l] * fmaxf(0.0f, -dot(nrm, ld3(C.light_dir[l])));
      for (int k = 0; k < 3; k++) rgb[k] = fminf(1.0f, R.rgb[hit * 3 + k] * inten);
    }
    uint8_t* out = pixels + ((size_t)env * npix + pix) * 3;
    for (int k = 0; k < 3; k++) out[k] = (uint8_t)(rgb[k] * 255.0f + 0.5f);
  }
}

}  // namespace so100
