// Device-side constant model + small math for the fused bin-a-cube step kernels (sm_100a).
//
// The packed so100_model (include/so100_model.h) is narrowed to float32 and re-arranged on
// the host (so100_b200.cu: build_dev_model) into:
//   * DevModel  -- uniform scalars / small tables, lives in __constant__ memory;
//   * DevGeom[] / DevPair[] / float4 verts[] -- per-lane indexed tables in global memory
//     (read-only, L1-resident: the whole set is < 64 KB).
// The kernel is specialised for the topology of gym_so100/assets/so100_transfer_cube.xml:
// a fixed base, a serial chain of NL = 6 hinge links (5 arm joints + the jaw) and one free
// body (the cube); the host refuses any model that does not have this shape.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace so100 {

constexpr int NL = 6;            // hinge links of the arm chain (bodies 4,5,6,7,8,10)
constexpr int NV = 12;           // dofs: 6 hinges + 6 free
constexpr int NQ = 13;
constexpr int NGEOM = 25;        // collidable geoms
constexpr int NPAIR_MAX = 192;
constexpr int NC = 24;           // contact capacity per env (== SO100_MAX_CONTACTS)
constexpr int JS = 13;           // row stride of the contact Jacobian (odd: conflict-free)
constexpr int STATE_WORDS = 64;  // HBM record per env (256 B, two 128 B lines)

// HBM state record layout (float words)
constexpr int S_QPOS = 0, S_QVEL = 13, S_CTRL = 25, S_WARM = 31, S_GOAL = 43, S_STEP = 46,
              S_TOTAL = 47, S_EPISODE = 48;

enum PairMode { MODE_BOX_MULTI = 0, MODE_BOX_SINGLE = 1, MODE_HULL = 2 };

struct DevGeom {
  int link;        // 0..5 arm link, 6 cube, -1 static
  int boxlike;     // 1: box geom or exact-cuboid mesh (SAT path), 0: general hull (GJK/EPA)
  int vadr, vnum;  // hull vertices (float4 pool)
  int mjid;
  float center[3]; // OBB centre in the body frame (static geoms: world)
  float half[3];   // OBB half sizes
  float rbound;
  float wmat[9];   // static geoms: world axes (row-major); dynamic: unused
  float org[3];    // static geoms: world origin of the body frame the hull vertices live in
};

struct DevPair {
  int g1, g2;      // collidable-geom indices, g1 has the lower (type, id)
  int mode, dim;
  float f0, f1;    // sliding / torsional friction
  float K, B;      // reference-acceleration stiffness / damping (solref, dmax folded in)
  float solimp[5]; // clamped
  float dtran, drot;  // diagApprox: body_invweight0 sums
  float omd0, dd;     // 1 - solimp[0] and solimp[1] - solimp[0], formed in fp64 on the host: with the cube's
                      // solimp clamped to 0.9999, (1 - imp) in float32 would lose 3 digits (R = (1-imp)/imp * diag)
  short l1, l2;       // link of g1 / g2 (0..5 arm link, 6 cube, -1 static): saves a dependent table load in the solver
};

struct DevModel {
  float timestep, gx, gy, gz, impratio, inv_scale;  // inv_scale = 1/(meaninertia*nv)
  int nsub, npair, ngeom, max_episode_steps, goal_max_steps, curriculum_steps;
  // kinematic chain
  float base_pos[3], base_quat[4];
  float link_pos[NL][3], link_quat[NL][4], link_axis[NL][3], link_ipos[NL][3];
  float link_mass[NL], link_Ib[NL][6];   // body-frame inertia tensor (xx,xy,xz,yy,yz,zz)
  float armature[NL];
  float cube_mass, cube_I[3];
  // constraint constants
  float fr_R[NV], fr_D[NV], fr_floss[NV], fr_B;
  float lim_lo[NL], lim_hi[NL], lim_invw[NL], lim_K, lim_B, lim_solimp[5];
  // actuators
  float kp[NL], kv[NL], ctrl_lo[NL], ctrl_hi[NL], frc_lo[NL], frc_hi[NL];
  // sites
  float ee_off[3];        // ee_site in the Fixed_Jaw (link 4) frame
  float cube_site_off[3]; // cube_site in the cube frame
  float bin_center[3];    // world (static body)
  double bin_min[3], bin_max[3];  // single_arm.py:64-75 in float64, compared against float32 cube_pos
  // task
  float start_pose[NL], act_lo[NL], act_hi[NL], act_range[NL];
  float box_lo[3], box_range[3];
  float cube_half, goal_threshold;
  float bin_goal_lo[3], bin_goal_hi[3], lift_xy, lift_zlo, lift_zhi;
  int cg_cube, cg_table;
  uint32_t pad_mask;
};

// ------------------------------------------------------------------ tiny vector algebra
struct V3 { float x, y, z; };
__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float comp(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ V3 normalized(V3 v) { float s = rsqrtf(fmaxf(dot(v, v), 1e-30f)); return v * s; }

struct Q4 { float w, x, y, z; };
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
  Q4 r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}
__device__ __forceinline__ Q4 qnormalize(Q4 q) {
  float s = rsqrtf(fmaxf(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, 1e-30f));
  q.w *= s; q.x *= s; q.y *= s; q.z *= s;
  return q;
}
// rotate v by unit quaternion q
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
  V3 u = mk(q.x, q.y, q.z);
  V3 t = cross(u, v) * 2.0f;
  return v + t * q.w + cross(u, t);
}
// row-major rotation matrix of unit quaternion
__device__ __forceinline__ void q2mat(Q4 q, float* m) {
  float w = q.w, x = q.x, y = q.y, z = q.z;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
__device__ __forceinline__ V3 mulmv(const float* m, V3 v) {   // M v
  return mk(fmaf(m[0], v.x, fmaf(m[1], v.y, m[2] * v.z)), fmaf(m[3], v.x, fmaf(m[4], v.y, m[5] * v.z)),
            fmaf(m[6], v.x, fmaf(m[7], v.y, m[8] * v.z)));
}
__device__ __forceinline__ V3 mulmtv(const float* m, V3 v) {  // M^T v
  return mk(fmaf(m[0], v.x, fmaf(m[3], v.y, m[6] * v.z)), fmaf(m[1], v.x, fmaf(m[4], v.y, m[7] * v.z)),
            fmaf(m[2], v.x, fmaf(m[5], v.y, m[8] * v.z)));
}
__device__ __forceinline__ V3 mcol(const float* m, int k) { return mk(m[k], m[3 + k], m[6 + k]); }
// symmetric 3x3 (xx,xy,xz,yy,yz,zz) times vector
__device__ __forceinline__ V3 symv(const float* s, V3 v) {
  return mk(s[0] * v.x + s[1] * v.y + s[2] * v.z, s[1] * v.x + s[3] * v.y + s[4] * v.z,
            s[2] * v.x + s[4] * v.y + s[5] * v.z);
}

// Philox4x32-10 (same spec as the oracle: counter = (env_lo, env_hi, episode, purpose), key = seed)
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                           uint32_t k1, uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float u01(uint32_t r) { return __uint2float_rn(r >> 8) * (1.0f / 16777216.0f); }

// lower-triangular index of (i >= j)
__device__ __forceinline__ int tri(int i, int j) { return (i * (i + 1)) / 2 + j; }

}  // namespace so100
