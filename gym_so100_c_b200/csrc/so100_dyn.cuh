// Position and velocity stages of one substep, one tile per env:
// forward kinematics as a prefix scan of rigid transforms over the 6-link chain, the arm mass matrix
// from pairwise lever-arm products, the RNE bias force via prefix-sum scans, position actuators and
// the unconstrained acceleration (6x6 Cholesky in registers).
#pragma once
#include <type_traits>
#include "so100_scratch.cuh"

namespace so100 {

// inclusive prefix sum over lanes 0..5 (other lanes carry garbage that never flows down)
template <unsigned LPE> __device__ __forceinline__ V3 scan6(const Tile<LPE>& t, V3 v, int lane) {
#pragma unroll
  for (int d = 1; d < 8; d <<= 1) {
    V3 o = shfl_up3(t, v, d);
    if (lane >= d) v = v + o;
  }
  return v;
}

// link frames from qpos (S->st[S_QPOS..]); DYN additionally fills the world CoM and inertia of every link
template <bool DYN, unsigned LPE, class ES> __device__ void kinematics(const Tile<LPE>& t, ES* S) {
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
    st3(S->f.lpos[lane], p);
#pragma unroll
    for (int k = 0; k < 9; k++) S->f.lmat[lane][k] = R[k];
    if (lane < NL) {
      st3(S->f.axis[lane], mulmv(R, ld3(c_m.link_axis[lane])));
      if constexpr (DYN) {
        st3(S->com[lane], p + mulmv(R, ld3(c_m.link_ipos[lane])));
        // Iw = R Ib R^T
        const float* I = c_m.link_Ib[lane];
        float T[9];
#pragma unroll
        for (int r = 0; r < 3; r++) {
          T[r * 3 + 0] = R[r * 3] * I[0] + R[r * 3 + 1] * I[1] + R[r * 3 + 2] * I[2];
          T[r * 3 + 1] = R[r * 3] * I[1] + R[r * 3 + 1] * I[3] + R[r * 3 + 2] * I[4];
          T[r * 3 + 2] = R[r * 3] * I[2] + R[r * 3 + 1] * I[4] + R[r * 3 + 2] * I[5];
        }
        float* Iw = S->Iw[lane];
        Iw[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2];
        Iw[1] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5];
        Iw[2] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8];
        Iw[3] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5];
        Iw[4] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8];
        Iw[5] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8];
      }
    }
  }
  t.sync();
}

// packed lower-triangle index e < 21 -> row
__device__ __forceinline__ int tri_row6(int e) { return (e >= 1) + (e >= 3) + (e >= 6) + (e >= 10) + (e >= 15); }

// u_il = a_i x (c_l - o_i), y_il = I_l a_i for l >= i; then M_ij = sum_{l>=i} m_l u_il.u_jl + a_j.y_il
template <unsigned LPE> __device__ void mass_matrix(const Tile<LPE>& t, KinS* S) {
  const int lane = t.thread_rank();
  for (int e = lane; e < 21; e += LPE) {
    const int l = tri_row6(e), i = e - tri(l, 0);
    V3 ai = ld3(S->f.axis[i]);
    st3(S->U[e], cross(ai, ld3(S->com[l]) - ld3(S->f.lpos[i])));
    st3(S->Y[e], symv(S->Iw[l], ai));
  }
  t.sync();
  for (int e = lane; e < 21; e += LPE) {
    const int i = tri_row6(e), j = e - tri(i, 0);
    V3 aj = ld3(S->f.axis[j]);
    float s = (i == j) ? c_m.armature[i] : 0.0f;
    for (int l = i; l < NL; l++)
      s += c_m.link_mass[l] * dot(ld3(S->U[tri(l, i)]), ld3(S->U[tri(l, j)])) + dot(aj, ld3(S->Y[tri(l, i)]));
    S->d.Mfull[i][j] = s;
    S->d.Mfull[j][i] = s;
  }
  t.sync();
}

// RNE bias via prefix scans, actuators, qfrc_smooth
template <unsigned LPE> __device__ void smooth_forces(const Tile<LPE>& t, KinS* S) {
  const int lane = t.thread_rank();
  V3 ax = mk(0, 0, 0), o = mk(0, 0, 0);
  float qd = 0;
  if (lane < NL) { ax = ld3(S->f.axis[lane]); o = ld3(S->f.lpos[lane]); qd = S->st[S_QVEL + lane]; }
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
    V3 c = ld3(S->com[lane]) - o;
    V3 ac = ao + cross(al, c) + cross(w, cross(w, c));
    V3 F = ac * c_m.link_mass[lane];
    V3 N = symv(S->Iw[lane], al) + cross(w, symv(S->Iw[lane], w));
    st3(&S->FN[lane][0], F);
    st3(&S->FN[lane][3], N);
  }
  t.sync();
  if (lane < NL) {
    float bias = 0;
    V3 ai = ld3(S->f.axis[lane]);
    for (int l = lane; l < NL; l++)
      bias += dot(ld3(S->U[tri(l, lane)]), ld3(&S->FN[l][0])) + dot(ai, ld3(&S->FN[l][3]));
    // position actuator: clip(kp (clip(ctrl) - q) - kv qd)
    float u = fminf(fmaxf(S->st[S_CTRL + lane], c_m.ctrl_lo[lane]), c_m.ctrl_hi[lane]);
    float f = c_m.kp[lane] * u - c_m.kp[lane] * S->st[S_QPOS + lane] - c_m.kv[lane] * qd;
    f = fminf(fmaxf(f, c_m.frc_lo[lane]), c_m.frc_hi[lane]);
    S->d.qfs[lane] = f - bias;
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
    S->d.qfs[lane] = -b;
  }
  t.sync();
}

// Compile-time loop: f(std::integral_constant<int, B>), ..., f(std::integral_constant<int, E - 1>).  The 6x6 factorisations below
// index their register arrays with these constants only: with ordinary `#pragma unroll` loops ptxas kept the triangular
// factor in LOCAL memory (51 LDL/STL in the light solve kernel, 146 in the dense ones, on the critical path of every Newton
// iteration).
template <int B, int E, class F> __device__ __forceinline__ void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}
__host__ __device__ constexpr int tri_c(int i, int j) { return (i * (i + 1)) / 2 + j; }

// In-register Cholesky of a 6x6 SPD block (FULL: row-major 6x6, else packed lower triangle): L[tri(i, j)] for i > j and
// 1 / L_jj on the diagonal
template <bool FULL> __device__ __forceinline__ void chol6_factor(const float* A, float* L) {
  static_for<0, NL>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    static_for<0, i + 1>([&](auto jc) {
      constexpr int j = decltype(jc)::value;
      L[tri_c(i, j)] = FULL ? A[i * NL + j] : A[tri_c(i, j)];
    });
  });
  static_for<0, NL>([&](auto jc) {
    constexpr int j = decltype(jc)::value;
    float d = L[tri_c(j, j)];
    static_for<0, j>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      d = fmaf(-L[tri_c(j, k)], L[tri_c(j, k)], d);
    });
    d = rsqrtf(fmaxf(d, 1e-20f));
    L[tri_c(j, j)] = d;   // 1 / L_jj
    static_for<j + 1, NL>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      float sum = L[tri_c(i, j)];
      static_for<0, j>([&](auto kc) {
        constexpr int k = decltype(kc)::value;
        sum = fmaf(-L[tri_c(i, k)], L[tri_c(j, k)], sum);
      });
      L[tri_c(i, j)] = sum * d;
    });
  });
}
// x <- L^-1 x
__device__ __forceinline__ void chol6_fwd(const float* L, float* x) {
  static_for<0, NL>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    float sum = x[i];
    static_for<0, i>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      sum = fmaf(-L[tri_c(i, k)], x[k], sum);
    });
    x[i] = sum * L[tri_c(i, i)];
  });
}
// x <- L^-T x
__device__ __forceinline__ void chol6_bwd(const float* L, float* x) {
  static_for<0, NL>([&](auto rc) {
    constexpr int i = NL - 1 - decltype(rc)::value;
    float sum = x[i];
    static_for<i + 1, NL>([&](auto kc) {
      constexpr int k = decltype(kc)::value;
      sum = fmaf(-L[tri_c(k, i)], x[k], sum);
    });
    x[i] = sum * L[tri_c(i, i)];
  });
}
// x = sign * A^-1 b
template <bool FULL> __device__ __forceinline__ void chol6_solve(const float* A, const float* b6, float sign, float* x) {
  float L[21];
  chol6_factor<FULL>(A, L);
  static_for<0, NL>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    x[i] = sign * b6[i];
  });
  chol6_fwd(L, x);
  chol6_bwd(L, x);
}

// qacc_smooth = M^-1 qfrc_smooth into S->d.qas: lane 0 factors the 6x6 arm block in registers, the cube block is diagonal
template <unsigned LPE> __device__ void smooth_acc(const Tile<LPE>& t, KinS* S) {
  const int lane = t.thread_rank();
  if (lane == 0) {
    float x[NL];
    chol6_solve<true>(&S->d.Mfull[0][0], S->d.qfs, 1.0f, x);
#pragma unroll
    for (int i = 0; i < NL; i++) S->d.qas[i] = x[i];
  } else if (lane >= NL && lane < NV) {
    S->d.qas[lane] = S->d.qfs[lane] / (lane < 9 ? c_m.cube_mass : c_m.cube_I[lane - 9]);
  }
  t.sync();
}

}  // namespace so100
