// lcr_kernels.cuh -- the env.step() hot path as warp-per-env CUDA (sm_100a).  Templated on the
// arithmetic type (float = product, double = verification build) and the number of cubes.
//
// Pipeline per substep (what mujoco.mj_step does for this model family; the reference calls it 20x
// per env.step at reach_cube_env.py:276-277):
//   kinematics -> inertia (CRB form, arm 6x6 + diagonal cubes) -> bias (RNE) -> collision ->
//   constraint rows (limits + elliptic contacts: Jacobian, impedance, regularisation, aref) ->
//   smooth forces (damping, position servos with +-frcrange) -> primal Newton solve with exact
//   line search -> implicitfast velocity update -> semi-implicit position update.
// Around it: apply_action incl. the damped-least-squares IK (reach_cube_env.py:148-279),
// get_observation (:281-295), reward / success (:313-348), reset (:297-311), TimeLimit(50).
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "lcr_device.cuh"

namespace lcr {

#define LCR_LS_MAXSMEM (227 * 1024 - 64)  // dynamic shared memory of one lockstep CTA (the static part is one int)
#define LANE (threadIdx.x & 31)
#define DI __device__ __forceinline__

template <typename T> DI T c_minval() { return (T)1e-15; }

// ---------------------------------------------------------------- warp helpers
template <typename T> DI T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}
// first-index argmax: larger value wins, ties go to the smaller index
template <typename T> DI void warp_argmax(T& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T ov = __shfl_xor_sync(FULLMASK, v, o);
    int oi = __shfl_xor_sync(FULLMASK, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
}

// float: two hardware warp reductions (REDUX) on an order-preserving integer key instead of five dependent shuffle
// rounds; same result (largest value, ties to the smaller index; -0 and +0 compare equal like the float compare)
template <> DI void warp_argmax<float>(float& v, int& idx) {
  const unsigned u = __float_as_uint(v + 0.0f);
  const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  const unsigned kmax = __reduce_max_sync(FULLMASK, key);
  idx = (int)__reduce_min_sync(FULLMASK, key == kmax ? (unsigned)idx : 0xffffffffu);
  v = __uint_as_float((kmax & 0x80000000u) ? (kmax & 0x7fffffffu) : ~kmax);
}

// row a and column b (a >= b) of entry e of a row-major packed lower triangle: e = a (a + 1) / 2 + b
DI void tri_index(int e, int& a, int& b) {
  a = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
  if (a * (a + 1) / 2 > e) a--;
  else if ((a + 1) * (a + 2) / 2 <= e) a++;
  b = e - a * (a + 1) / 2;
}

// ---------------------------------------------------------------- small math
template <typename T> DI T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> DI void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DI void quat_mul(T* r, const T* a, const T* b) {
  T w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  T x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  T y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  T z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
template <typename T> DI void quat_to_mat(T* m, const T* q) {
  T w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
template <typename T> DI void quat_normalize(T* q) {
  T n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < c_minval<T>()) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  T inv = 1 / n;
  q[0] *= inv; q[1] *= inv; q[2] *= inv; q[3] *= inv;
}
template <typename T> DI void mat_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
    z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DI void matT_vec(T* r, const T* m, const T* v) {
  T x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
    z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> DI T clampT(T x, T lo, T hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------- PCG64 (numpy) on the device
DI unsigned long long pcg64_next(unsigned long long* s) {
  const unsigned long long mh = 0x2360ED051FC65DA4ULL, ml = 0x4385DF649FCCF645ULL;
  unsigned long long sh = s[0], sl = s[1];
  unsigned long long lo = sl * ml;
  unsigned long long hi = __umul64hi(sl, ml) + sh * ml + sl * mh;
  unsigned long long nlo = lo + s[3];
  unsigned long long nhi = hi + s[2] + (nlo < lo ? 1ULL : 0ULL);
  s[0] = nhi; s[1] = nlo;
  unsigned long long x = nhi ^ nlo;
  unsigned rot = (unsigned)(nhi >> 58);
  return (x >> rot) | (x << ((64 - rot) & 63));
}
DI double pcg64_double(unsigned long long* s) { return (double)(pcg64_next(s) >> 11) * (1.0 / 9007199254740992.0); }

// ---------------------------------------------------------------- warp-cooperative dense SPD solves
// In-place Cholesky of the leading n x n block of A (row stride ld) in shared memory: lower
// triangle becomes L.  Lane i owns row i (n <= 32).
template <typename T> __device__ __noinline__ void warp_cholesky(T* A, int ld, int n) {
  const int i = LANE;
  for (int k = 0; k < n; k++) {
    T akk = A[k * ld + k];
    if (akk < c_minval<T>()) akk = c_minval<T>();
    T piv = sqrt(akk), lik = 0;
    if (i > k && i < n) { lik = A[i * ld + k] / piv; A[i * ld + k] = lik; }
    if (i == k) A[k * ld + k] = piv;
    __syncwarp();
    if (i > k && i < n)
      for (int j = k + 1; j <= i; j++) A[i * ld + j] -= lik * A[j * ld + k];
    __syncwarp();
  }
}
// Solve L L^T x = b; lane i passes b_i and receives x_i (lanes >= n pass anything, receive 0).
template <typename T> __device__ __noinline__ T warp_chol_solve(const T* L, int ld, int n, T b) {
  const int i = LANE;
  T y = (i < n) ? b : (T)0;
  for (int k = 0; k < n; k++) {
    T yk = __shfl_sync(FULLMASK, y, k) / L[k * ld + k];
    if (i == k) y = yk;
    else if (i > k && i < n) y -= L[i * ld + k] * yk;
  }
  for (int k = n - 1; k >= 0; k--) {
    T xk = __shfl_sync(FULLMASK, y, k) / L[k * ld + k];
    if (i == k) y = xk;
    else if (i < k) y -= L[k * ld + i] * xk;
  }
  return y;
}

// Register version for a compile-time size: lane i holds row i of the SPD matrix (a[0..i] = A[i][0..i], zeros in lanes >= N).
// On return a[k] = L[i][k] (k <= i) and c[j] = L[j][i] (j > i): lane i also holds column i of L, which is what the
// back substitution needs.  Same operations in the same order as warp_cholesky / warp_chol_solve, but the factor
// never touches shared memory and the loops are fully unrolled (a few shuffles and FMAs per column).
template <typename T, int N> DI void chol_reg(T (&a)[N], T (&c)[N]) {
  const int i = LANE;
#pragma unroll
  for (int k = 0; k < N; k++) {
    T akk = __shfl_sync(FULLMASK, a[k], k);
    if (akk < c_minval<T>()) akk = c_minval<T>();
    const T piv = sqrt(akk);
    const T lik = (i > k) ? a[k] / piv : (T)0;
    a[k] = (i == k) ? piv : lik;
#pragma unroll
    for (int j = k + 1; j < N; j++) {
      const T ljk = __shfl_sync(FULLMASK, lik, j);
      if (i >= j) a[j] -= lik * ljk;
      if (i == k) c[j] = ljk;
    }
  }
}
template <typename T, int N> DI T chol_reg_solve(const T (&a)[N], const T (&c)[N], T b) {
  const int i = LANE;
  T y = (i < N) ? b : (T)0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    const T yk = __shfl_sync(FULLMASK, y, k) / __shfl_sync(FULLMASK, a[k], k);
    if (i == k) y = yk;
    else if (i > k && i < N) y -= a[k] * yk;
  }
#pragma unroll
  for (int k = N - 1; k >= 0; k--) {
    const T xk = __shfl_sync(FULLMASK, y, k) / __shfl_sync(FULLMASK, a[k], k);
    if (i == k) y = xk;
    else if (i < k) y -= c[k] * xk;
  }
  return y;
}
// x = A^-1 b for the N x N SPD matrix stored row-major (stride ld) in shared memory; lane i passes b_i, gets x_i
template <typename T, int N> DI T chol_reg_factor_solve(const T* A, int ld, T b) {
  const int i = LANE;
  T a[N], c[N];
#pragma unroll
  for (int k = 0; k < N; k++) { a[k] = (i < N && k <= i) ? A[i * ld + k] : (T)0; c[k] = 0; }
  chol_reg<T, N>(a, c);
  return chol_reg_solve<T, N>(a, c, b);
}

// ---------------------------------------------------------------- kinematics
template <typename T, int NC>
__device__ __noinline__ void kinematics(Ws<T, NC>& w, const DevModel<T>& m) {
  const int lane = LANE;
  T p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
  const T* qpos = w.qpos();
#pragma unroll 1
  for (int b = 0; b < LCR_NABODY; b++) {  // serial chain, computed redundantly by every lane
    T R[9], t[3], bq[4] = {m.body_quat[b][0], m.body_quat[b][1], m.body_quat[b][2], m.body_quat[b][3]};
    T bp[3] = {m.body_pos[b][0], m.body_pos[b][1], m.body_pos[b][2]};
    quat_to_mat(R, q);
    mat_vec(t, R, bp);
    p[0] += t[0]; p[1] += t[1]; p[2] += t[2];
    quat_mul(q, q, bq);
    if (b >= 1) {
      const int j = b - 1;
      T ja[3] = {m.jnt_axis[j][0], m.jnt_axis[j][1], m.jnt_axis[j][2]}, ax[3], ql[4], sn, cs;
      quat_to_mat(R, q);
      mat_vec(ax, R, ja);
      sincos((T)0.5 * qpos[j], &sn, &cs);
      ql[0] = cs; ql[1] = sn * ja[0]; ql[2] = sn * ja[1]; ql[3] = sn * ja[2];
      quat_mul(q, q, ql);
      if (lane == 0) { w.axis[j][0] = ax[0]; w.axis[j][1] = ax[1]; w.axis[j][2] = ax[2]; }
    }
    if (lane == 0) {
      w.xpos[b][0] = p[0]; w.xpos[b][1] = p[1]; w.xpos[b][2] = p[2];
      w.xquat[b][0] = q[0]; w.xquat[b][1] = q[1]; w.xquat[b][2] = q[2]; w.xquat[b][3] = q[3];
    }
  }
  __syncwarp();
  if (lane < LCR_NABODY) {
    const int b = lane;
    T qq[4] = {w.xquat[b][0], w.xquat[b][1], w.xquat[b][2], w.xquat[b][3]}, R[9], t[3], qi[4];
    quat_to_mat(R, qq);
    T ip[3] = {m.body_ipos[b][0], m.body_ipos[b][1], m.body_ipos[b][2]};
    T iq[4] = {m.body_iquat[b][0], m.body_iquat[b][1], m.body_iquat[b][2], m.body_iquat[b][3]};
    mat_vec(t, R, ip);
#pragma unroll
    for (int k = 0; k < 9; k++) w.xmat[b][k] = R[k];
#pragma unroll
    for (int k = 0; k < 3; k++) w.xipos[b][k] = w.xpos[b][k] + t[k];
    quat_mul(qi, qq, iq);
    T Ri[9];
    quat_to_mat(Ri, qi);
    // world inertia Iw = Ri diag(I) Ri^T, symmetric storage xx xy xz yy yz zz
    T I0 = m.body_inertia[b][0], I1 = m.body_inertia[b][1], I2 = m.body_inertia[b][2];
    w.Iw[b][0] = Ri[0] * I0 * Ri[0] + Ri[1] * I1 * Ri[1] + Ri[2] * I2 * Ri[2];
    w.Iw[b][1] = Ri[0] * I0 * Ri[3] + Ri[1] * I1 * Ri[4] + Ri[2] * I2 * Ri[5];
    w.Iw[b][2] = Ri[0] * I0 * Ri[6] + Ri[1] * I1 * Ri[7] + Ri[2] * I2 * Ri[8];
    w.Iw[b][3] = Ri[3] * I0 * Ri[3] + Ri[4] * I1 * Ri[4] + Ri[5] * I2 * Ri[5];
    w.Iw[b][4] = Ri[3] * I0 * Ri[6] + Ri[4] * I1 * Ri[7] + Ri[5] * I2 * Ri[8];
    w.Iw[b][5] = Ri[6] * I0 * Ri[6] + Ri[7] * I1 * Ri[7] + Ri[8] * I2 * Ri[8];
    if (b == m.site_body) {
      T sp[3] = {m.site_pos[0], m.site_pos[1], m.site_pos[2]};
      mat_vec(t, R, sp);
      T* s = w.site_xpos();
      s[0] = w.xpos[b][0] + t[0]; s[1] = w.xpos[b][1] + t[1]; s[2] = w.xpos[b][2] + t[2];
    }
  } else if (lane < LCR_NABODY + Scene<NC>::NBOX && lane >= LCR_NABODY + Scene<NC>::NCUBE) {
    // static wall boxes: constant pose in the slots after the cubes
    const int b = lane, wi = lane - LCR_NABODY - Scene<NC>::NCUBE;
#pragma unroll
    for (int k = 0; k < 3; k++) w.xpos[b][k] = m.wall_pos[wi][k];
#pragma unroll
    for (int k = 0; k < 4; k++) w.xquat[b][k] = k == 0 ? (T)1 : (T)0;
#pragma unroll
    for (int k = 0; k < 9; k++) w.xmat[b][k] = (k & 3) == 0 ? (T)1 : (T)0;
  } else if (lane < LCR_NABODY + Scene<NC>::NCUBE) {
    const int c = lane - LCR_NABODY, b = lane;
    const T* qp = qpos + LCR_NARM + 7 * c;
    T qq[4] = {qp[3], qp[4], qp[5], qp[6]}, R[9];
    quat_normalize(qq);
    quat_to_mat(R, qq);
    T* cx = w.cube_xpos(c);
#pragma unroll
    for (int k = 0; k < 3; k++) { w.xpos[b][k] = qp[k]; cx[k] = qp[k]; }
#pragma unroll
    for (int k = 0; k < 4; k++) w.xquat[b][k] = qq[k];
#pragma unroll
    for (int k = 0; k < 9; k++) w.xmat[b][k] = R[k];
  }
  __syncwarp();
}

// ---------------------------------------------------------------- arm inertia matrix + bias forces
template <typename T, int NC>
__device__ __noinline__ void inertia_and_bias(Ws<T, NC>& w, const DevModel<T>& m) {
  const int lane = LANE;
  // M: lane e < 21 -> lower-triangle entry (i, j), i >= j
  if (lane < 21) {
    int i, j;
    tri_index(lane, i, j);
    T zi[3] = {w.axis[i][0], w.axis[i][1], w.axis[i][2]}, zj[3] = {w.axis[j][0], w.axis[j][1], w.axis[j][2]};
    T acc = 0;
    for (int b = i + 1; b < LCR_NABODY; b++) {
      T ri[3], rj[3], ci[3], cj[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { ri[k] = w.xipos[b][k] - w.xpos[i + 1][k]; rj[k] = w.xipos[b][k] - w.xpos[j + 1][k]; }
      cross3(ci, zi, ri);
      cross3(cj, zj, rj);
      const T* I = w.Iw[b];
      T Iz[3] = {I[0] * zj[0] + I[1] * zj[1] + I[2] * zj[2], I[1] * zj[0] + I[3] * zj[1] + I[4] * zj[2],
                 I[2] * zj[0] + I[4] * zj[1] + I[5] * zj[2]};
      acc += m.body_mass[b] * dot3(ci, cj) + dot3(zi, Iz);
    }
    if (i == j) acc += m.jnt_armature[i];
    w.M[i][j] = acc;
    w.M[j][i] = acc;
  }
  // RNE forward recursion (serial; every lane computes it, lane 0 stores)
  {
    T wv[3] = {0, 0, 0}, al[3] = {0, 0, 0}, a[3] = {-m.gravity[0], -m.gravity[1], -m.gravity[2]};
    const T* qvel = w.qvel();
#pragma unroll 1
    for (int b = 1; b < LCR_NABODY; b++) {
      const int j = b - 1;
      T zq[3] = {w.axis[j][0] * qvel[j], w.axis[j][1] * qvel[j], w.axis[j][2] * qvel[j]}, r[3], t1[3], t2[3];
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = w.xpos[b][k] - w.xpos[b - 1][k];
      cross3(t1, al, r);
      cross3(t2, wv, r);
      cross3(t2, wv, t2);
#pragma unroll
      for (int k = 0; k < 3; k++) a[k] += t1[k] + t2[k];  // uses w, al of the parent
      cross3(t1, wv, zq);
#pragma unroll
      for (int k = 0; k < 3; k++) { al[k] += t1[k]; wv[k] += zq[k]; }
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) { w.rw[b][k] = wv[k]; w.ral[b][k] = al[k]; w.ra[b][k] = a[k]; }
      }
    }
  }
  __syncwarp();
  if (lane >= 1 && lane < LCR_NABODY) {
    const int b = lane;
    T wv[3] = {w.rw[b][0], w.rw[b][1], w.rw[b][2]}, al[3] = {w.ral[b][0], w.ral[b][1], w.ral[b][2]};
    T rc[3], t1[3], t2[3];
#pragma unroll
    for (int k = 0; k < 3; k++) rc[k] = w.xipos[b][k] - w.xpos[b][k];
    cross3(t1, al, rc);
    cross3(t2, wv, rc);
    cross3(t2, wv, t2);
    T mass = m.body_mass[b];
#pragma unroll
    for (int k = 0; k < 3; k++) w.F[b][k] = mass * (w.ra[b][k] + t1[k] + t2[k]);
    const T* I = w.Iw[b];
    T Iww[3] = {I[0] * wv[0] + I[1] * wv[1] + I[2] * wv[2], I[1] * wv[0] + I[3] * wv[1] + I[4] * wv[2],
                I[2] * wv[0] + I[4] * wv[1] + I[5] * wv[2]};
    T Iwa[3] = {I[0] * al[0] + I[1] * al[1] + I[2] * al[2], I[1] * al[0] + I[3] * al[1] + I[4] * al[2],
                I[2] * al[0] + I[4] * al[1] + I[5] * al[2]};
    cross3(t1, wv, Iww);
#pragma unroll
    for (int k = 0; k < 3; k++) w.Nn[b][k] = Iwa[k] + t1[k];
  }
  __syncwarp();
  if (lane < LCR_NARM) {
    const int j = lane;
    T tq[3] = {0, 0, 0};
    for (int b = j + 1; b < LCR_NABODY; b++) {
      T r[3], t[3], F[3] = {w.F[b][0], w.F[b][1], w.F[b][2]};
#pragma unroll
      for (int k = 0; k < 3; k++) r[k] = w.xipos[b][k] - w.xpos[j + 1][k];
      cross3(t, r, F);
#pragma unroll
      for (int k = 0; k < 3; k++) tq[k] += w.Nn[b][k] + t[k];
    }
    w.bias[j] = w.axis[j][0] * tq[0] + w.axis[j][1] * tq[1] + w.axis[j][2] * tq[2];
  } else if (lane < Ws<T, NC>::NVV) {
    const int d = lane - LCR_NARM, c = d / 6, k = d % 6;
    w.bias[lane] = (k < 3) ? -m.cube_mass[c] * m.gravity[k] : (T)0;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- collision
template <typename T> DI void make_frame(T* f) {
  T y[3] = {0, 0, 0};
  if (f[1] < (T)0.5 && f[1] > (T)-0.5) y[1] = 1; else y[2] = 1;
  T d = dot3(f, y);
#pragma unroll
  for (int k = 0; k < 3; k++) y[k] -= d * f[k];
  T inv = 1 / sqrt(dot3(y, y));
#pragma unroll
  for (int k = 0; k < 3; k++) f[3 + k] = y[k] * inv;
  cross3(f + 6, f, f + 3);
}

// Append one contact (warp-uniform arguments; lane 0 writes).  ncon / nefc are warp-uniform
// registers owned by make_constraints.  Returns true if stored.
template <typename T, int NC>
DI bool add_contact(Ws<T, NC>& w, const DevModel<T>& m, int& ncon, int& nefc, const CPar<T>* par, int b1, int b2, const T* pos, const T* normal, T dist) {
  const int dim = par->dim;
  if (ncon >= Ws<T, NC>::MAXCON || nefc + dim > Ws<T, NC>::MAXEFC) {
    if (LANE == 0) { w.diag[4]++; w.ovf = 1; }
    return false;
  }
  const int ci = ncon;
  if (LANE == 0) {
    T f[9];
    f[0] = normal[0]; f[1] = normal[1]; f[2] = normal[2];
    make_frame(f);
#pragma unroll
    for (int k = 0; k < 3; k++) w.c_pos[ci][k] = pos[k];
#pragma unroll
    for (int k = 0; k < 9; k++) w.c_frame[ci][k] = f[k];
    w.c_dist[ci] = dist;
    w.c_par[ci] = (short)m.par_index(par);
    w.c_b1[ci] = (signed char)b1;
    w.c_b2[ci] = (signed char)b2;
    w.c_efc[ci] = (short)nefc;
  }
  ncon = ci + 1;
  nefc += dim;
  return true;
}

template <typename T, int NC>
DI void collide_floor_cube(Ws<T, NC>& w, const DevModel<T>& m, int& ncon, int& nefc, int c) {
  const int lane = LANE, b = LCR_NABODY + c, i = lane & 7;
  T v[3] = {m.cube_size[c][0] * ((i & 1) ? 1 : -1), m.cube_size[c][1] * ((i & 2) ? 1 : -1), m.cube_size[c][2] * ((i & 4) ? 1 : -1)};
  T wv[3];
  mat_vec(wv, w.xmat[b], v);
  const T dist = w.xpos[b][2], ld = wv[2], d = dist + ld;
  const bool hit = lane < 8 && !(d > 0 || ld > 0) && d < 0;
  unsigned mask = __ballot_sync(FULLMASK, hit);
  int cnt = 0;
  const T n[3] = {0, 0, 1};
  while (mask && cnt < 4) {
    const int src = __ffs(mask) - 1;
    mask &= mask - 1;
    T pos[3], dd = __shfl_sync(FULLMASK, d, src);
    pos[0] = w.xpos[b][0] + __shfl_sync(FULLMASK, wv[0], src);
    pos[1] = w.xpos[b][1] + __shfl_sync(FULLMASK, wv[1], src);
    pos[2] = w.xpos[b][2] + __shfl_sync(FULLMASK, wv[2], src) - (T)0.5 * dd;
    if (add_contact(w, m, ncon, nefc, &m.par_floor_cube[c], -1, b, pos, n, dd)) cnt++;
  }
}

}  // namespace lcr
#include "lcr_convex.cuh"
namespace lcr {

// support vertices of mesh g along 4 world directions at once (warp-cooperative, lane-strided)
template <typename T, int NC>
DI void mesh_support4(const Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, int g, const T (*dirs)[3],
                      int* idx, T (*pts)[3]) {
  const int lane = LANE, b = m.mesh_body[g], adr = m.mesh_vertadr[g], num = m.mesh_vertnum[g];
  T dl[4][3], bv[4];
  int bi[4];
#pragma unroll
  for (int t = 0; t < 4; t++) { matT_vec(dl[t], w.xmat[b], dirs[t]); bv[t] = (T)-1e30; bi[t] = 0x7fffffff; }
  for (int base = lane; base < num; base += 32 * kScanUnroll) {  // batched loads, see shape_support
    T vx[kScanUnroll], vy[kScanUnroll], vz[kScanUnroll];
#pragma unroll
    for (int u = 0; u < kScanUnroll; u++) {
      const int i = base + 32 * u;
      load_vert(verts, adr + (i < num ? i : num - 1), vx[u], vy[u], vz[u]);
    }
#pragma unroll
    for (int u = 0; u < kScanUnroll; u++) {
      const int i = base + 32 * u;
#pragma unroll
      for (int t = 0; t < 4; t++) {
        T s = vx[u] * dl[t][0] + vy[u] * dl[t][1] + vz[u] * dl[t][2];
        if (i < num && s > bv[t]) { bv[t] = s; bi[t] = i; }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 4; t++) {
    warp_argmax(bv[t], bi[t]);
    idx[t] = bi[t];
    T vl[3];
    load_vert(verts, adr + bi[t], vl[0], vl[1], vl[2]);  // (cold path: four winners, one reload each)
    mat_vec(pts[t], w.xmat[b], vl);
#pragma unroll
    for (int k = 0; k < 3; k++) pts[t][k] += w.xpos[b][k];
  }
}

template <typename T, int NC>
__device__ __noinline__ void collide_floor_meshes(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, int& ncon, int& nefc) {
  const int lane = LANE;
  bool cand = false;
  if (lane < m.nmesh && m.mesh_body[lane] != 0) cand = !(w.gc[lane][2] - m.mesh_rbound[lane] > 0);
  unsigned mask = __ballot_sync(FULLMASK, cand);
  const T n[3] = {0, 0, 1};
  const T dirs[4][3] = {{0, 0, -1}, {(T)1e-3, 0, -1}, {(T)-0.5e-3, (T)0.8660254037844386e-3, -1}, {(T)-0.5e-3, (T)-0.8660254037844386e-3, -1}};
  while (mask) {
    const int g = __ffs(mask) - 1;
    mask &= mask - 1;
    int idx[4];
    T pts[4][3];
    mesh_support4(w, m, verts, g, dirs, idx, pts);
    if (pts[0][2] >= 0) continue;
    int used[4], cnt = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      T d = pts[t][2];
      if (d >= 0) continue;
      bool dup = false;
      for (int k = 0; k < cnt; k++) dup |= (used[k] == idx[t]);
      if (dup) continue;
      T pos[3] = {pts[t][0], pts[t][1], pts[t][2] - (T)0.5 * d};
      if (add_contact(w, m, ncon, nefc, &m.par_floor_mesh[g], -1, m.mesh_body[g], pos, n, d)) used[cnt++] = idx[t];
    }
  }
}

// ---------------------------------------------------------------- constraint rows
template <typename T> DI T impedance(const T* si, T pos) {
  if (si[0] == si[1] || si[2] <= c_minval<T>()) return (T)0.5 * (si[0] + si[1]);
  T x = fabs(pos) / si[2];
  if (x >= 1) return si[1];
  if (x <= 0) return si[0];
  T y;
  if (si[4] == 1) y = x;
  else if (si[4] == 2) y = (x <= si[3]) ? x * x / si[3] : 1 - (1 - x) * (1 - x) / (1 - si[3]);
  else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
  else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
  return si[0] + y * (si[1] - si[0]);
}

template <typename T, int NC> DI T body_invweight(const DevModel<T>& m, int b, int rot) { return b < 0 ? (T)0 : m.body_invweight0[b][rot]; }

// Fill row i of J for contact row (axis ax, point pos, bodies b1 -> b2): ax . (Jac_b2 - Jac_b1).
template <typename T, int NC>
DI void fill_jac_row(Ws<T, NC>& w, int i, int b1, int b2, const T* pos, const T* ax, bool rot) {
  constexpr int NVV = Ws<T, NC>::NVV;
  T* row = w.J[i];
#pragma unroll
  for (int d = 0; d < NVV; d++) row[d] = 0;
  const int n1 = (b1 > 0 && b1 < LCR_NABODY) ? b1 : 0, n2 = (b2 > 0 && b2 < LCR_NABODY) ? b2 : 0;
  const int nmax = n1 > n2 ? n1 : n2;
#pragma unroll 1
  for (int d = 0; d < nmax; d++) {  // hinge d moves arm body b iff d < b
    const int sgn = (d < n2 ? 1 : 0) - (d < n1 ? 1 : 0);
    if (sgn == 0) continue;
    const T* z = w.axis[d];
    T val;
    if (rot) val = dot3(ax, z);
    else {
      T r[3] = {pos[0] - w.xpos[d + 1][0], pos[1] - w.xpos[d + 1][1], pos[2] - w.xpos[d + 1][2]}, c[3];
      cross3(c, z, r);
      val = dot3(ax, c);
    }
    row[d] = sgn > 0 ? val : -val;
  }
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    const int b = side ? b1 : b2;
    if (b < LCR_NABODY) continue;
    const T sg = side ? (T)-1 : (T)1;
    T* rc = row + LCR_NARM + 6 * (b - LCR_NABODY);
    T r[3] = {pos[0] - w.xpos[b][0], pos[1] - w.xpos[b][1], pos[2] - w.xpos[b][2]};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      T bx[3] = {w.xmat[b][k], w.xmat[b][3 + k], w.xmat[b][6 + k]};
      if (rot) rc[3 + k] += sg * dot3(ax, bx);
      else {
        T c[3];
        cross3(c, bx, r);
        rc[k] += sg * ax[k];
        rc[3 + k] += sg * dot3(ax, c);
      }
    }
  }
}

template <typename T, int NC>
__device__ __noinline__ void make_constraints(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, bool precomputed) {
  constexpr int NVV = Ws<T, NC>::NVV;
  const int lane = LANE;
  const T* qpos = w.qpos();
  const T* qvel = w.qvel();
  int ncon = 0, nefc = 0, nlim = 0;
  if (lane == 0) w.diag[4] = 0;
  // joint limits: lane l < 12 -> (joint l/2, side l&1)
  {
    bool viol = false;
    T dist = 0;
    const int j = lane >> 1, side = lane & 1;
    if (lane < 2 * LCR_NARM) {
      dist = side == 0 ? qpos[j] - m.jnt_range[j][0] : m.jnt_range[j][1] - qpos[j];
      viol = dist < 0;
    }
    const unsigned vmask = __ballot_sync(FULLMASK, viol);
    if (viol) {
      const int i = __popc(vmask & ((1u << lane) - 1));
      for (int d = 0; d < NVV; d++) w.J[i][d] = 0;
      w.J[i][j] = side == 0 ? (T)1 : (T)-1;
      w.e_unit[i] = (short)(-1 - j);
      w.e_r[i] = 0;
      w.e_pos[i] = dist;
    }
    nefc = nlim = __popc(vmask);
  }
  __syncwarp();
  const int cmask = m.collision_mask;
  if (!precomputed) {  // fused path: broadphase and narrowphase jobs inline
    collect_candidates(w, m);
    run_jobs_inline(w, m, verts);
  }
  // generation order (= drop order at the caps): floor-cube, cube-cube, wall-cube, cube-mesh, wall-mesh, floor-mesh, mesh-mesh
  if (cmask & LCR_COLLIDE_FLOOR_CUBE)
    for (int c = 0; c < Scene<NC>::NCUBE; c++) collide_floor_cube(w, m, ncon, nefc, c);
  if (Scene<NC>::NCUBE == 2 && (cmask & LCR_COLLIDE_CUBE_CUBE)) collide_box_box(w, m, ncon, nefc, 0, 1, &m.par_cube_cube);
  if (Scene<NC>::NWALL > 0 && (cmask & LCR_COLLIDE_WALL_CUBE))
    for (int wi = 0; wi < Scene<NC>::NWALL; wi++)
      for (int c = 0; c < Scene<NC>::NCUBE; c++) collide_box_box(w, m, ncon, nefc, Scene<NC>::NCUBE + wi, c, &m.par_wall_cube[wi][c]);
  if (cmask & (LCR_COLLIDE_CUBE_MESH | LCR_COLLIDE_WALL_MESH)) consume_candidates(w, m, ncon, nefc, true);
  if (cmask & LCR_COLLIDE_FLOOR_MESH) collide_floor_meshes(w, m, verts, ncon, nefc);
  if (cmask & LCR_COLLIDE_MESH_MESH) consume_candidates(w, m, ncon, nefc, false);
#ifdef LCR_FLOW_DEBUG
  {  // debug: the counts are warp-uniform registers -- every lane must hold the same values
    const int n0 = __shfl_sync(FULLMASK, ncon, 0), e0 = __shfl_sync(FULLMASK, nefc, 0);
    if (__any_sync(FULLMASK, ncon != n0 || nefc != e0 || ncon > Ws<T, NC>::MAXCON || nefc > Ws<T, NC>::MAXEFC)) {
      if (lane == 24) printf("make_constraints: lane counts differ blk %d warp %d lane0 ncon %d nefc %d lane24 ncon %d nefc %d ncand %d nlim %d pre %d\n", blockIdx.x, threadIdx.x >> 5, n0, e0, ncon, nefc, w.ncand, nlim, (int)precomputed);
      __trap();
    }
  }
#endif
  if (lane == 0) { w.ncon = ncon; w.nefc = nefc; w.nlim = nlim; }
  __syncwarp();
  // contact rows: lane <-> row
  for (int ci = lane; ci < ncon; ci += 32) {
    const int dim = m.par(w.c_par[ci])->dim, e0 = w.c_efc[ci];
    for (int r = 0; r < dim; r++) { w.e_unit[e0 + r] = (short)ci; w.e_r[e0 + r] = (signed char)r; }
  }
  __syncwarp();
  for (int i = nlim + lane; i < nefc; i += 32) {
    const int ci = w.e_unit[i], r = w.e_r[i], b1 = w.c_b1[ci], b2 = w.c_b2[ci];
    const T* ax = w.c_frame[ci] + 3 * (r % 3);
    const T* pos = w.c_pos[ci];
    const bool rot = r >= 3;
    fill_jac_row(w, i, b1, b2, pos, ax, rot);
    w.e_pos[i] = r == 0 ? w.c_dist[ci] : (T)0;
  }
  __syncwarp();
  // impedance, regularisation R (stored in e_D for now), reference acceleration
  for (int i = lane; i < nefc; i += 32) {
    T v = 0;
    for (int d = 0; d < NVV; d++) v += w.J[i][d] * qvel[d];
    const int u = w.e_unit[i];
    const CPar<T>* par;
    T diagA;
    if (u < 0) { par = &m.par_limit[-1 - u]; diagA = m.dof_invweight0[-1 - u]; }
    else {
      par = m.par(w.c_par[u]);
      const int rot = w.e_r[i] >= 3;
      diagA = body_invweight<T, NC>(m, w.c_b1[u], rot) + body_invweight<T, NC>(m, w.c_b2[u], rot);
    }
    const T pos = w.e_pos[i], imp = impedance(par->si, pos);
    T R = (1 - imp) * diagA / imp;
    if (R < c_minval<T>()) R = c_minval<T>();
    w.e_D[i] = R;
    w.e_aref[i] = -par->B * v - par->K * imp * pos;
  }
  __syncwarp();
  // elliptic friction rows: R from the normal row and impratio; mu
  for (int base = nlim; base < nefc; base += 32) {  // warp-uniform trip count; per chunk: read, sync, write
    const int i = base + lane;
    const bool act = i < nefc;
    T Rnew = 0;
    if (act) {
      const int u = w.e_unit[i], r = w.e_r[i];
      const CPar<T>* par = m.par(w.c_par[u]);
      const T R0 = w.e_D[w.c_efc[u]];
      T ir = m.impratio < c_minval<T>() ? c_minval<T>() : m.impratio;
      const T R1 = R0 / ir;
      if (r == 0) { Rnew = R0; w.c_mu[u] = par->dim > 1 ? par->fr[0] * sqrt(R1 / R0) : (T)0; }
      else Rnew = R1 * par->fr[0] * par->fr[0] / (par->fr[r - 1] * par->fr[r - 1]);
    }
    __syncwarp();
    if (act) w.e_D[i] = Rnew;
    __syncwarp();
  }
  for (int i = lane; i < nefc; i += 32) w.e_D[i] = 1 / w.e_D[i];
  if (lane == 0) {
    w.diag[0] = ncon; w.diag[1] = nefc;
    if (nefc > w.diag[3]) w.diag[3] = nefc;
  }
  __syncwarp();
}

// ---------------------------------------------------------------- smooth forces
template <typename T, int NC>
__device__ __noinline__ void smooth_forces(Ws<T, NC>& w, const DevModel<T>& m) {
  constexpr int NVV = Ws<T, NC>::NVV;
  const int lane = LANE;
  const T* qpos = w.qpos();
  const T* qvel = w.qvel();
  T sm = 0;
  if (lane < LCR_NARM) {
    const int j = lane;
    T u = clampT(w.ctrl()[j], m.act_ctrlrange[j][0], m.act_ctrlrange[j][1]);
    T f = m.act_kp[j] * (u - qpos[j]) - m.act_kv[j] * qvel[j];
    f = clampT(f, m.jnt_frcrange[j][0], m.jnt_frcrange[j][1]);
    sm = -m.jnt_damping[j] * qvel[j] - w.bias[j] + f;
  } else if (lane < NVV) sm = -w.bias[lane];
  if (lane < NVV) w.smooth[lane] = sm;
  // qacc_smooth of the arm block: M_arm x = smooth (Cholesky in registers)
  __syncwarp();
  T x = chol_reg_factor_solve<T, LCR_NARM>(&w.M[0][0], LCR_NARM, sm);
  if (lane >= LCR_NARM && lane < NVV) {
    const int d = lane - LCR_NARM, c = d / 6;
    x = sm / ((d % 6) < 3 ? m.cube_mass[c] : m.cube_inertia[c]);
  }
  if (lane < NVV) w.qacc_smooth[lane] = x;
  __syncwarp();
}

// y_lane = (M x)_lane with M = blockdiag(M_arm, cube diagonals); x read from shared memory
template <typename T, int NC> DI T mul_M(const Ws<T, NC>& w, const DevModel<T>& m, const T* x) {
  const int lane = LANE;
  T a = 0;
  if (lane < LCR_NARM) {
#pragma unroll
    for (int d = 0; d < LCR_NARM; d++) a += w.M[lane][d] * x[d];
  } else if (lane < Ws<T, NC>::NVV) {
    const int d = lane - LCR_NARM, c = d / 6;
    a = ((d % 6) < 3 ? m.cube_mass[c] : m.cube_inertia[c]) * x[lane];
  }
  return a;
}

// ---------------------------------------------------------------- primal Newton solver
// Evaluate one elliptic contact at x = jar + alpha*jv.  zone: 0 top (satisfied), 1 bottom (quadratic
// in all rows), 2 middle (cone).  Mirrors the cost used by MuJoCo's primal solvers.
// With FULL the lane also publishes, per row k of the contact, the force and the pieces of the
// Hessian in the factored form
//     J_c^T Hc J_c = sum_k wrow_k J_k J_k^T + c1 (J_c^T g)(J_c^T g)^T - c2 (J_c^T p)(J_c^T p)^T
// (bottom zone: wrow = D, c1 = c2 = 0;  middle zone: wrow_0 = 0, wrow_k = c2 f_k^2,
//  g = dNT/djar, p_k = f_k u_k / T, c1 = Dm, c2 = -mu NT Dm / T >= 0).
template <typename T> struct Eval3 { T cost, d1, d2; };
template <typename T, int NC, bool with_jv, bool FULL>
__device__ __noinline__ Eval3<T> contact_eval(Ws<T, NC>& w, const DevModel<T>& m, int ci, T alpha) {
  T cost, d1, d2;
  const CPar<T>* par = m.par(w.c_par[ci]);
  const int dim = par->dim, i0 = w.c_efc[ci];
  const T mu = w.c_mu[ci];
  T x[6], u[6], fri[6], jv[6];
  cost = d1 = d2 = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j < dim) {
      jv[j] = with_jv ? w.e_jv[i0 + j] : (T)0;
      x[j] = w.e_jar[i0 + j] + alpha * jv[j];
    } else { jv[j] = 0; x[j] = 0; }
  }
  fri[0] = mu;
  T T2 = 0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    if (j > 0) fri[j] = j < dim ? par->fr[j - 1] : (T)0;
    u[j] = x[j] * fri[j];
    if (j > 0) T2 += u[j] * u[j];
  }
  const T N = u[0], Tn = sqrt(T2);
  int zone;
  if (dim == 1) zone = x[0] < 0 ? 1 : 0;
  else if (N >= mu * Tn || (Tn <= 0 && N >= 0)) zone = 0;
  else if (mu * N + Tn <= 0 || (Tn <= 0 && N < 0)) zone = 1;
  else zone = 2;
  if (zone == 0) {
    if (FULL) {
#pragma unroll
      for (int j = 0; j < 6; j++) if (j < dim) { w.e_force[i0 + j] = 0; w.e_w[i0 + j] = 0; }
      w.c_c1[ci] = 0; w.c_c2[ci] = 0;
    }
    return Eval3<T>{cost, d1, d2};
  }
  if (zone == 1) {
#pragma unroll
    for (int j = 0; j < 6; j++)
      if (j < dim) {
        const T D = w.e_D[i0 + j];
        cost += (T)0.5 * D * x[j] * x[j]; d1 += D * x[j] * jv[j]; d2 += D * jv[j] * jv[j];
        if (FULL) { w.e_force[i0 + j] = -D * x[j]; w.e_w[i0 + j] = D; }
      }
    if (FULL) { w.c_c1[ci] = 0; w.c_c2[ci] = 0; }
    return Eval3<T>{cost, d1, d2};
  }
  const T Dm = w.e_D[i0] / (mu * mu * (1 + mu * mu)), NT = N - mu * Tn, iT = 1 / Tn;
  cost = (T)0.5 * Dm * NT * NT;
  if (with_jv) {
    T N1 = mu * jv[0], T1 = 0, up2 = 0;
#pragma unroll
    for (int j = 1; j < 6; j++) { T up = fri[j] * jv[j]; T1 += u[j] * up; up2 += up * up; }
    T1 *= iT;
    const T T2d = up2 * iT - T1 * T1 * iT, NT1 = N1 - mu * T1, NT2 = -mu * T2d;
    d1 = Dm * NT * NT1;
    d2 = Dm * (NT1 * NT1 + NT * NT2);
  }
  if (FULL) {
    const T f0 = -Dm * NT * mu, c2 = -mu * NT * iT * Dm;
    w.e_force[i0] = f0; w.e_w[i0] = 0; w.e_g[i0] = mu; w.e_p[i0] = 0;
#pragma unroll
    for (int j = 1; j < 6; j++)
      if (j < dim) {
        const T pj = fri[j] * u[j] * iT;
        w.e_force[i0 + j] = -f0 * pj;
        w.e_w[i0 + j] = c2 * fri[j] * fri[j];
        w.e_g[i0 + j] = -mu * pj;
        w.e_p[i0 + j] = pj;
      }
    w.c_c1[ci] = Dm; w.c_c2[ci] = c2;
  }
  return Eval3<T>{cost, d1, d2};
}

// (M x)_dof with M = blockdiag(M_arm, cube diagonals); x read from shared memory
template <typename T, int NC> DI T mul_M_dof(const Ws<T, NC>& w, const DevModel<T>& m, const T* x, int dof) {
  T a = 0;
  if (dof < LCR_NARM) {
#pragma unroll
    for (int d = 0; d < LCR_NARM; d++) a += w.M[dof][d] * x[d];
  } else if (dof < Ws<T, NC>::NVV) {
    const int d = dof - LCR_NARM, c = d / 6;
    a = ((d % 6) < 3 ? m.cube_mass[c] : m.cube_inertia[c]) * x[dof];
  }
  return a;
}

// cost at qacc over the dof island [d0, NVV): fills e_jar, Ma; if FULL also e_force, grad and the Hessian pieces
// (e_w, e_g, e_p, c_c1, c_c2).  Lane l owns dof d0 + l.
template <typename T, int NC, bool FULL>
__device__ __noinline__ T total_cost(Ws<T, NC>& w, const DevModel<T>& m, const T* qacc, int d0) {
  constexpr int NVV = Ws<T, NC>::NVV;
  const int lane = LANE, nefc = w.nefc, ncon = w.ncon, nlim = w.nlim, dof = d0 + lane;
  for (int i = lane; i < nefc; i += 32) {
    T a = -w.e_aref[i];
    for (int d = d0; d < NVV; d++) a += w.J[i][d] * qacc[d];
    w.e_jar[i] = a;
  }
  T cost = 0;
  if (dof < NVV) {
    const T ma = mul_M_dof(w, m, qacc, dof);
    w.Ma[dof] = ma;
    cost = (T)0.5 * (ma - w.smooth[dof]) * (qacc[dof] - w.qacc_smooth[dof]);
  }
  __syncwarp();
  if (lane < nlim) {
    const T x = w.e_jar[lane];
    T f = 0, ww = 0;
    if (x < 0) { cost += (T)0.5 * w.e_D[lane] * x * x; f = -w.e_D[lane] * x; ww = w.e_D[lane]; }
    if (FULL) { w.e_force[lane] = f; w.e_w[lane] = ww; }
  }
  for (int ci = lane; ci < ncon; ci += 32) cost += contact_eval<T, NC, false, FULL>(w, m, ci, (T)0).cost;
  cost = warp_sum(cost);
  if (FULL) {
    __syncwarp();
    if (dof < NVV) {
      T g = w.Ma[dof] - w.smooth[dof];
#pragma unroll 4
      for (int i = 0; i < nefc; i++) g -= w.J[i][dof] * w.e_force[i];
      w.grad[dof] = g;
    }
  }
  __syncwarp();
  return cost;
}

// Primal Newton solve.  If no constraint row touches an arm dof (no limit rows, no contact on links 1..6) the arm
// block of the problem is decoupled: qacc_arm = qacc_smooth_arm exactly and the solve runs on the cube dofs only
// (island [6, NVV): 6x6 or 12x12 Hessian instead of 12x12 / 18x18).
// CTA_SYNC (lockstep kernel): every warp of the CTA calls this together (`active` false for warps without work)
// and the Newton iterations are separated by CTA barriers, so that the warps of a CTA walk the same code at the
// same time and share its instruction-cache lines; the arithmetic per env is unchanged.
template <typename T, int NC, bool CTA_SYNC>
__device__ __noinline__ void solve_constraints(Ws<T, NC>& w, const DevModel<T>& m, T tol, bool active = true) {
  constexpr int NVV = Ws<T, NC>::NVV;
  constexpr int NENT = NVV * (NVV + 1) / 2, EPL = (NENT + 31) / 32;
  const int lane = LANE, nefc = active ? w.nefc : 0, ncon = w.ncon, nlim = w.nlim;
  T* warm = w.warm();
  bool done = !active;
  if (active && nefc == 0) {
    if (lane < NVV) { T a = w.qacc_smooth[lane]; w.qacc[lane] = a; warm[lane] = a; }
    if (lane == 0) w.diag[2] = 0;
    __syncwarp();
    done = true;
    if (!CTA_SYNC) return;
  }
  int d0 = 0, n = NVV, dof = lane, nent = 0, niter = 0;
  T scale = 0, cost = 0;
  int ea[EPL], eb[EPL];
#pragma unroll
  for (int k = 0; k < EPL; k++) { ea[k] = 0; eb[k] = 0; }
  if (!done) {
    bool arm = nlim > 0;
    for (int ci = lane; ci < ncon; ci += 32) arm |= (w.c_b1[ci] > 0 && w.c_b1[ci] < LCR_NABODY) || (w.c_b2[ci] > 0 && w.c_b2[ci] < LCR_NABODY);
    d0 = __any_sync(FULLMASK, arm) ? 0 : LCR_NARM; n = NVV - d0; dof = d0 + lane;
    nent = n * (n + 1) / 2;
    if (lane < d0) warm[lane] = w.qacc_smooth[lane];  // decoupled dofs take no warm start
    __syncwarp();
    // start from the warm start if it is cheaper than qacc_smooth; the FULL evaluation at the warm start is kept
    // when it wins (the common case for persistent contacts), so the chosen point is evaluated only once
    const T cs = total_cost<T, NC, false>(w, m, w.qacc_smooth, d0);
    const T cw = total_cost<T, NC, true>(w, m, warm, d0);
    const bool use_warm = cw < cs;
    if (lane < NVV) w.qacc[lane] = (use_warm && lane >= d0) ? warm[lane] : w.qacc_smooth[lane];
    __syncwarp();
    scale = 1 / (m.meaninertia * (T)NVV);
    cost = use_warm ? cw : total_cost<T, NC, true>(w, m, w.qacc, d0);
    // lower-triangle entries (island-local indices) owned by this lane
#pragma unroll
    for (int k = 0; k < EPL; k++) {
      int e = lane + 32 * k;
      if (e >= nent) e = 0;
      tri_index(e, ea[k], eb[k]);
    }
  }
  for (int iter = 0; iter < m.iterations; iter++) {
    if (CTA_SYNC) { if (!__syncthreads_or(!done)) break; }
    else if (done) break;
    if (done) continue;
    done = true;  // every `break` below ends this env's solve; cleared again at the bottom if it goes on
    do {
      // ---- Hessian H = M + sum_rows w_i J_i J_i^T + sum_cone-contacts (c1 G G^T - c2 P P^T)
      T h[EPL];
#pragma unroll
      for (int k = 0; k < EPL; k++) {
        const int a = d0 + ea[k], b = d0 + eb[k];
        T v = 0;
        if (a < LCR_NARM) v = w.M[a][b];
        else if (a == b) { const int d = a - LCR_NARM, c = d / 6; v = (d % 6) < 3 ? m.cube_mass[c] : m.cube_inertia[c]; }
        h[k] = v;
      }
#pragma unroll 4
      for (int i = 0; i < nefc; i++) {  // no skip of zero-weight rows: the loads of several rows stay in flight
        const T ww = w.e_w[i];
#pragma unroll
        for (int k = 0; k < EPL; k++) h[k] += ww * w.J[i][d0 + ea[k]] * w.J[i][d0 + eb[k]];
      }
      for (int ci = 0; ci < ncon; ci++) {
        const T c1 = w.c_c1[ci];
        if (c1 == 0) continue;  // warp-uniform
        const T c2 = w.c_c2[ci];
        const int i0 = w.c_efc[ci], dim = m.par(w.c_par[ci])->dim;
        T G = 0, P = 0;
        if (dof < NVV)
          for (int k = 0; k < dim; k++) { const T jk = w.J[i0 + k][dof]; G += w.e_g[i0 + k] * jk; P += w.e_p[i0 + k] * jk; }
#pragma unroll
        for (int k = 0; k < EPL; k++) {
          const T Ga = __shfl_sync(FULLMASK, G, ea[k]), Gb = __shfl_sync(FULLMASK, G, eb[k]);
          const T Pa = __shfl_sync(FULLMASK, P, ea[k]), Pb = __shfl_sync(FULLMASK, P, eb[k]);
          h[k] += c1 * Ga * Gb - c2 * Pa * Pb;
        }
      }
#pragma unroll
      for (int k = 0; k < EPL; k++)
        if (lane + 32 * k < nent) w.H[ea[k]][eb[k]] = h[k];
      __syncwarp();
      T s;
      {
        const T gi = lane < n ? w.grad[dof] : (T)0;
#ifndef LCR_NO_CHOL_REG
        if (n == 6) s = -chol_reg_factor_solve<T, 6>(&w.H[0][0], NVV + 1, gi);
        else if (n == 12) s = -chol_reg_factor_solve<T, 12>(&w.H[0][0], NVV + 1, gi);
        else
#endif
        { warp_cholesky(&w.H[0][0], NVV + 1, n); s = -warp_chol_solve(&w.H[0][0], NVV + 1, n, gi); }
      }
      if (lane < n) w.search[dof] = s;
      const T snorm = sqrt(warp_sum(lane < n ? s * s : (T)0));
      __syncwarp();
      if (snorm < c_minval<T>()) break;
      // ---- exact line search (safeguarded Newton on alpha)
      T mv = 0;
      if (lane < n) { mv = mul_M_dof(w, m, w.search, dof); w.Mv[dof] = mv; }
      for (int i = lane; i < nefc; i += 32) {
        T a = 0;
        for (int d = d0; d < NVV; d++) a += w.J[i][d] * w.search[d];
        w.e_jv[i] = a;
      }
      T g1 = 0, g2 = 0;
      if (lane < n) { g1 = s * (w.Ma[dof] - w.smooth[dof]); g2 = s * mv; }
      g1 = warp_sum(g1);
      g2 = warp_sum(g2);
      __syncwarp();
      const T gtol = tol * m.ls_tolerance * snorm / scale;
      T alpha = 0, lo = 0, hi = -1;
      for (int ls = 0; ls <= m.ls_iterations; ls++) {
        T d1 = 0, d2 = 0;
        if (lane < nlim) {
          const T jv = w.e_jv[lane], x = w.e_jar[lane] + alpha * jv;
          if (x < 0) { d1 = w.e_D[lane] * x * jv; d2 = w.e_D[lane] * jv * jv; }
        }
        for (int ci = lane; ci < ncon; ci += 32) {
          const Eval3<T> ev = contact_eval<T, NC, true, false>(w, m, ci, alpha);
          d1 += ev.d1; d2 += ev.d2;
        }
        d1 = warp_sum(d1) + g1 + alpha * g2;
        d2 = warp_sum(d2) + g2;
        if (fabs(d1) < gtol || ls == m.ls_iterations) break;
        if (d1 < 0) lo = alpha; else hi = alpha;
        T an = alpha - d1 / d2;
        if (hi >= 0 && (an <= lo || an >= hi)) an = (T)0.5 * (lo + hi);
        if (an == alpha) break;
        alpha = an;
      }
      if (!(alpha > 0)) break;
      if (lane < n) w.qacc[dof] += alpha * s;
      __syncwarp();
      const T old = cost;
      cost = total_cost<T, NC, true>(w, m, w.qacc, d0);
      niter = iter + 1;
      const T gn = sqrt(warp_sum(lane < n ? w.grad[dof] * w.grad[dof] : (T)0));
      if (scale * (old - cost) < tol || scale * gn < tol) break;
      done = false;
    } while (0);
  }
  if (active && nefc > 0) {
    if (lane < NVV) warm[lane] = w.qacc[lane];
    if (lane == 0) w.diag[2] = niter;
    __syncwarp();
  }
}

// ---------------------------------------------------------------- mj_forward / mj_step
template <typename T> DI T solver_tol(const DevModel<T>& m);
template <> DI double solver_tol<double>(const DevModel<double>& m) { return m.tolerance; }
template <> DI float solver_tol<float>(const DevModel<float>& m) { return fmaxf(m.tolerance, 1e-6f); }

template <typename T, int NC>
__device__ __noinline__ void forward(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts) {
#ifdef LCR_FLOW_DEBUG
  if (!Scene<NC>::BIG && LANE == 0 && gridDim.x == 148 && blockDim.x == 512) printf("forward() blk %d warp %d substep %d redo %d ints %d %d\n", blockIdx.x, threadIdx.x >> 5, w.substep, w.redo_forward, w.ints[0], w.ints[1]);
#endif
  kinematics(w, m);
  inertia_and_bias(w, m);
  make_constraints(w, m, verts, false);
  smooth_forces(w, m);
  solve_constraints<T, NC, false>(w, m, solver_tol<T>(m));
}

template <typename T, int NC> DI void reset_data(Ws<T, NC>& w, const DevModel<T>& m) {  // mj_resetData
  const int lane = LANE;
  for (int i = lane; i < Ws<T, NC>::NF - LCR_NAUX + 1; i += 32) w.st[i] = 0;  // qpos qvel ctrl warm time
  __syncwarp();
  if (lane < Scene<NC>::NCUBE) {
    T* qp = w.qpos() + LCR_NARM + 7 * lane;
    qp[0] = m.cube_qpos0[lane][0]; qp[1] = m.cube_qpos0[lane][1]; qp[2] = m.cube_qpos0[lane][2]; qp[3] = 1;
  }
  __syncwarp();
}

template <typename T> DI bool bad_val(T x) { return !(x == x) || x > (T)1e10 || x < (T)-1e10; }

// mj_checkPos / mj_checkVel: reset the env (mj_resetData) on NaN / huge values
template <typename T, int NC> DI void check_state(Ws<T, NC>& w, const DevModel<T>& m) {
  constexpr int NVV = Ws<T, NC>::NVV, NQ = Ws<T, NC>::NQ;
  const int lane = LANE;
  const bool bad = (lane < NQ && bad_val(w.qpos()[lane])) || (lane < NVV && bad_val(w.qvel()[lane]));
  if (__any_sync(FULLMASK, bad)) { reset_data(w, m); if (lane == 0) w.diag[5]++; }
}
// mj_checkAcc: returns true if qacc was bad and the env has been reset (caller must redo mj_forward)
template <typename T, int NC> DI bool check_acc(Ws<T, NC>& w, const DevModel<T>& m) {
  const int lane = LANE;
  const bool bad = lane < Ws<T, NC>::NVV && bad_val(w.qacc[lane]);
  if (__any_sync(FULLMASK, bad)) { reset_data(w, m); if (lane == 0) w.diag[5]++; return true; }
  return false;
}

// implicitfast velocity update + semi-implicit position update (mj_implicit + mj_advance)
template <typename T, int NC>
__device__ __noinline__ void integrate(Ws<T, NC>& w, const DevModel<T>& m) {
  constexpr int NVV = Ws<T, NC>::NVV;
  const int lane = LANE;
  T* qpos = w.qpos();
  T* qvel = w.qvel();
  const T h = m.timestep;
  // (M + h diag(damping + kv)) a = M qacc on the arm block; cubes keep qacc
  T rhs = mul_M(w, m, w.qacc);
  T a;
  {
    T ar[LCR_NARM], cr[LCR_NARM];
#pragma unroll
    for (int k = 0; k < LCR_NARM; k++) { ar[k] = (lane < LCR_NARM && k <= lane) ? w.M[lane][k] : (T)0; cr[k] = 0; }
    if (lane < LCR_NARM) {
      const T dd = h * (m.jnt_damping[lane] + m.act_kv[lane]);
#pragma unroll
      for (int k = 0; k < LCR_NARM; k++) if (k == lane) ar[k] += dd;
    }
    chol_reg<T, LCR_NARM>(ar, cr);
    a = chol_reg_solve<T, LCR_NARM>(ar, cr, rhs);
  }
  if (lane >= LCR_NARM && lane < NVV) a = w.qacc[lane];
  if (lane < NVV) qvel[lane] += h * a;
  __syncwarp();
  if (lane < LCR_NARM) qpos[lane] += h * qvel[lane];
  else if (lane < LCR_NARM + Scene<NC>::NCUBE) {
    const int c = lane - LCR_NARM;
    T* qp = qpos + LCR_NARM + 7 * c;
    const T* qv = qvel + LCR_NARM + 6 * c;
    qp[0] += h * qv[0]; qp[1] += h * qv[1]; qp[2] += h * qv[2];
    T q[4] = {qp[3], qp[4], qp[5], qp[6]};
    quat_normalize(q);
    const T wn = sqrt(qv[3] * qv[3] + qv[4] * qv[4] + qv[5] * qv[5]);
    if (wn * h > 0) {
      T sn, cs;
      sincos((T)0.5 * wn * h, &sn, &cs);
      sn /= wn;
      T dq[4] = {cs, sn * qv[3], sn * qv[4], sn * qv[5]}, r[4];
      quat_mul(r, q, dq);
      quat_normalize(r);
      q[0] = r[0]; q[1] = r[1]; q[2] = r[2]; q[3] = r[3];
    }
    qp[3] = q[0]; qp[4] = q[1]; qp[5] = q[2]; qp[6] = q[3];
  }
  if (lane == 0) w.aux()[0] += h;
  __syncwarp();
}

template <typename T, int NC>
__device__ __noinline__ void substep(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts) {
  check_state(w, m);
  forward(w, m, verts);
  if (check_acc(w, m)) forward(w, m, verts);
  integrate(w, m);
}

// ---------------------------------------------------------------- env glue
// width of an observation row (get_observation of the six envs; lcr_obs_dim of the C-ABI)
DI int obs_dim(int task) { return (task == LCR_TASK_REACH || task == LCR_TASK_LIFT || task == LCR_TASK_PUSH_LOOP) ? 15 : 18; }
template <typename T, int NC> DI void write_obs(Ws<T, NC>& w, const DevModel<T>& m, float* obs, float* rec = nullptr) {
  const int lane = LANE, task = m.task;
  const T* qpos = w.qpos();
  const T* qvel = w.qvel();
  const bool has_target = task == LCR_TASK_PUSH || task == LCR_TASK_PICK_PLACE;
  const int od = obs_dim(task);
  if (lane < od) {
    T v;
    if (lane < 6) v = qpos[lane];
    else if (lane < 12) v = qvel[lane - 6];
    else if (lane < 15) v = has_target ? w.target()[lane - 12] : qpos[6 + lane - 12];
    else v = has_target ? qpos[6 + lane - 15] : qpos[13 + lane - 15];
    if (obs) obs[lane] = (float)v;
    if (rec) rec[lane] = (float)v;
  }
}
// observation row + the scalars of one env (lane 0 holds reward / flags)
template <typename T, int NC> DI void write_outputs(Ws<T, NC>& w, const DevModel<T>& m, const StepIO& io, int env, float r, bool te, bool tr, bool su) {
  const int od = obs_dim(m.task);
  float* rec = io.rec ? io.rec + (size_t)env * (od + 4) : nullptr;
  write_obs(w, m, io.obs + (size_t)env * od, rec);
  if (LANE == 0) {
    io.reward[env] = r; io.term[env] = te; io.trunc[env] = tr; io.succ[env] = su;
    if (rec) { rec[od] = r; rec[od + 1] = te ? 1.0f : 0.0f; rec[od + 2] = tr ? 1.0f : 0.0f; rec[od + 3] = su ? 1.0f : 0.0f; }
  }
}

template <typename T, int NC>
__device__ __noinline__ void env_reset(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts) {
  const int lane = LANE, task = m.task;
  T* qpos = w.qpos();
  if (lane == 0) {
    for (int j = 0; j < 6; j++) qpos[j] = 0;
    if (task == LCR_TASK_PUSH_LOOP) {
      // push_cube_loop_env.py:302-320: the cube is drawn inside the region of the CURRENT goal (target[0], persists across resets)
      const int cg = w.target()[0] != (T)0;
      T* qp = qpos + 6;
      for (int k = 0; k < 3; k++) {
        const double hi = m.goal_high[k], lo = k < 2 ? hi * -1.0 : hi;
        const double p = lo + (hi - lo) * pcg64_double(w.rng);
        qp[k] = (T)(k < 2 ? p + m.goal_center[cg][k] : p);
      }
      qp[3] = 1; qp[4] = qp[5] = qp[6] = 0;
    } else
    for (int c = 0; c < Scene<NC>::NCUBE; c++) {
      T* qp = qpos + 6 + 7 * c;
      for (int k = 0; k < 3; k++) qp[k] = (T)(m.cube_low[k] + (m.cube_high[k] - m.cube_low[k]) * pcg64_double(w.rng));
      qp[3] = 1; qp[4] = qp[5] = qp[6] = 0;
    }
    if (task == LCR_TASK_PUSH || task == LCR_TASK_PICK_PLACE)
      for (int k = 0; k < 3; k++)
        w.target()[k] = (T)(float)(m.target_low[k] + (m.target_high[k] - m.target_low[k]) * pcg64_double(w.rng));
    w.ints[0] = 0;
    w.ints[1] = 0;
  }
  __syncwarp();
  forward(w, m, verts);
}

// one damped-least-squares update of the IK (reach_cube_env.py:191-218) from the kinematics in `w`: returns true (q untouched) if the site
// is within 1 cm of the target, otherwise advances q (lane a < 6 holds joint a)
template <typename T, int NC>
DI bool ik_update(Ws<T, NC>& w, const DevModel<T>& m, const T* target, T& q) {
  const int lane = LANE;
  const T* site = w.site_xpos();
  T err[3] = {target[0] - site[0], target[1] - site[1], target[2] - site[2]};
  if (sqrt(dot3(err, err)) < (T)0.01) return true;
  // lane a < 6: Jacobian column a, A[a][:] = J^T J + 0.15 I, rhs_a = J^T err
  T col[3] = {0, 0, 0};
  if (lane < m.site_body && lane < 6) {
    T r[3] = {site[0] - w.xpos[lane + 1][0], site[1] - w.xpos[lane + 1][1], site[2] - w.xpos[lane + 1][2]};
    cross3(col, w.axis[lane], r);
  }
  T rhs = dot3(col, err);
  T ar[LCR_NARM], cr[LCR_NARM];
#pragma unroll
  for (int b = 0; b < 6; b++) {
    T cb[3] = {__shfl_sync(FULLMASK, col[0], b), __shfl_sync(FULLMASK, col[1], b), __shfl_sync(FULLMASK, col[2], b)};
    ar[b] = (lane < 6 && b <= lane) ? dot3(col, cb) + (lane == b ? (T)0.15 : (T)0) : (T)0;
    cr[b] = 0;
  }
  chol_reg<T, LCR_NARM>(ar, cr);
  T qd = chol_reg_solve<T, LCR_NARM>(ar, cr, rhs);
  const T n = sqrt(warp_sum(lane < 6 ? qd * qd : (T)0));
  if (n > 1) qd /= n;
  if (lane < 6) q = clampT(q + (T)0.5 * qd, m.jnt_range[lane][0], m.jnt_range[lane][1]);
  return false;
}

// damped least squares IK (reach_cube_env.py:148-221); teleport = faithful in-step behaviour
template <typename T, int NC>
__device__ __noinline__ void inverse_kinematics(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, const T* target,
                                                T* q_out, bool teleport) {
  const int lane = LANE;
  T* qpos = w.qpos();
  T q = lane < 6 ? qpos[lane] : (T)0;
  const T save = q;
  for (int it = 0; it < 10; it++) {
    if (lane < 6) qpos[lane] = q;
    __syncwarp();
    if (teleport) forward(w, m, verts); else kinematics(w, m);
    if (ik_update(w, m, target, q)) break;
    __syncwarp();
  }
  if (lane < 6) q_out[lane] = q;
  if (!teleport) {
    if (lane < 6) qpos[lane] = save;
    __syncwarp();
  }
  __syncwarp();
}

__device__ const double kTargetLow[6] = {-3.14159, -1.5708, -1.48353, -1.91986, -2.96706, -1.74533};
__device__ const double kTargetHigh[6] = {3.14159, 1.22173, 1.74533, 1.91986, 2.96706, 0.0523599};

// env_step_begin in two halves around the IK of the ee action mode, so that a CTA that owns ONE env (the BIG passes) can run the IK's
// mj_forward passes with its helper warps (env_step_begin_cta); env_step_begin = the one-warp composition, same arithmetic.
// step_pre: 0 = the env was auto-reset instead of stepped (outputs written), 1 = joint mode, ctrl set, 2 = ee mode: IK target in w.search[0..2]
template <typename T, int NC>
__device__ __noinline__ int step_pre(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, const StepIO& io, int env) {
  const int lane = LANE, task = m.task;
  if (m.autoreset && w.ints[1]) {
    env_reset(w, m, verts);
    if (Scene<NC>::BIG || !w.ovf) write_outputs(w, m, io, env, 0.0f, false, false, false);
    __syncwarp();
    return 0;
  }
  if (lane == 0) w.diag[3] = 0;
  const int na = (m.action_mode ? 3 : 5) + (m.block_gripper ? 0 : 1);
  const float* action = io.actions + (size_t)env * na;
  const bool gripper_task = task == LCR_TASK_LIFT || task == LCR_TASK_PICK_PLACE || task == LCR_TASK_STACK;
  T* qpos = w.qpos();
  T a = lane < na ? clampT((T)action[lane], (T)-1, (T)1) : (T)0;
  if (m.action_mode == 1) {
    T* sc = w.search;  // free before the first and after the last mj_forward of the IK loop
    const T* site = w.site_xpos();
    // `ee_action * 0.05` is a float32 product in the reference (float32 array times a Python float), rounded before the sum
    if (lane < 3) { T t = site[lane] + (T)__fmul_rn((float)a, 0.05f); if (lane == 2 && t < 0) t = 0; sc[lane] = t; }
    __syncwarp();
    return 2;
  }
  T tq = 0;
  const T alast = __shfl_sync(FULLMASK, a, na - 1);
  if (lane < 5) tq = clampT(a + qpos[lane], (T)kTargetLow[lane], (T)kTargetHigh[lane]);
  else if (lane == 5) tq = gripper_task ? clampT(alast + qpos[5], (T)kTargetLow[5], (T)kTargetHigh[5]) : (T)0;
  if (lane < 6) w.ctrl()[lane] = tq;
  __syncwarp();
  return 1;
}
// ee mode, after the IK: q (lane a < 6 = joint a) -> ctrl, gripper from the action (lift_cube_env.py:253-257)
template <typename T, int NC>
DI void step_post_ik(Ws<T, NC>& w, const DevModel<T>& m, const StepIO& io, int env, T q) {
  const int lane = LANE, task = m.task;
  const int na = 3 + (m.block_gripper ? 0 : 1);
  const float* action = io.actions + (size_t)env * na;
  const bool gripper_task = task == LCR_TASK_LIFT || task == LCR_TASK_PICK_PLACE || task == LCR_TASK_STACK;
  const T a = lane < na ? clampT((T)action[lane], (T)-1, (T)1) : (T)0;
  T tq = lane < 6 ? q : (T)0;
  if (lane == 5 && !gripper_task) tq = 0;
  const T ga = na > 3 ? __shfl_sync(FULLMASK, a, 3) : (T)0;
  if (lane == 5 && gripper_task) tq = clampT(w.qpos()[5] + (T)__fmul_rn((float)ga, 0.2f), m.act_ctrlrange[5][0], m.act_ctrlrange[5][1]);
  if (lane < 6) w.ctrl()[lane] = tq;
  __syncwarp();
}

template <typename T, int NC>
__device__ __noinline__ bool env_step_begin(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, const StepIO& io, int env) {
  // returns false if the env was auto-reset instead of stepped (outputs already written)
  const int pre = step_pre(w, m, verts, io, env);
  if (pre != 2) return pre != 0;
  T* sc = w.search;
  T tgt[3] = {sc[0], sc[1], sc[2]};
  __syncwarp();
  inverse_kinematics(w, m, verts, tgt, sc, true);
  __syncwarp();
  const T q = LANE < 6 ? sc[LANE] : (T)0;
  step_post_ik(w, m, io, env, q);
  return true;
}

// get_reward / get_cube_overlap of PushCubeLoop (push_cube_loop_env.py:337-383).  The reference mixes numpy float32
// (cube_position), numpy float64 (model arrays) and Python scalars; the dtype of every intermediate follows NumPy >= 2
// promotion (NEP 50: a Python scalar adopts the dtype of the numpy operand), tracked explicitly like in the oracle.
// Correctly rounded intrinsics: the float32 build compiles with approximate division and with FMA contraction.
struct PyNum { double v; int k; };  // k: 0 numpy float64, 1 numpy float32, 2 Python scalar (weak)
DI PyNum pn(double v, int k) { PyNum r; r.v = v; r.k = k; return r; }
DI PyNum pn_op(PyNum a, PyNum b, int op) {  // 0 +, 1 -, 2 *, 3 /
  const int k = (a.k == 0 || b.k == 0) ? 0 : ((a.k == 1 || b.k == 1) ? 1 : 2);
  if (k == 1) {
    const float x = (float)a.v, y = (float)b.v;
    const float z = op == 0 ? __fadd_rn(x, y) : op == 1 ? __fsub_rn(x, y) : op == 2 ? __fmul_rn(x, y) : __fdiv_rn(x, y);
    return pn((double)z, 1);
  }
  return pn(op == 0 ? __dadd_rn(a.v, b.v) : op == 1 ? __dsub_rn(a.v, b.v) : op == 2 ? __dmul_rn(a.v, b.v) : __ddiv_rn(a.v, b.v), k);
}
DI PyNum pn_min(PyNum a, PyNum b) { return b.v < a.v ? b : a; }  // Python min / max return the first extremal operand
DI PyNum pn_max(PyNum a, PyNum b) { return b.v > a.v ? b : a; }

// lane 0 only; returns success, flips the current goal (target[0]) on success
template <typename T, int NC>
__device__ __noinline__ bool loop_reward(Ws<T, NC>& w, const DevModel<T>& m, float& reward) {
  const T* qpos = w.qpos();
  const int cg = w.target()[0] != (T)0;
  const PyNum half = pn(0.015 / 2, 2);  // self.cube_size (push_cube_loop_env.py:124)
  PyNum ov[2];
#pragma unroll 1
  for (int k = 0; k < 2; k++) {
    const PyNum c = pn((double)(float)qpos[6 + k], 1);  // data.qpos[6:9].astype(np.float32)
    const PyNum g = pn(m.goal_center[cg][k], 0), wg = pn(m.goal_high[k], 0);
    const PyNum up = pn_min(pn_op(c, half, 0), pn_op(g, wg, 0)), dn = pn_max(pn_op(c, half, 1), pn_op(g, wg, 1));
    ov[k] = pn_max(pn(0, 2), pn_op(up, dn, 1));
  }
  const PyNum overlap = pn_op(pn_op(ov[0], ov[1], 2), pn(0.0075 * 0.0075 * 4, 2), 3);
  if (overlap.v > 0.95) {
    reward = 5.0f;
    w.target()[0] = cg ? (T)0 : (T)1;
    return true;
  }
  if (overlap.v > 0.0) { reward = (float)pn_op(overlap, pn(1, 2), 1).v; return false; }
  const double edge = __dadd_rn(m.goal_high[1] * -1.0, m.goal_center[cg][1]);
  const double diff = __dsub_rn((double)(float)qpos[7], edge), dist = sqrt(__dmul_rn(diff, diff));
  double r = __dsub_rn(__ddiv_rn(-dist, 0.16), 1.0);
  r = r > -2 ? r : -2;
  r = -1 < r ? -1 : r;
  reward = (float)r;
  return false;
}

template <typename T, int NC>
__device__ __noinline__ void env_step_end(Ws<T, NC>& w, const DevModel<T>& m, const StepIO& io, int env) {
  const int task = m.task;
  float r = 0;
  bool te = false, tr = false, su = false;
  if (LANE == 0) {
    if (task == LCR_TASK_PUSH_LOOP) {  // push_cube_loop_env.py:322-335: never terminates; TimeLimit truncates
      su = loop_reward(w, m, r);
    } else {
      T pa[3], pb[3];
      const T* site = w.site_xpos();
      const T* c0 = w.cube_xpos(0);
      if (task == LCR_TASK_REACH || task == LCR_TASK_LIFT) { for (int k = 0; k < 3; k++) { pa[k] = site[k]; pb[k] = c0[k]; } }
      else if (task == LCR_TASK_STACK) { const T* c1 = w.cube_xpos(Scene<NC>::NCUBE - 1); for (int k = 0; k < 3; k++) { pa[k] = c1[k]; pb[k] = c0[k]; } pb[2] += (T)0.03; }
      else { for (int k = 0; k < 3; k++) { pa[k] = c0[k]; pb[k] = w.target()[k]; } }
      T dx = pa[0] - pb[0], dy = pa[1] - pb[1], dz = pa[2] - pb[2];
      T d = sqrt(dx * dx + dy * dy + dz * dz);
      if (task == LCR_TASK_LIFT) r = (float)((c0[2] - m.height_threshold) + d);
      else {
        su = d < m.distance_threshold;
        te = su;
        r = m.reward_type == 0 ? -(float)(d > m.distance_threshold) : (float)(-d);
      }
    }
    const int el = ++w.ints[0];
    tr = m.max_episode_steps > 0 && el >= m.max_episode_steps;
    w.ints[1] = (te || tr) ? 1 : 0;
  }
  write_outputs(w, m, io, env, r, te, tr, su);
  __syncwarp();
}

DI void redo_push(const Redo& rd, int env) {
  if (LANE == 0) rd.list[atomicAdd(rd.count, 1)] = env;
}
// true if this env must leave the fast path (never on the BIG path: what exceeds the big caps is dropped and counted)
template <typename T, int NC> DI bool moved(const Ws<T, NC>& w) { return !Scene<NC>::BIG && w.ovf != 0; }

template <typename T, int NC>
DI void env_step(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, const StepIO& io, int env) {
  if (!env_step_begin(w, m, verts, io, env)) return;
#pragma unroll 1
  for (int k = 0; k < m.n_substeps; k++) substep(w, m, verts);
  // (an env that outgrew the fast workspace is redone over the big one: no outputs from this pass)
  if (!moved(w)) env_step_end(w, m, io, env);
}

// ---------------------------------------------------------------- state staging HBM <-> shared
// One warp moves its env's record with coalesced 128-bit accesses: lane l <-> uint4 #l.
template <typename T, int NC>
DI void load_state(Ws<T, NC>& w, const DevState<T>& s, int env) {
  constexpr int NV4 = Ws<T, NC>::NFP * (int)sizeof(T) / 16;
  const uint4* src = reinterpret_cast<const uint4*>(s.st + (size_t)env * Ws<T, NC>::NFP);
  uint4* dst = reinterpret_cast<uint4*>(w.st);
  for (int i = LANE; i < NV4; i += 32) dst[i] = src[i];
  if (LANE < LCR_IB_WORDS / 4)
    reinterpret_cast<uint4*>(w.ints)[LANE] = reinterpret_cast<const uint4*>(s.ib + (size_t)env * LCR_IB_WORDS)[LANE];
  constexpr int NSA4 = Ws<T, NC>::SA_BYTES / 16;
  const uint4* sas = reinterpret_cast<const uint4*>(s.sa + (size_t)env * Ws<T, NC>::SA_BYTES);
  uint4* sad = reinterpret_cast<uint4*>(w.sa_dir);
  for (int i = LANE; i < NSA4; i += 32) sad[i] = sas[i];
  if (LANE == 0) { w.ovf = 0; w.substep = 0; w.jobs_left = 0; }
  __syncwarp();
}
template <typename T, int NC>
DI void store_state(Ws<T, NC>& w, const DevState<T>& s, int env) {
  constexpr int NV4 = Ws<T, NC>::NFP * (int)sizeof(T) / 16;
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(s.st + (size_t)env * Ws<T, NC>::NFP);
  const uint4* src = reinterpret_cast<const uint4*>(w.st);
  for (int i = LANE; i < NV4; i += 32) dst[i] = src[i];
  if (LANE < LCR_IB_WORDS / 4)
    reinterpret_cast<uint4*>(s.ib + (size_t)env * LCR_IB_WORDS)[LANE] = reinterpret_cast<const uint4*>(w.ints)[LANE];
  constexpr int NSA4 = Ws<T, NC>::SA_BYTES / 16;
  uint4* sad = reinterpret_cast<uint4*>(s.sa + (size_t)env * Ws<T, NC>::SA_BYTES);
  const uint4* sas = reinterpret_cast<const uint4*>(w.sa_dir);
  for (int i = LANE; i < NSA4; i += 32) sad[i] = sas[i];
}

// ---------------------------------------------------------------- kernels (one warp per env)
extern __shared__ __align__(16) unsigned char lcr_smem[];

// env handled by CTA `i` of a fused kernel: identity, or entry i of a device-side list (the BIG redo pass)
DI int list_env(const int* __restrict__ list, int i) { return list ? list[i] : i; }

// One CTA = one warp = one env, the whole step in one launch.  With `list` the grid strides over a device-side env list
// (BIG pass over the envs that outgrew the fast workspace).  A fast env that hits a cap is not written back: redo list.
template <typename T, int NC>
__global__ void __launch_bounds__(32, Scene<NC>::BIG ? 2 : 16) k_step(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                                                      StepIO io, Redo redo, const int* __restrict__ list, const int* __restrict__ count, int first) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const DevModel<T>& m = *dm;
  const int n = list ? *count : s.n;
  for (int i = first + blockIdx.x; i < n; i += gridDim.x) {
    const int env = list_env(list, i);
    load_state(w, s, env);
    env_step(w, m, verts, io, env);
    if (moved(w)) redo_push(redo, env);
    else store_state(w, s, env);
    __syncwarp();
  }
}

template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_reset(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                              const uint8_t* __restrict__ mask, float* __restrict__ obs) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = blockIdx.x;
  if (mask != nullptr && !mask[env]) return;
  load_state(w, s, env);
  const DevModel<T>& m = *dm;
  const int od = obs_dim(m.task);
  env_reset(w, m, verts);
  if (obs) write_obs(w, m, obs + (size_t)env * od);
  store_state(w, s, env);
}

template <typename T, int NC>
__global__ void __launch_bounds__(32, Scene<NC>::BIG ? 2 : 16) k_substeps(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s, int nsub,
                                                                          Redo redo, const int* __restrict__ list, const int* __restrict__ count) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int n = list ? *count : s.n;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const int env = list_env(list, i);
    load_state(w, s, env);
    if (LANE == 0) w.diag[3] = 0;
    if (nsub == 0) forward(w, *dm, verts);
    for (int k = 0; k < nsub; k++) substep(w, *dm, verts);
    if (moved(w)) redo_push(redo, env);
    else store_state(w, s, env);
    __syncwarp();
  }
}

template <typename T, int NC>
__global__ void __launch_bounds__(32) k_ik(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                           const float* __restrict__ target, float* __restrict__ q_out) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  T* outq = w.search;
  const int env = blockIdx.x, lane = LANE;
  load_state(w, s, env);
  T tgt[3] = {(T)target[3 * (size_t)env], (T)target[3 * (size_t)env + 1], (T)target[3 * (size_t)env + 2]};
  inverse_kinematics(w, *dm, verts, tgt, outq, false);
  __syncwarp();
  if (lane < 6) q_out[6 * (size_t)env + lane] = (float)outq[lane];
  // state is not written back: lcr_ik leaves the simulation untouched
}

// ---------------------------------------------------------------- lockstep execution
// One CTA = W warps = W envs that walk mj_step phase by phase TOGETHER.  The fused one-warp-per-CTA kernel is
// instruction-fetch bound (ncu: `no_instruction` = 17 of the 24 cycles between two issues of a warp): a substep
// is ~110 KB of SASS against a 32 KB L1.5 instruction cache, and 14 independent warps per SM each sit in a
// different place of it.  Here the warps of a CTA are re-aligned by CTA barriers at the phase boundaries (and
// optionally between Newton iterations), so a cache line fetched for one warp is hit by the others; the
// narrowphase jobs of all W envs go into one CTA-wide pool that any warp drains (the workspaces are in shared
// memory, so a warp can run another env's MPR), which evens out the most unevenly distributed work of the step.
// Per-env arithmetic is identical to k_step (bitwise identical state and outputs).
#define LCR_LS_BAR_TOP 1      // barrier at the top of every substep
#define LCR_LS_BAR_CON 2      // barrier before the constraint build (implied by the job pool)
#define LCR_LS_BAR_SOL 4      // barrier before the solver
#define LCR_LS_SYNC_NEWTON 8  // barriers between Newton iterations
#define LCR_LS_JOB_POOL 16    // CTA-wide narrowphase job pool
#define LCR_LS_RESUME 32      // BIG pass over migrated envs: seats hold env | substep << 24, the env resumes from its parked substep-start state

template <typename T, int NC>
DI void cta_jobs(Ws<T, NC>* wsa, int W, const DevModel<T>& m, const T* __restrict__ verts, int* next) {
  constexpr int MC = Ws<T, NC>::MAXCAND;
  int tot = 0;
  for (int e = 0; e < W; e++) { const int c = wsa[e].ncand; tot += c < MC ? c : MC; }
  for (;;) {
    int j = 0;
    if (LANE == 0) j = atomicAdd(next, 1);
    j = __shfl_sync(FULLMASK, j, 0);
    if (j >= tot) break;
    int e = 0;
    for (;; e++) {
      int c = wsa[e].ncand;
      c = c < MC ? c : MC;
      if (j < c) break;
      j -= c;
    }
    Ws<T, NC>& we = wsa[e];
    T r[8];
    narrowphase_job<T, NC, true>(we, m, verts, we.cand_key[j], r);
    __syncwarp();
    if (LANE < 8) cand_res(we)[j][LANE] = r[LANE];
  }
}

// Work-aware env order for the lockstep kernel.  Envs are ranked by the constraint count they reached in their
// previous step (diag[3] = max nefc; contacts persist, so it predicts the cost of the next step; envs that will only
// be auto-reset are the cheapest), most expensive first, and seated in one of two ways:
//  * STRIPED (few envs per SM: the step time is the chain of the most expensive env): rank r goes to CTA r % nCTA,
//    seat r / nCTA -- every CTA gets one env of each cost stratum, so an expensive env sits with cheap ones whose
//    warps drain its narrowphase jobs, and CTAs are ordered by their most expensive member (the grid tail is cheap);
//  * SORTED (many waves of CTAs: the step time is the total work): rank r takes seat r -- CTAs hold envs of equal
//    cost, so nobody waits at the phase barriers for a much slower neighbour, and the few all-expensive CTAs start
//    first and overlap the cheap ones.
// One CTA, two passes over the 64-byte int records; the order inside a bucket is arbitrary and does not affect any
// result.  Empty seats (n not a multiple of W) hold -1.
// Envs whose previous step needed the BIG workspace (diag[3] >= tbig rows) are not seated at all: they go to the list
// `big` (big[0] = count, then env ids) and start over the big workspace right away, on a stream of their own, instead of
// being found out and redone after the main kernels (contacts persist: an env beyond the fast caps usually stays there).
#define LCR_NBUCKET 16
template <typename T>
__global__ void __launch_bounds__(1024) k_sched(DevState<T> s, int* __restrict__ perm, int W, int striped, int* __restrict__ big, int tbig) {
  __shared__ int hist[LCR_NBUCKET], start[LCR_NBUCKET];
  const int ncta = (s.n + W - 1) / W;
  if (threadIdx.x < LCR_NBUCKET) hist[threadIdx.x] = 0;
  for (int k = threadIdx.x; k < ncta * W; k += blockDim.x) perm[k] = -1;
  __syncthreads();
  for (int e = threadIdx.x; e < s.n; e += blockDim.x) {
    const int32_t* ib = s.ib + (size_t)e * LCR_IB_WORDS;
    int key = ib[1] ? 0 : 1 + ib[LCR_NINT + 3] / 8;
    key = key < LCR_NBUCKET - 1 ? key : LCR_NBUCKET - 2;
    if (big && !ib[1] && ib[LCR_NINT + 3] >= tbig) key = LCR_NBUCKET - 1;
    atomicAdd(&hist[key], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int k = LCR_NBUCKET - 1; k >= 0; k--) { start[k] = acc; acc += hist[k]; }
    if (big) big[0] = hist[LCR_NBUCKET - 1];
  }
  __syncthreads();
  const int nbig = big ? hist[LCR_NBUCKET - 1] : 0;
  for (int e = threadIdx.x; e < s.n; e += blockDim.x) {
    const int32_t* ib = s.ib + (size_t)e * LCR_IB_WORDS;
    int key = ib[1] ? 0 : 1 + ib[LCR_NINT + 3] / 8;
    key = key < LCR_NBUCKET - 1 ? key : LCR_NBUCKET - 2;
    const bool isbig = big && !ib[1] && ib[LCR_NINT + 3] >= tbig;
    if (isbig) key = LCR_NBUCKET - 1;
    int r = atomicAdd(&start[key], 1);
    if (isbig) { big[1 + r] = e; continue; }
    r -= nbig;
    perm[striped ? (r % ncta) * W + r / ncta : r] = e;
  }
}

// env_step_begin for a CTA that owns ONE env (the BIG passes: warp 0 owns the env, the others help): the same arithmetic, but the
// narrowphase jobs of the IK's mj_forward passes -- dozens of MPR runs per pass in the deeply penetrating poses that need the big
// workspace, ten passes per step in the ee action mode -- are drained by all warps of the CTA instead of the owner alone.
// Every warp of the CTA calls this; returns env_step_begin's value in the owner warp, false in the helpers.
template <typename T, int NC>
__device__ bool env_step_begin_cta(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, const StepIO& io, int env, bool owner, int* job_next) {
  __shared__ int flag;
  if (owner) {
    const int pre = step_pre(w, m, verts, io, env);
    if (LANE == 0) flag = pre;
  }
  __syncthreads();
  const int pre = flag;
  if (pre != 2) return owner && pre != 0;
  T tgt[3] = {0, 0, 0}, q = 0;
  if (owner) {
    tgt[0] = w.search[0]; tgt[1] = w.search[1]; tgt[2] = w.search[2];
    q = LANE < 6 ? w.qpos()[LANE] : (T)0;
  }
  for (int it = 0; it < 10; it++) {
    if (owner) {  // mj_forward, first half
      if (LANE < 6) w.qpos()[LANE] = q;
      __syncwarp();
      kinematics(w, m);
      inertia_and_bias(w, m);
      collect_candidates(w, m);
      if (LANE == 0) *job_next = 0;
    }
    __syncthreads();
    cta_jobs(&w, 1, m, verts, job_next);
    __syncthreads();
    if (owner) {  // second half, then the IK update
      make_constraints(w, m, verts, true);
      smooth_forces(w, m);
      solve_constraints<T, NC, false>(w, m, solver_tol<T>(m));
      const bool converged = ik_update(w, m, tgt, q);
      __syncwarp();
      if (LANE == 0) flag = converged ? 0 : 1;
    }
    __syncthreads();
    if (!flag) break;
  }
  if (owner) step_post_ik(w, m, io, env, q);
  return owner;
}

template <typename T, int NC, bool PROF>
__global__ void __launch_bounds__(512, 1) k_step_ls(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s, StepIO io, Redo redo,
                                                    int flags, const int* __restrict__ perm, int epc, long long* __restrict__ prof,
                                                    const int* __restrict__ count, const void* __restrict__ parked = nullptr) {
  // blockDim.x / 32 warps, the first `epc` of them own an env (seat blockIdx.x * epc + warp), the others only help
  // with narrowphase jobs; shared memory holds epc workspaces
  Ws<T, NC>* wsa = reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  __shared__ int job_next;
  const int W = epc, warp = threadIdx.x >> 5, slot = blockIdx.x * epc + warp;
  const bool owner = warp < epc;
  // (count, optional: the seats are a device-side list of that many envs -- the BIG pass; seats beyond it are empty)
  int env = owner && !(count != nullptr && slot >= *count) ? (perm != nullptr ? perm[slot] : slot) : -1;  // -1 = empty seat
  int k0 = 0;  // first substep to run (LCR_LS_RESUME: the substep whose constraint rows outgrew the fast workspace)
  if ((flags & LCR_LS_RESUME) && env >= 0) { k0 = (env >> 24) & 127; env &= 0xffffff; }
  const bool valid = env >= 0 && env < s.n;
  if (!__syncthreads_or(valid)) return;
  Ws<T, NC>& w = wsa[owner ? warp : 0];
  const DevModel<T>& m = *dm;
  // optional per-env phase timing (debug hook, prof == nullptr in production): clock64 deltas summed over the substeps
  long long tp[PROF ? 10 : 1] = {0}, t0 = 0;
#define LCR_TICK(k) do { if (PROF) { (void)*(volatile int*)&job_next; /* BAR.SYNC defers blocking to the next memory access */ \
    const long long t1_ = clock64(); tp[k] += t1_ - t0; t0 = t1_; } } while (0)
  if (PROF) t0 = clock64();
  bool go = false;
  // one env per CTA with helper warps (the BIG passes): the step begin (IK forward passes of the ee mode) is CTA-cooperative too
  const bool coop_begin = Scene<NC>::BIG && W == 1 && (flags & LCR_LS_JOB_POOL) && !(flags & LCR_LS_RESUME);
  if (valid) {
    if (flags & LCR_LS_RESUME) {
      // migrated from the phased chain: the state block (state, ints, diag, rng) of the parked fast workspace is the start state of
      // substep k0 and has the same layout in both workspaces; the separating-axis cache comes along (performance only)
      typedef Ws<T, (NC & 15)> WsF;
      typedef Ws<T, NC> WsN;
      static_assert(offsetof(WsF, xpos) == offsetof(WsN, xpos), "state block layout");
      const unsigned char* src = reinterpret_cast<const unsigned char*>(reinterpret_cast<const WsF*>(parked) + env);
      for (int i = LANE; i < (int)offsetof(WsF, xpos) / 16; i += 32) reinterpret_cast<uint4*>(&w)[i] = reinterpret_cast<const uint4*>(src)[i];
      for (int i = LANE; i < WsF::SA_BYTES / 16; i += 32)
        reinterpret_cast<uint4*>(w.sa_dir)[i] = reinterpret_cast<const uint4*>(src + offsetof(WsF, sa_dir))[i];
      if (LANE == 0) { w.ovf = 0; w.ncand = 0; w.redo_forward = 0; w.substep = 0; w.jobs_left = 0; }
      __syncwarp();
      go = true;
    } else {
      load_state(w, s, env);
      if (!coop_begin) go = env_step_begin(w, m, verts, io, env);
    }
  }
  if (coop_begin) go = env_step_begin_cta(w, m, verts, io, env, owner, &job_next);
  if (!go && owner) { if (LANE == 0) w.ncand = 0; __syncwarp(); }
  const T tol = solver_tol<T>(m);
  LCR_TICK(0);
  if (flags & LCR_LS_RESUME) {  // (one env per CTA: its helper warps start at the same substep)
    __shared__ int k0_cta;
    if (threadIdx.x == 0) k0_cta = k0;
    __syncthreads();
    k0 = k0_cta;
  }
#pragma unroll 1
  for (int k = k0; k < m.n_substeps; k++) {
    if (flags & LCR_LS_BAR_TOP) __syncthreads();
    LCR_TICK(1);  // wait at the top of the substep
    if (go) { check_state(w, m); kinematics(w, m); inertia_and_bias(w, m); }
    if (flags & LCR_LS_JOB_POOL) {
      if (go) collect_candidates(w, m);
      if (threadIdx.x == 0) job_next = 0;
      LCR_TICK(2);  // kinematics, inertia, broadphase
      __syncthreads();
      LCR_TICK(3);  // wait before the job pool
      cta_jobs(wsa, W, m, verts, &job_next);
      LCR_TICK(4);  // narrowphase jobs
      __syncthreads();
      LCR_TICK(5);  // wait after the job pool
      if (go) make_constraints(w, m, verts, true);
    } else {
      if (flags & LCR_LS_BAR_CON) __syncthreads();
      if (go) make_constraints(w, m, verts, false);
    }
    if (go) smooth_forces(w, m);
    LCR_TICK(6);  // constraint rows, smooth forces
    if (flags & LCR_LS_BAR_SOL) __syncthreads();
    LCR_TICK(7);  // wait before the solver
    if (flags & LCR_LS_SYNC_NEWTON) solve_constraints<T, NC, true>(w, m, tol, go);
    else if (go) solve_constraints<T, NC, false>(w, m, tol);
    LCR_TICK(8);  // Newton solve
    if (go) {
      if (check_acc(w, m)) forward(w, m, verts);
      integrate(w, m);
    }
    LCR_TICK(9);  // integrate
  }
  if (valid) {
    if (moved(w)) redo_push(redo, env);  // outgrew the fast workspace: redone over the big one, nothing written here
    else {
      if (go) env_step_end(w, m, io, env);
      store_state(w, s, env);
    }
  }
  if (PROF && valid && LANE == 0) {
    LCR_TICK(0);
    for (int k = 0; k < 10; k++) prof[(size_t)env * 10 + k] = tp[k];
  }
#undef LCR_TICK
}

// ---------------------------------------------------------------- phased execution
// The same device functions as k_step, but one small kernel per phase of mj_step with the per-env workspace
// parked in HBM/L2 between launches.  Every launch has a small, homogeneous instruction footprint (the fused
// kernel is instruction-fetch bound: ~200 KB of code per substep against a 32 KB L1.5 I-cache) and variable-cost
// phases (collision, Newton solve) no longer hold the uniform ones back.  Staging is coalesced 128-bit copies.
// (Superseded by the flow kernel of lcr_flow.cuh, which runs the same phases from device-side queues inside one
// persistent launch; kept as the launch-per-phase reference the flow kernel is tested against bit by bit.)
template <typename T, int NC>
DI void store_ws(const Ws<T, NC>& w, Ws<T, NC>* g, int env, bool with_J) {
  typedef Ws<T, NC> WsT;
  constexpr int HEAD = (int)(offsetof(WsT, J) / 16), JROW = WsT::JS * (int)sizeof(T);
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(g + env);
  const uint4* src = reinterpret_cast<const uint4*>(&w);
  for (int i = LANE; i < HEAD; i += 32) dst[i] = src[i];
  if (with_J) {
    const int n4 = (w.nefc * JROW + 15) / 16;
    for (int i = LANE; i < n4; i += 32) dst[HEAD + i] = src[HEAD + i];
  }
}

// Staging by region.  The members of Ws are declared in the order state | kinematics | dynamics vectors | Hessian |
// contacts | rows | counts + cache | candidates | J with 16-byte aligned region starts; a phase kernel only moves the
// regions it reads / the ones it changed (the parked workspaces do not fit L2 for large batches, so every byte is HBM
// traffic and every dependent 512 bytes a round trip).
#define LCR_OFF(member) ((int)offsetof(WsT, member))
template <typename T, int NC>
DI void load_ws_range(Ws<T, NC>& w, const Ws<T, NC>* g, int env, int off0, int off1) {
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(g + env) + off0);
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(&w) + off0);
  const int n4 = (off1 - off0) / 16;
  constexpr int B = 8;
  for (int base = LANE; base < n4; base += 32 * B) {
    uint4 t[B];
#pragma unroll
    for (int k = 0; k < B; k++) { const int i = base + 32 * k; if (i < n4) t[k] = src[i]; }
#pragma unroll
    for (int k = 0; k < B; k++) { const int i = base + 32 * k; if (i < n4) dst[i] = t[k]; }
  }
}
template <typename T, int NC>
DI void store_ws_range(const Ws<T, NC>& w, Ws<T, NC>* g, int env, int off0, int off1) {
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(g + env) + off0);
  const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(&w) + off0);
  const int n4 = (off1 - off0) / 16;
  for (int i = LANE; i < n4; i += 32) dst[i] = src[i];
}
template <typename T, int NC>
DI void load_ws_J(Ws<T, NC>& w, const Ws<T, NC>* g, int env) {  // the live rows of J (needs w.nefc)
  typedef Ws<T, NC> WsT;
  load_ws_range(w, g, env, LCR_OFF(J), LCR_OFF(J) + ((w.nefc * WsT::JS * (int)sizeof(T) + 15) / 16) * 16);
}
template <typename T, int NC>
DI void store_ws_J(const Ws<T, NC>& w, Ws<T, NC>* g, int env) {
  typedef Ws<T, NC> WsT;
  store_ws_range(w, g, env, LCR_OFF(J), LCR_OFF(J) + ((w.nefc * WsT::JS * (int)sizeof(T) + 15) / 16) * 16);
}

// seat -> env of the phased chain: identity, or the work-aware order written by k_sched (heaviest envs of the previous
// step first, dealt out to the env groups like cards; -1 = padding seat)
DI int ph_env(const int* __restrict__ perm, int seat) { return perm ? perm[seat] : seat; }

template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_begin(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                                     Ws<T, NC>* __restrict__ gws, StepIO io, Redo redo, int env0, const int* __restrict__ perm) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = ph_env(perm, env0 + blockIdx.x);
  if (env < 0) return;
  load_state(w, s, env);
  const DevModel<T>& m = *dm;
  const bool go = env_step_begin(w, m, verts, io, env);
  if (LANE == 0) { w.skip = go ? 0 : 1; w.redo_forward = 0; w.nefc = 0; w.ncon = 0; w.nlim = 0; }
  if (moved(w)) {  // the IK / reset forward passes outgrew the fast workspace: the whole step is redone over the big one
    if (LANE == 0) w.skip = 1;
    redo_push(redo, env);
  } else if (!go) store_state(w, s, env);
  store_ws(w, gws, env, false);
}

// [integrate the previous substep] -> checks -> kinematics -> inertia / bias -> smooth forces
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_dyn(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, Ws<T, NC>* __restrict__ gws, int first, int env0, const int* __restrict__ perm,
                                                   int* __restrict__ jobq) {
  typedef Ws<T, NC> WsT;
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = ph_env(perm, env0 + blockIdx.x);
  if (env < 0 || gws[env].skip) return;
  // reads: state, dynamics vectors (M, qacc for the integration of the previous substep), counts + cache, candidate block
  load_ws_range(w, gws, env, 0, LCR_OFF(xpos));
  load_ws_range(w, gws, env, LCR_OFF(M), LCR_OFF(H));
  load_ws_range(w, gws, env, LCR_OFF(ncon), LCR_OFF(J));
  __syncwarp();
  const DevModel<T>& m = *dm;
  bool redone = false;
  if (!first) {
    if (w.redo_forward) { const int ov = w.ovf; forward(w, m, verts); if (LANE == 0) { w.redo_forward = 0; w.ovf = ov; } __syncwarp(); redone = true; }
    integrate(w, m);
  }
  check_state(w, m);
  kinematics(w, m);
  inertia_and_bias(w, m);
  smooth_forces(w, m);
  collect_candidates(w, m);
  __syncwarp();
  if (jobq) {  // one queue entry per convex candidate of this env: env << 6 | candidate (LCR_JOBQ_* layout, see k_ph_jobq)
    const int nj = w.ncand < WsT::MAXCAND ? w.ncand : WsT::MAXCAND;
    if (nj > 0) {
      int base = 0;
      if (LANE == 0) base = atomicAdd(&jobq[LCR_JOBQ_COUNT], nj);
      base = __shfl_sync(FULLMASK, base, 0);
      for (int k = LANE; k < nj; k += 32) jobq[LCR_JOBQ_ITEMS + base + k] = (env << 6) | k;
    }
  }
  // writes: state, kinematics, dynamics vectors, candidate block (+ the cache if mj_forward was re-run)
  store_ws_range(w, gws, env, 0, LCR_OFF(H));
  store_ws_range(w, gws, env, redone ? LCR_OFF(ncon) : LCR_OFF(cand_key), LCR_OFF(J));
}
// narrowphase jobs: one warp per (env, slot); the workspace stays in HBM/L2 and is only read, results go to the
// candidate result rows.  No shared memory, so the hull vertices stay L1 resident.
#ifndef LCR_NSLOT
#define LCR_NSLOT 8
#endif
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_job(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, Ws<T, NC>* __restrict__ gws, int env0, const int* __restrict__ perm) {
  const int env = ph_env(perm, env0 + blockIdx.x / LCR_NSLOT), slot = blockIdx.x % LCR_NSLOT;
  if (env < 0) return;
  Ws<T, NC>& w = gws[env];
  if (w.skip) return;
  const int n = w.ncand < Ws<T, NC>::MAXCAND ? w.ncand : Ws<T, NC>::MAXCAND;
  T (*res)[8] = cand_res(w);
  for (int k = slot; k < n; k += LCR_NSLOT) {
    T r[8];
    narrowphase_job<T, NC, false>(w, *dm, verts, w.cand_key[k], r);
    if (LANE < 8) res[k][LANE] = r[LANE];
  }
}

// The same jobs from ONE queue per chain (filled by k_ph_dyn, heaviest envs first): a fixed grid of warps pulls entries with an
// atomic ticket until the queue is empty, so no warp is launched for an env without candidates and the jobs of a folded arm are
// spread over as many warps as it has penetrating pairs.  jobq: [LCR_JOBQ_COUNT] entries, [LCR_JOBQ_NEXT] next ticket, items from
// LCR_JOBQ_ITEMS; both counters are cleared by k_ph_col (which runs after this kernel and before the next k_ph_dyn).
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_jobq(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, Ws<T, NC>* __restrict__ gws, int* __restrict__ jobq) {
  const int total = *reinterpret_cast<volatile int*>(&jobq[LCR_JOBQ_COUNT]);
  for (;;) {
    int j = 0;
    if (LANE == 0) j = atomicAdd(&jobq[LCR_JOBQ_NEXT], 1);
    j = __shfl_sync(FULLMASK, j, 0);
    if (j >= total) return;
    const int item = jobq[LCR_JOBQ_ITEMS + j], env = item >> 6, k = item & 63;
    Ws<T, NC>& w = gws[env];
    T r[8];
    narrowphase_job<T, NC, false>(w, *dm, verts, w.cand_key[k], r);
    if (LANE < 8) cand_res(w)[k][LANE] = r[LANE];
  }
}

template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_col(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, Ws<T, NC>* __restrict__ gws, int env0, const int* __restrict__ perm,
                                                   int substep, int* __restrict__ mig, int* __restrict__ jobq) {
  typedef Ws<T, NC> WsT;
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  if (jobq && blockIdx.x == 0 && LANE == 0) { jobq[LCR_JOBQ_COUNT] = 0; jobq[LCR_JOBQ_NEXT] = 0; }  // (the job kernel of this substep is done)
  const int env = ph_env(perm, env0 + blockIdx.x);
  if (env < 0 || gws[env].skip) return;
  // reads: state, kinematics, the job results (they alias e_w / e_g / e_p), counts + cache, candidate block
  load_ws_range(w, gws, env, 0, LCR_OFF(M));
  load_ws_range(w, gws, env, LCR_OFF(e_w), LCR_OFF(e_unit));
  load_ws_range(w, gws, env, LCR_OFF(ncon), LCR_OFF(J));
  __syncwarp();
  make_constraints(w, *dm, verts, true);
  __syncwarp();
  if (w.ovf && mig) {
    // the rows of this substep do not fit the fast workspace: the env migrates to the big one NOW (mig[0] = count, then env | substep << 24;
    // one list per chain and substep, consumed by the BIG pass launched behind this kernel) and resumes there from the parked start
    // state of this substep, beside the rest of the chain -- instead of being redone from the step start after the chain
    int pos = 0;
    if (LANE == 0) pos = atomicAdd(&mig[0], 1);
    pos = __shfl_sync(FULLMASK, pos, 0);
    if (pos < LCR_MIGCAP) {
      if (LANE == 0) { mig[1 + pos] = env | (substep << 24); gws[env].skip = 1; }
      return;
    }
  }
  // writes: diag (state block), contacts, row parameters, row -> contact maps, counts + cache, J
  store_ws_range(w, gws, env, 0, LCR_OFF(xpos));
  store_ws_range(w, gws, env, LCR_OFF(c_pos), LCR_OFF(e_jar));
  store_ws_range(w, gws, env, LCR_OFF(e_unit), LCR_OFF(cand_key));
  store_ws_J(w, gws, env);
  if (w.ovf && LANE == 0) gws[env].ovf = 1;  // (the flags block is not written back by this phase)
}
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_sol(const DevModel<T>* __restrict__ dm, Ws<T, NC>* __restrict__ gws, int env0, const int* __restrict__ perm) {
  typedef Ws<T, NC> WsT;
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = ph_env(perm, env0 + blockIdx.x);
  if (env < 0 || gws[env].skip) return;
  // reads: state (warm start), dynamics vectors, contact scalars (not positions / frames), row parameters, counts, flags, J
  load_ws_range(w, gws, env, 0, LCR_OFF(xpos));
  load_ws_range(w, gws, env, LCR_OFF(M), LCR_OFF(H));
  load_ws_range(w, gws, env, LCR_OFF(c_dist), LCR_OFF(e_jar));
  load_ws_range(w, gws, env, LCR_OFF(ncon), LCR_OFF(sa_dir));
  load_ws_range(w, gws, env, LCR_OFF(cand_key), LCR_OFF(J));
  __syncwarp();
  load_ws_J(w, gws, env);
  __syncwarp();
  const DevModel<T>& m = *dm;
  solve_constraints<T, NC, false>(w, m, solver_tol<T>(m));
  if (check_acc(w, m)) { if (LANE == 0) w.redo_forward = 1; }
  __syncwarp();
  // writes: state (warm start, diag; qpos / qvel if the env was reset), dynamics vectors (qacc), flags
  store_ws_range(w, gws, env, 0, LCR_OFF(xpos));
  store_ws_range(w, gws, env, LCR_OFF(M), LCR_OFF(H));
  store_ws_range(w, gws, env, LCR_OFF(cand_key), LCR_OFF(J));
}
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_ph_end(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                                   Ws<T, NC>* __restrict__ gws, StepIO io, Redo redo, int env0, const int* __restrict__ perm) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = ph_env(perm, env0 + blockIdx.x);
  if (env < 0 || gws[env].skip) return;
  typedef Ws<T, NC> WsT;
  load_ws_range(w, gws, env, 0, LCR_OFF(xpos));          // state
  load_ws_range(w, gws, env, LCR_OFF(M), LCR_OFF(H));    // M, qacc
  load_ws_range(w, gws, env, LCR_OFF(ncon), LCR_OFF(J)); // counts + cache (stored with the state), flags
  __syncwarp();
  if (moved(w)) { redo_push(redo, env); return; }  // a substep outgrew the fast workspace: redone over the big one
  const DevModel<T>& m = *dm;
  if (m.n_substeps > 0) {  // the integration of the last substep (with n_substeps == 0 there is none: mj_step was never called)
    if (w.redo_forward) forward(w, m, verts);
    integrate(w, m);
  }
  env_step_end(w, m, io, env);
  store_state(w, s, env);
}

// debug / test hook: mj_forward on the current state (not written back) and dump of the contact list; always over the
// BIG workspace, so that the list is never cut by the fast caps
template <typename T, int NC>
__global__ void __launch_bounds__(32, 2) k_debug_contacts(const DevModel<T>* __restrict__ dm, const T* __restrict__ verts, DevState<T> s,
                                                          double* __restrict__ out, int32_t* __restrict__ ncon_out) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = blockIdx.x;
  load_state(w, s, env);
  forward(w, *dm, verts);
  const int ncon = w.ncon;
  if (LANE == 0) ncon_out[env] = ncon;
  for (int ci = LANE; ci < ncon; ci += 32) {
    double* o = out + ((size_t)env * LCR_MAXCON_BIG + ci) * 12;
    for (int k = 0; k < 3; k++) { o[k] = (double)w.c_pos[ci][k]; o[3 + k] = (double)w.c_frame[ci][k]; }
    o[6] = (double)w.c_dist[ci]; o[7] = w.c_b1[ci]; o[8] = w.c_b2[ci]; o[9] = dm->par(w.c_par[ci])->dim; o[10] = (double)w.c_mu[ci]; o[11] = w.c_efc[ci];
  }
}

// empty separating-axis cache of env e in HBM: keys -1, next 0 (the block layout does not depend on NC)
template <typename T> DI void sa_empty(DevState<T> s, int e) {
  unsigned char* blk = s.sa + (size_t)e * Ws<T, 1>::SA_BYTES;
  short* key = reinterpret_cast<short*>(blk + Ws<T, 1>::SA_WORDS * sizeof(T));
  for (int k = 0; k < LCR_NSA; k++) key[k] = -1;
  int* next = reinterpret_cast<int*>(blk + Ws<T, 1>::SA_WORDS * sizeof(T) + LCR_NSA * 2);
  for (int k = 0; k < 4; k++) next[k] = 0;
}
// row-major float64 <-> per-env records of T; rng = the PCG64 state of the env's reset stream (4 x u64)
// world poses of the kinematic tree for the renderer (image observations, reach_cube_env.py:288-292): mj_kinematics on the current
// state (nothing written back); out [n][NB][12] float32 = xpos[3] | xmat[9] of the 7 arm bodies, then the boxes (cubes, walls)
template <typename T, int NC>
__global__ void __launch_bounds__(32, 16) k_poses(const DevModel<T>* __restrict__ dm, DevState<T> s, float* __restrict__ out) {
  Ws<T, NC>& w = *reinterpret_cast<Ws<T, NC>*>(lcr_smem);
  const int env = blockIdx.x;
  load_state(w, s, env);
  kinematics(w, *dm);
  __syncwarp();
  constexpr int NB = Ws<T, NC>::NB;
  float* o = out + (size_t)env * NB * 12;
  for (int i = LANE; i < NB * 12; i += 32) {
    const int b = i / 12, k = i % 12;
    o[i] = (float)(k < 3 ? w.xpos[b][k] : w.xmat[b][k - 3]);
  }
}

template <typename T>
__global__ void k_get_state(DevState<T> s, int nq, int nv, double* qpos, double* qvel, double* ctrl, double* warm, double* aux, int32_t* ints,
                            unsigned long long* rng) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  const T* r = s.st + (size_t)e * s.nfp;
  int f = 0;
  for (int k = 0; k < nq; k++, f++) if (qpos) qpos[(size_t)e * nq + k] = (double)r[f];
  for (int k = 0; k < nv; k++, f++) if (qvel) qvel[(size_t)e * nv + k] = (double)r[f];
  for (int k = 0; k < 6; k++, f++) if (ctrl) ctrl[(size_t)e * 6 + k] = (double)r[f];
  for (int k = 0; k < nv; k++, f++) if (warm) warm[(size_t)e * nv + k] = (double)r[f];
  for (int k = 0; k < LCR_NAUX; k++, f++) if (aux) aux[(size_t)e * LCR_NAUX + k] = (double)r[f];
  if (ints) for (int k = 0; k < LCR_NINT; k++) ints[(size_t)e * LCR_NINT + k] = s.ib[(size_t)e * LCR_IB_WORDS + k];
  if (rng) {
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(s.ib + (size_t)e * LCR_IB_WORDS + 8);
    for (int k = 0; k < 4; k++) rng[4 * (size_t)e + k] = src[k];
  }
}
template <typename T>
__global__ void k_set_state(DevState<T> s, int nq, int nv, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                            const double* aux, const int32_t* ints, const unsigned long long* rng) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  T* r = s.st + (size_t)e * s.nfp;
  int f = 0;
  for (int k = 0; k < nq; k++, f++) if (qpos) r[f] = (T)qpos[(size_t)e * nq + k];
  for (int k = 0; k < nv; k++, f++) if (qvel) r[f] = (T)qvel[(size_t)e * nv + k];
  for (int k = 0; k < 6; k++, f++) if (ctrl) r[f] = (T)ctrl[(size_t)e * 6 + k];
  for (int k = 0; k < nv; k++, f++) if (warm) r[f] = (T)warm[(size_t)e * nv + k];
  for (int k = 0; k < LCR_NAUX; k++, f++) if (aux) r[f] = (T)aux[(size_t)e * LCR_NAUX + k];
  if (ints) for (int k = 0; k < LCR_NINT; k++) s.ib[(size_t)e * LCR_IB_WORDS + k] = ints[(size_t)e * LCR_NINT + k];
  if (rng) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(s.ib + (size_t)e * LCR_IB_WORDS + 8);
    for (int k = 0; k < 4; k++) dst[k] = rng[4 * (size_t)e + k];
  }
  sa_empty<T>(s, e);  // the separating-axis cache is not part of the checkpointed state (it never changes a result)
}
template <typename T>
__global__ void k_init_state(const DevModel<T>* dm, DevState<T> s) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  T* r = s.st + (size_t)e * s.nfp;
  for (int f = 0; f < s.nfp; f++) r[f] = 0;
  for (int c = 0; c < dm->ncube; c++) {
    for (int k = 0; k < 3; k++) r[6 + 7 * c + k] = dm->cube_qpos0[c][k];
    r[6 + 7 * c + 3] = 1;
  }
  int32_t* ib = s.ib + (size_t)e * LCR_IB_WORDS;
  for (int k = 0; k < LCR_IB_WORDS; k++) ib[k] = 0;
  ib[8 + 2] = 1;  // rng[1] (state_lo) = 1
  ib[8 + 6] = 1;  // rng[3] (inc_lo) = 1
  sa_empty<T>(s, e);
}
template <typename T>
__global__ void k_get_diag(DevState<T> s, int32_t* out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n) return;
  for (int k = 0; k < LCR_NDIAG; k++) out[(size_t)e * LCR_NDIAG + k] = s.ib[(size_t)e * LCR_IB_WORDS + LCR_NINT + k];
}
// per-env PCG64 states; with a mask only the selected envs are reseeded (the streams of the others go on)
template <typename T>
__global__ void k_seed(DevState<T> s, const unsigned long long* st, const uint8_t* __restrict__ mask) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= s.n || (mask && !mask[e])) return;
  unsigned long long* rng = reinterpret_cast<unsigned long long*>(s.ib + (size_t)e * LCR_IB_WORDS + 8);
  for (int k = 0; k < 4; k++) rng[k] = st[4 * (size_t)e + k];
}

// obs | reward | terminated | truncated | success as one float32 record per env, from separate arrays (the step kernels
// write the record themselves when lcr_step_rec is given one; this is the stand-alone version of the C-ABI)
static __global__ void k_pack(const float* __restrict__ obs, const float* __restrict__ reward, const uint8_t* __restrict__ term,
                       const uint8_t* __restrict__ trunc, const uint8_t* __restrict__ succ, float* __restrict__ rec, int n, int od) {
  const int w = od + 4, total = n * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int e = i / w, c = i - e * w;
    float v;
    if (c < od) v = obs[(size_t)e * od + c];
    else if (c == od) v = reward[e];
    else if (c == od + 1) v = (float)term[e];
    else if (c == od + 2) v = (float)trunc[e];
    else v = (float)succ[e];
    rec[i] = v;
  }
}


// Batched trajectory recorder (the device side of vec.TrajectoryRecorder; replaces HDF5_Recorder.capture_frame and the episode
// cut of RecordHDF5Wrapper.step, envs/wrappers/record_hdf5.py:40-45,116-137).  One warp per env: appends the row
// arm_qpos[6] | arm_qvel[6] | action[A] of this step to the env's open trajectory traj[env][t]; when the episode ended in this
// step the trajectory moves to the next free slot of the finished-episode pool (slot = atomic counter, meta = env, length)
// and the env starts a new one.  No host involvement: the host drains the pool whenever it likes (count[0] = episodes in the
// pool, count[1] = episodes dropped because the pool was full).
static __global__ void k_rec_append(const float* __restrict__ obs, int od, const float* __restrict__ act, int A, const uint8_t* __restrict__ term,
                             const uint8_t* __restrict__ trunc, int n, int h, float* __restrict__ traj, int32_t* __restrict__ len,
                             float* __restrict__ pool, int32_t* __restrict__ meta, int32_t* __restrict__ count, int cap) {
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31, W = 12 + A;
  if (env >= n) return;
  int t = len[env];
  float* tr = traj + (size_t)env * h * W;
  const int row = t < h ? t : h - 1;  // (an episode longer than the horizon keeps overwriting its last row)
  if (lane < W) tr[(size_t)row * W + lane] = lane < 12 ? obs[(size_t)env * od + lane] : act[(size_t)env * A + lane - 12];
  t++;
  __syncwarp();
  if (!(term[env] | trunc[env])) {
    if (lane == 0) len[env] = t;
    return;
  }
  int slot = 0;
  if (lane == 0) slot = atomicAdd(&count[0], 1);
  slot = __shfl_sync(FULLMASK, slot, 0);
  const int L = t < h ? t : h;
  if (slot < cap) {
    float* dst = pool + (size_t)slot * h * W;
    for (int i = lane; i < L * W; i += 32) dst[i] = tr[i];
    if (lane == 0) { meta[2 * slot] = env; meta[2 * slot + 1] = L; }
  } else if (lane == 0) {
    atomicAdd(&count[1], 1);
  }
  if (lane == 0) len[env] = 0;
}

}  // namespace lcr
#include "lcr_flow.cuh"
namespace lcr {

// ---------------------------------------------------------------- launchers
#define LCR_SMEM_MAX (227 * 1024)
template <typename T, int S>
void LaunchNC<T, S>::prepare() {
  constexpr int B = S | LCR_NC_BIG;
  const int fast = (int)sizeof(Ws<T, S>), big = (int)sizeof(Ws<T, B>);
  cudaFuncSetAttribute(k_step<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_step<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_reset<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_substeps<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_substeps<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_ik<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_poses<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_debug_contacts<T, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_ph_begin<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_ph_dyn<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_ph_col<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_ph_sol<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_ph_end<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, fast);
  cudaFuncSetAttribute(k_step_ls<T, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, LCR_LS_MAXSMEM);
  cudaFuncSetAttribute(k_step_ls<T, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, LCR_LS_MAXSMEM);
  cudaFuncSetAttribute(k_step_ls<T, B, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  cudaFuncSetAttribute(k_flow<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, flow_smem());
  // all of the SM's L1/shared array as shared memory: several CTAs of a few workspaces each must fit one SM
  cudaFuncSetAttribute(k_step_ls<T, S, false>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_step_ls<T, S, true>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_flow<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k_step<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  // (experiment knob, off by default: the SAME maximal shared-memory carve-out for every kernel of the phased chain, including the
  // narrowphase job kernel that uses no shared memory, so that CTAs of different kernels of the chain can share an SM.  Measured on
  // B200, PushCube 16 384: 20.5 -> 23.6 ms per step -- the job kernel loses the L1 that holds the hull vertices)
  if (getenv("LCR_PH_CARVEOUT") && atoi(getenv("LCR_PH_CARVEOUT")) != 0) {
    const int co = (int)cudaSharedmemCarveoutMaxShared;
    cudaFuncSetAttribute(k_ph_begin<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_ph_dyn<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_ph_job<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_ph_col<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_ph_sol<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_ph_end<T, S>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
    cudaFuncSetAttribute(k_step_ls<T, B, false>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
  }
}
template <typename T, int S> size_t LaunchNC<T, S>::ws_bytes() { return sizeof(Ws<T, S>); }

template <typename T, int S>
void LaunchNC<T, S>::reset(const DevModel<T>* dm, const T* verts, DevState<T> s, const uint8_t* mask, float* obs, cudaStream_t st) {
  k_reset<T, S><<<s.n, 32, sizeof(Ws<T, S>), st>>>(dm, verts, s, mask, obs);
}
template <typename T, int S>
void LaunchNC<T, S>::step(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo redo, cudaStream_t st) {
  k_step<T, S><<<s.n, 32, sizeof(Ws<T, S>), st>>>(dm, verts, s, io, redo, nullptr, nullptr, 0);
}
// The envs of the list, from their unchanged start state, over the big workspace.  These are the most expensive envs of
// the batch (100 - 180 constraint rows, dozens of penetrating hull pairs) and the step waits for them, so each gets a CTA
// of its own with LCR_BIG_WARPS warps: the lockstep kernel with ONE env per CTA, whose other warps drain the env's
// narrowphase job pool.  The grid is fixed (the list length lives on the device; CTAs without a seat return at once); a
// list longer than the grid -- never seen -- is finished by the one-warp kernel from entry LCR_BIG_GRID on.
#define LCR_BIG_WARPS 8
#define LCR_BIG_GRID 2048
// BIG pass over the envs that migrated out of the phased chain in one (group, substep): mig[0] = count, then env | substep << 24
template <typename T, int S>
void LaunchNC<T, S>::step_big_resume(const DevModel<T>* dm, const T* verts, DevState<T> s, const void* gws, StepIO io, const int* mig, cudaStream_t st) {
  constexpr int B = S | LCR_NC_BIG;
  k_step_ls<T, B, false><<<LCR_MIGCAP, 32 * LCR_BIG_WARPS, sizeof(Ws<T, B>), st>>>(dm, verts, s, io, Redo{nullptr, nullptr},
                                                                                 LCR_LS_BAR_TOP | LCR_LS_BAR_CON | LCR_LS_JOB_POOL | LCR_LS_RESUME,
                                                                                 mig + 1, 1, nullptr, mig, gws);
}
template <typename T, int S>
void LaunchNC<T, S>::step_big(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo list, cudaStream_t st) {
  constexpr int B = S | LCR_NC_BIG;
  const int grid = std::min(s.n, LCR_BIG_GRID);
  k_step_ls<T, B, false><<<grid, 32 * LCR_BIG_WARPS, sizeof(Ws<T, B>), st>>>(dm, verts, s, io, Redo{nullptr, nullptr}, LCR_LS_BAR_TOP | LCR_LS_BAR_CON | LCR_LS_JOB_POOL,
                                                                          list.list, 1, nullptr, list.count);
  if (s.n > LCR_BIG_GRID)
    k_step<T, B><<<296, 32, sizeof(Ws<T, B>), st>>>(dm, verts, s, io, Redo{nullptr, nullptr}, list.list, list.count, LCR_BIG_GRID);
}
// lockstep kernel: CTAs of `warps` envs; warps <= 0 picks the largest CTA that fits one SM
template <typename T, int S>
int LaunchNC<T, S>::lockstep_warps(int warps) {
  const int fit = (int)(LCR_LS_MAXSMEM / sizeof(Ws<T, S>));
  if (warps <= 0) warps = fit;
  return std::max(1, std::min(std::min(warps, fit), 16));
}
// `grid` CTAs of `warps` warps, the first `epc` of which own an env (seats from perm, or env = seat if perm is null)
template <typename T, int S>
void LaunchNC<T, S>::step_lockstep(const DevModel<T>* dm, const T* verts, DevState<T> s, StepIO io, Redo redo, int grid, int warps, int epc, int flags,
                                   const int* perm, long long* prof, cudaStream_t st) {
  if (prof) k_step_ls<T, S, true><<<grid, 32 * warps, sizeof(Ws<T, S>) * epc, st>>>(dm, verts, s, io, redo, flags, perm, epc, prof, nullptr);
  else k_step_ls<T, S, false><<<grid, 32 * warps, sizeof(Ws<T, S>) * epc, st>>>(dm, verts, s, io, redo, flags, perm, epc, prof, nullptr);
}
// one chain of 2 + 4 * n_substeps launches over the env range [env0, env0 + cnt) on stream st.  mig (optional): [n_substeps][1 + LCR_MIGCAP]
// migration lists of this chain (zeroed by the caller); behind the constraint-row kernel of substep k the BIG pass over list k is
// launched on side[k] (forked from st by ev_fork[k]; the caller joins side[k] through ev_join[k])
template <typename T, int S>
int LaunchNC<T, S>::step_phased(int n_substeps, const DevModel<T>* dm, const T* verts, DevState<T> s, void* gws_, StepIO io, Redo redo, int env0, int cnt,
                                const int* perm, cudaStream_t st, int* mig, cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join, int* jobq,
                                int* early) {
  typedef Ws<T, S> W;
  W* gws = reinterpret_cast<W*>(gws_);
  const size_t sm = sizeof(W);
  int nl = 2 + 4 * n_substeps;
  // (tried and removed: the Newton solve of substep k fused with the dynamics stage of substep k + 1 -- one launch and one staging
  //  round trip less per substep, bit-identical -- PushCube 16 384 15.1 -> 20.3 ms per step, StackTwoCubes 8 192 12.8 -> 15.6: like the
  //  col + sol merge of round 1, a larger heterogeneous kernel loses more in the instruction cache than the launch saves)
  // early (optional, with mig): [1 + cnt] list of the envs whose step BEGIN outgrew the fast workspace (the IK's forward passes of the ee
  // mode, the forward pass of an autoreset): their BIG pass over the whole step starts right behind k_ph_begin on side[n_substeps],
  // beside the chain, instead of after it with the redo pass
  const Redo first = early ? Redo{early, early + 1} : redo;
  k_ph_begin<T, S><<<cnt, 32, sm, st>>>(dm, verts, s, gws, io, first, env0, perm);
  if (early) {
    cudaEventRecord(ev_fork[n_substeps], st);
    cudaStreamWaitEvent(side[n_substeps], ev_fork[n_substeps], 0);
    step_big(dm, verts, s, io, first, side[n_substeps]);
    cudaEventRecord(ev_join[n_substeps], side[n_substeps]);
    nl++;
  }
  for (int k = 0; k < n_substeps; k++) {
    int* mk = mig ? mig + (size_t)k * (1 + LCR_MIGCAP) : nullptr;
    k_ph_dyn<T, S><<<cnt, 32, sm, st>>>(dm, verts, gws, k == 0, env0, perm, jobq);
    if (jobq) k_ph_jobq<T, S><<<std::min(cnt * 2, 148 * 16), 32, 0, st>>>(dm, verts, gws, jobq);
    else k_ph_job<T, S><<<cnt * LCR_NSLOT, 32, 0, st>>>(dm, verts, gws, env0, perm);
    k_ph_col<T, S><<<cnt, 32, sm, st>>>(dm, verts, gws, env0, perm, k, mk, jobq);
    if (mk) {
      cudaEventRecord(ev_fork[k], st);
      cudaStreamWaitEvent(side[k], ev_fork[k], 0);
      step_big_resume(dm, verts, s, gws_, io, mk, side[k]);
      cudaEventRecord(ev_join[k], side[k]);
      nl++;
    }
    k_ph_sol<T, S><<<cnt, 32, sm, st>>>(dm, gws, env0, perm);
  }
  k_ph_end<T, S><<<cnt, 32, sm, st>>>(dm, verts, s, gws, io, redo, env0, perm);
  return nl;
}
template <typename T, int S> int LaunchNC<T, S>::jobq_words(int cnt) { return LCR_JOBQ_ITEMS + cnt * Ws<T, S>::MAXCAND; }
// flow kernel: `warps` fast workspace slots per CTA (<= 16); BIG CTAs hold as many big workspaces as fit
template <typename T, int S> int LaunchNC<T, S>::flow_warps() { return std::max(1, std::min(16, (int)((LCR_SMEM_MAX - 256) / sizeof(Ws<T, S>)))); }
template <typename T, int S> int LaunchNC<T, S>::flow_bigslots() { return std::max(1, std::min(16, (int)((LCR_SMEM_MAX - 256) / sizeof(Ws<T, S | LCR_NC_BIG>)))); }
template <typename T, int S> int LaunchNC<T, S>::flow_smem() {
  return (int)std::max(sizeof(Ws<T, S>) * flow_warps(), sizeof(Ws<T, S | LCR_NC_BIG>) * flow_bigslots());
}
template <typename T, int S>
void LaunchNC<T, S>::step_flow(const DevModel<T>* dm, const T* verts, DevState<T> s, void* gws, StepIO io, const void* fq_, int grid, int nbigcta, int flags,
                               int t_hi, int t_big, unsigned long long* stats, cudaStream_t st) {
  const FlowQ& fq = *reinterpret_cast<const FlowQ*>(fq_);
  k_sched_flow<T><<<1, 1024, 0, st>>>(s, fq, t_hi, t_big);
  static const int dbg_warps = getenv("LCR_FLOW_WARPS") ? atoi(getenv("LCR_FLOW_WARPS")) : 0;  // debug: fewer warps per CTA
  const int warps = dbg_warps > 0 ? std::min(dbg_warps, flow_warps()) : flow_warps();
  k_flow<T, S><<<grid, 32 * warps, flow_smem(), st>>>(dm, verts, s, reinterpret_cast<Ws<T, S>*>(gws), io, fq, nbigcta, std::min(flow_bigslots(), warps), flags, stats);
}
template <typename T, int S>
void LaunchNC<T, S>::substeps(const DevModel<T>* dm, const T* verts, DevState<T> s, int n, Redo redo, cudaStream_t st) {
  k_substeps<T, S><<<s.n, 32, sizeof(Ws<T, S>), st>>>(dm, verts, s, n, redo, nullptr, nullptr);
}
template <typename T, int S>
void LaunchNC<T, S>::substeps_big(const DevModel<T>* dm, const T* verts, DevState<T> s, int n, Redo list, cudaStream_t st) {
  constexpr int B = S | LCR_NC_BIG;
  k_substeps<T, B><<<std::min(s.n, 296), 32, sizeof(Ws<T, B>), st>>>(dm, verts, s, n, Redo{nullptr, nullptr}, list.list, list.count);
}
template <typename T, int S>
void LaunchNC<T, S>::ik(const DevModel<T>* dm, const T* verts, DevState<T> s, const float* target, float* q_out, cudaStream_t st) {
  k_ik<T, S><<<s.n, 32, sizeof(Ws<T, S>), st>>>(dm, verts, s, target, q_out);
}
template <typename T, int S> int LaunchNC<T, S>::pose_slots() { return Ws<T, S>::NB; }
template <typename T, int S>
void LaunchNC<T, S>::poses(const DevModel<T>* dm, DevState<T> s, float* out, cudaStream_t st) {
  k_poses<T, S><<<s.n, 32, sizeof(Ws<T, S>), st>>>(dm, s, out);
}
template <typename T, int S>
void LaunchNC<T, S>::debug_contacts(const DevModel<T>* dm, const T* verts, DevState<T> s, double* out, int32_t* ncon, cudaStream_t st) {
  constexpr int B = S | LCR_NC_BIG;
  k_debug_contacts<T, B><<<s.n, 32, sizeof(Ws<T, B>), st>>>(dm, verts, s, out, ncon);
}

// ---- kernels that do not depend on the scene class
template <typename T>
void Launch<T>::pack(const float* obs, const float* reward, const uint8_t* term, const uint8_t* trunc, const uint8_t* succ, float* rec, int n, int od,
                     cudaStream_t st) {
  const int total = n * (od + 4), blocks = std::min((total + 255) / 256, 148 * 8);
  k_pack<<<blocks, 256, 0, st>>>(obs, reward, term, trunc, succ, rec, n, od);
}
template <typename T>
void Launch<T>::rec_append(const float* obs, int od, const float* act, int A, const uint8_t* term, const uint8_t* trunc, int n, int h, float* traj,
                           int32_t* len, float* pool, int32_t* meta, int32_t* count, int cap, cudaStream_t st) {
  k_rec_append<<<(n + 3) / 4, 128, 0, st>>>(obs, od, act, A, term, trunc, n, h, traj, len, pool, meta, count, cap);
}
template <typename T>
void Launch<T>::sched(DevState<T> s, int* perm, int W, int striped, int* big, int tbig, cudaStream_t st) {
  k_sched<T><<<1, 1024, 0, st>>>(s, perm, W, striped, big, tbig);
}
template <typename T>
void Launch<T>::get_state(int ncube, DevState<T> s, double* qpos, double* qvel, double* ctrl, double* warm, double* aux, int32_t* ints,
                          unsigned long long* rng, cudaStream_t st) {
  k_get_state<T><<<(s.n + 127) / 128, 128, 0, st>>>(s, 6 + 7 * ncube, 6 + 6 * ncube, qpos, qvel, ctrl, warm, aux, ints, rng);
}
template <typename T>
void Launch<T>::set_state(int ncube, DevState<T> s, const double* qpos, const double* qvel, const double* ctrl, const double* warm,
                          const double* aux, const int32_t* ints, const unsigned long long* rng, cudaStream_t st) {
  k_set_state<T><<<(s.n + 127) / 128, 128, 0, st>>>(s, 6 + 7 * ncube, 6 + 6 * ncube, qpos, qvel, ctrl, warm, aux, ints, rng);
}
template <typename T>
void Launch<T>::init_state(const DevModel<T>* dm, DevState<T> s, cudaStream_t st) { k_init_state<T><<<(s.n + 127) / 128, 128, 0, st>>>(dm, s); }
template <typename T>
void Launch<T>::get_diag(DevState<T> s, int32_t* out, cudaStream_t st) { k_get_diag<T><<<(s.n + 127) / 128, 128, 0, st>>>(s, out); }
template <typename T>
void Launch<T>::seed(DevState<T> s, const unsigned long long* d_state, const uint8_t* d_mask, cudaStream_t st) {
  k_seed<T><<<(s.n + 127) / 128, 128, 0, st>>>(s, d_state, d_mask);
}

}  // namespace lcr
