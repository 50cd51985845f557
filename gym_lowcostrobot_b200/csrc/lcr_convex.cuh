// lcr_convex.cuh -- warp-cooperative convex narrowphase: Minkowski Portal Refinement for box-mesh and
// mesh-mesh pairs, SAT + face clipping for box-box, and the sphere / oriented-box broadphase.
//
// MPR follows the published algorithm of libccd's ccdMPRPenetration, which MuJoCo 3.2.x calls from
// mjc_Convex for every box/mesh pair (inside mujoco.mj_step, reference reach_cube_env.py:277).  The
// portal logic is scalar and runs redundantly (warp-uniform) in every lane; the support function of
// a mesh is the parallel part: lanes stride over the hull vertices and a shuffle butterfly picks the
// first maximum.  All of this is cold code (only reached when a broadphase test passes), kept out
// of line so that the per-substep hot path stays small.
#pragma once
#include "lcr_device.cuh"

namespace lcr {

#ifndef LCR_SCAN_UNROLL
#define LCR_SCAN_UNROLL 8
#endif
constexpr int kScanUnroll = LCR_SCAN_UNROLL;  // hull-scan loads in flight per lane

template <typename T> DI T ccd_eps();
template <> DI float ccd_eps<float>() { return 1.1920929e-07f; }
template <> DI double ccd_eps<double>() { return 2.220446049250313e-16; }
template <typename T> DI bool is_zero(T x) { return fabs(x) < ccd_eps<T>(); }
template <typename T> DI bool ccd_eq(T a, T b) {
  const T ab = fabs(a - b);
  if (ab < ccd_eps<T>()) return true;
  const T fa = fabs(a), fb = fabs(b);
  return ab < ccd_eps<T>() * (fb > fa ? fb : fa);
}
template <typename T> DI bool vec_is_origin(const T* a) { return ccd_eq(a[0], (T)0) && ccd_eq(a[1], (T)0) && ccd_eq(a[2], (T)0); }
template <typename T> DI void normalize3(T* v) {
  const T n = sqrt(dot3(v, v));
  if (n > 0) { const T inv = 1 / n; v[0] *= inv; v[1] *= inv; v[2] *= inv; }
}

template <typename T>
struct Shape {  // warp-uniform descriptor
  int kind;     // 0 box, 1 mesh
  int body;     // index into w.xpos / w.xmat
  int adr, num; // mesh vertex range
  T half[3];    // box half sizes
  T center[3];  // interior point (geom centre), world
};
template <typename T> struct SPoint { T v[3], v1[3], v2[3]; };

// hull vertex i as one (float) or two (double) 128-bit read-only loads; verts is [nvert][4], 16-byte aligned
DI void load_vert(const float* __restrict__ verts, int i, float& x, float& y, float& z) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(verts) + i);
  x = v.x; y = v.y; z = v.z;
}
DI void load_vert(const double* __restrict__ verts, int i, double& x, double& y, double& z) {
  const double2 a = __ldg(reinterpret_cast<const double2*>(verts) + 2 * (size_t)i);
  x = a.x; y = a.y; z = __ldg(verts + 4 * (size_t)i + 2);
}

// support point of one shape along world direction d (all lanes get the same result)
template <typename T, int NC>
__device__ __noinline__ void shape_support(const Ws<T, NC>& w, const T* __restrict__ verts, const Shape<T>& sh, const T* d, T* out) {
  T dl[3], p[3];
  const T* R = w.xmat[sh.body];
  matT_vec(dl, R, d);
  if (sh.kind == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) p[k] = dl[k] >= 0 ? sh.half[k] : -sh.half[k];
  } else {
    T bv = (T)-1e30, bx = 0, by = 0, bz = 0;
    int bi = 0x7fffffff;
    // chunks of kScanUnroll vertices per lane: all loads of a chunk are issued before the first use (the vertex pool is
    // served by L2 -- with the whole L1 carved out as shared memory -- so a scan costs one L2 round trip per chunk
    // instead of one per vertex); out-of-range slots re-read the last vertex and are ignored
    for (int base = LANE; base < sh.num; base += 32 * kScanUnroll) {
      T vx[kScanUnroll], vy[kScanUnroll], vz[kScanUnroll];
#pragma unroll
      for (int u = 0; u < kScanUnroll; u++) {
        const int i = base + 32 * u;
        load_vert(verts, sh.adr + (i < sh.num ? i : sh.num - 1), vx[u], vy[u], vz[u]);
      }
#pragma unroll
      for (int u = 0; u < kScanUnroll; u++) {
        const int i = base + 32 * u;
        const T s = vx[u] * dl[0] + vy[u] * dl[1] + vz[u] * dl[2];
        if (i < sh.num && s > bv) { bv = s; bi = i; bx = vx[u]; by = vy[u]; bz = vz[u]; }
      }
    }
    warp_argmax(bv, bi);
    // vertex i lives in lane i % 32: fetch the winner from that lane's registers instead of another trip to L2
    p[0] = __shfl_sync(FULLMASK, bx, bi & 31); p[1] = __shfl_sync(FULLMASK, by, bi & 31); p[2] = __shfl_sync(FULLMASK, bz, bi & 31);
  }
  mat_vec(out, R, p);
#pragma unroll
  for (int k = 0; k < 3; k++) out[k] += w.xpos[sh.body][k];
}

// Supports of two hulls at once (same result as two shape_support calls): the vertex loads of both scans are issued
// together, so a support pair costs the L2 round trips of the longer scan instead of the sum.
template <typename T, int NC>
__device__ __noinline__ void dual_mesh_support(const Ws<T, NC>& w, const T* __restrict__ verts, const Shape<T>& A, const Shape<T>& B, const T* da,
                                               const T* db, T* outA, T* outB) {
  const T* RA = w.xmat[A.body];
  const T* RB = w.xmat[B.body];
  T la[3], lb[3];
  matT_vec(la, RA, da);
  matT_vec(lb, RB, db);
  T va = (T)-1e30, vb = (T)-1e30, ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
  int ia = 0x7fffffff, ib = 0x7fffffff;
  constexpr int U = kScanUnroll;  // U vertices of each hull per lane and round, all 2U loads in flight together
  const int nmax = A.num > B.num ? A.num : B.num;
  for (int base = LANE; base < nmax; base += 32 * U) {
    T xa[U], ya[U], za[U], xb[U], yb[U], zb[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + 32 * u;
      load_vert(verts, A.adr + (i < A.num ? i : A.num - 1), xa[u], ya[u], za[u]);
      load_vert(verts, B.adr + (i < B.num ? i : B.num - 1), xb[u], yb[u], zb[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int i = base + 32 * u;
      const T sa = xa[u] * la[0] + ya[u] * la[1] + za[u] * la[2];
      const T sb = xb[u] * lb[0] + yb[u] * lb[1] + zb[u] * lb[2];
      if (i < A.num && sa > va) { va = sa; ia = i; ax = xa[u]; ay = ya[u]; az = za[u]; }
      if (i < B.num && sb > vb) { vb = sb; ib = i; bx = xb[u]; by = yb[u]; bz = zb[u]; }
    }
  }
  warp_argmax(va, ia);
  warp_argmax(vb, ib);
  T pa[3] = {__shfl_sync(FULLMASK, ax, ia & 31), __shfl_sync(FULLMASK, ay, ia & 31), __shfl_sync(FULLMASK, az, ia & 31)};
  T pb[3] = {__shfl_sync(FULLMASK, bx, ib & 31), __shfl_sync(FULLMASK, by, ib & 31), __shfl_sync(FULLMASK, bz, ib & 31)};
  mat_vec(outA, RA, pa);
  mat_vec(outB, RB, pb);
#pragma unroll
  for (int k = 0; k < 3; k++) { outA[k] += w.xpos[A.body][k]; outB[k] += w.xpos[B.body][k]; }
}

// DUAL: evaluate two hulls with dual_mesh_support (shorter dependent chain: the lockstep / fused kernels, where a few warps
// scan at a time); the phased job kernel, where every warp of the GPU scans at once, is faster with the plain scans
template <typename T, int NC, bool DUAL>
DI void md_support(const Ws<T, NC>& w, const T* __restrict__ verts, const Shape<T>& A, const Shape<T>& B, const T* dir, SPoint<T>& p) {
  // (both directions are private copies: `dir` may point into a caller array that the optimiser also uses as an output)
  const T dx = dir[0], dy = dir[1], dz = dir[2];
  T da[3] = {dx, dy, dz}, db[3] = {-dx, -dy, -dz};
  if (DUAL && A.kind == 1 && B.kind == 1) dual_mesh_support(w, verts, A, B, da, db, p.v1, p.v2);
  else {
    shape_support(w, verts, A, da, p.v1);
    shape_support(w, verts, B, db, p.v2);
  }
#pragma unroll
  for (int k = 0; k < 3; k++) p.v[k] = p.v1[k] - p.v2[k];
}

template <typename T> DI void portal_dir(const SPoint<T>* P, T* dir) {
  T a[3], b[3];
#pragma unroll
  for (int k = 0; k < 3; k++) { a[k] = P[2].v[k] - P[1].v[k]; b[k] = P[3].v[k] - P[1].v[k]; }
  cross3(dir, a, b);
  normalize3(dir);
}
template <typename T> DI bool reach_tolerance(const SPoint<T>* P, const SPoint<T>& v4, const T* dir, T tol) {
  const T dv1 = dot3(P[1].v, dir), dv2 = dot3(P[2].v, dir), dv3 = dot3(P[3].v, dir), dv4 = dot3(v4.v, dir);
  T d = dv4 - dv1;
  if (dv4 - dv2 < d) d = dv4 - dv2;
  if (dv4 - dv3 < d) d = dv4 - dv3;
  return ccd_eq(d, tol) || d < tol;
}
template <typename T> DI void expand_portal(SPoint<T>* P, const SPoint<T>& v4) {
  T c[3];
  cross3(c, v4.v, P[0].v);
  if (dot3(P[1].v, c) > 0) {
    if (dot3(P[2].v, c) > 0) P[1] = v4; else P[3] = v4;
  } else {
    if (dot3(P[3].v, c) > 0) P[2] = v4; else P[1] = v4;
  }
}
template <typename T> DI T origin_tri_dist2(const T* a, const T* b, const T* c, T* wit) {
  T ab[3], ac[3], ap[3] = {-a[0], -a[1], -a[2]}, bp[3] = {-b[0], -b[1], -b[2]}, cp[3] = {-c[0], -c[1], -c[2]};
#pragma unroll
  for (int k = 0; k < 3; k++) { ab[k] = b[k] - a[k]; ac[k] = c[k] - a[k]; }
  const T d1 = dot3(ab, ap), d2 = dot3(ac, ap), d3 = dot3(ab, bp), d4 = dot3(ac, bp), d5 = dot3(ab, cp), d6 = dot3(ac, cp);
  const T vc = d1 * d4 - d3 * d2, vb = d5 * d2 - d1 * d6, va = d3 * d6 - d5 * d4;
  if (d1 <= 0 && d2 <= 0) { wit[0] = a[0]; wit[1] = a[1]; wit[2] = a[2]; }
  else if (d3 >= 0 && d4 <= d3) { wit[0] = b[0]; wit[1] = b[1]; wit[2] = b[2]; }
  else if (vc <= 0 && d1 >= 0 && d3 <= 0) { const T v = d1 / (d1 - d3); for (int k = 0; k < 3; k++) wit[k] = a[k] + v * ab[k]; }
  else if (d6 >= 0 && d5 <= d6) { wit[0] = c[0]; wit[1] = c[1]; wit[2] = c[2]; }
  else if (vb <= 0 && d2 >= 0 && d6 <= 0) { const T ww = d2 / (d2 - d6); for (int k = 0; k < 3; k++) wit[k] = a[k] + ww * ac[k]; }
  else if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) {
    const T ww = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    for (int k = 0; k < 3; k++) wit[k] = b[k] + ww * (c[k] - b[k]);
  } else {
    const T den = 1 / (va + vb + vc), v = vb * den, ww = vc * den;
    for (int k = 0; k < 3; k++) wit[k] = a[k] + ab[k] * v + ac[k] * ww;
  }
  return dot3(wit, wit);
}
template <typename T> DI void find_pos(const SPoint<T>* P, T* pos) {
  T dir[3], b[4], c[3];
  portal_dir(P, dir);
  cross3(c, P[1].v, P[2].v); b[0] = dot3(c, P[3].v);
  cross3(c, P[3].v, P[2].v); b[1] = dot3(c, P[0].v);
  cross3(c, P[0].v, P[1].v); b[2] = dot3(c, P[3].v);
  cross3(c, P[2].v, P[1].v); b[3] = dot3(c, P[0].v);
  T sum = b[0] + b[1] + b[2] + b[3];
  if (is_zero(sum) || sum < 0) {
    b[0] = 0;
    cross3(c, P[2].v, P[3].v); b[1] = dot3(c, dir);
    cross3(c, P[3].v, P[1].v); b[2] = dot3(c, dir);
    cross3(c, P[1].v, P[2].v); b[3] = dot3(c, dir);
    sum = b[1] + b[2] + b[3];
  }
  const T inv = 1 / sum;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T p1 = 0, p2 = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { p1 += b[i] * P[i].v1[k]; p2 += b[i] * P[i].v2[k]; }
    pos[k] = (T)0.5 * (p1 + p2) * inv;
  }
}

// returns 1 and fills depth / pdir / pos if the shapes penetrate; 0 if a separating direction was found (returned
// in pdir: max over A-B of x.pdir <= 0); -1 if separated without a usable direction (warp-uniform)
template <typename T, int NC, bool DUAL>
__device__ __noinline__ int mpr_penetration(const Ws<T, NC>& w, const T* __restrict__ verts, const Shape<T>& A, const Shape<T>& B, T& depth,
                                             T* pdir, T* pos) {
  const T tol = (T)1e-6;
  SPoint<T> P[4], v4;
  T dir[3], va[3], vb[3], d;
#pragma unroll
  for (int k = 0; k < 3; k++) { P[0].v1[k] = A.center[k]; P[0].v2[k] = B.center[k]; P[0].v[k] = A.center[k] - B.center[k]; }
  if (vec_is_origin(P[0].v)) P[0].v[0] += ccd_eps<T>() * 10;
#pragma unroll
  for (int k = 0; k < 3; k++) dir[k] = -P[0].v[k];
  normalize3(dir);
  md_support<T, NC, DUAL>(w, verts, A, B, dir, P[1]);
  d = dot3(P[1].v, dir);
  if (is_zero(d) || d < 0) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; return 0; }
  cross3(dir, P[0].v, P[1].v);
  if (is_zero(dot3(dir, dir))) {
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = (T)0.5 * (P[1].v1[k] + P[1].v2[k]);
    if (vec_is_origin(P[1].v)) { depth = 0; pdir[0] = pdir[1] = pdir[2] = 0; return 1; }
    depth = sqrt(dot3(P[1].v, P[1].v));
    pdir[0] = P[1].v[0]; pdir[1] = P[1].v[1]; pdir[2] = P[1].v[2];
    normalize3(pdir);
    return 1;
  }
  normalize3(dir);
  md_support<T, NC, DUAL>(w, verts, A, B, dir, P[2]);
  d = dot3(P[2].v, dir);
  if (is_zero(d) || d < 0) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; return 0; }
#pragma unroll
  for (int k = 0; k < 3; k++) { va[k] = P[1].v[k] - P[0].v[k]; vb[k] = P[2].v[k] - P[0].v[k]; }
  cross3(dir, va, vb);
  normalize3(dir);
  if (dot3(dir, P[0].v) > 0) {
    const SPoint<T> t = P[1]; P[1] = P[2]; P[2] = t;
    dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2];
  }
#pragma unroll 1
  for (int guard = 0;; guard++) {
    if (guard > 100) return -1;
    md_support<T, NC, DUAL>(w, verts, A, B, dir, P[3]);
    d = dot3(P[3].v, dir);
    if (is_zero(d) || d < 0) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; return 0; }
    bool cont = false;
    cross3(va, P[1].v, P[3].v); d = dot3(va, P[0].v);
    if (d < 0 && !is_zero(d)) { P[2] = P[3]; cont = true; }
    if (!cont) {
      cross3(va, P[3].v, P[2].v); d = dot3(va, P[0].v);
      if (d < 0 && !is_zero(d)) { P[1] = P[3]; cont = true; }
    }
    if (!cont) break;
#pragma unroll
    for (int k = 0; k < 3; k++) { va[k] = P[1].v[k] - P[0].v[k]; vb[k] = P[2].v[k] - P[0].v[k]; }
    cross3(dir, va, vb);
    normalize3(dir);
  }
#pragma unroll 1
  for (int guard = 0;; guard++) {
    if (guard > 100) return -1;
    portal_dir(P, dir);
    d = dot3(dir, P[1].v);
    if (is_zero(d) || d > 0) break;
    md_support<T, NC, DUAL>(w, verts, A, B, dir, v4);
    d = dot3(v4.v, dir);
    if (!(is_zero(d) || d > 0)) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; return 0; }
    if (reach_tolerance(P, v4, dir, tol)) return -1;
    expand_portal(P, v4);
  }
#pragma unroll 1
  for (int it = 0;; it++) {
    portal_dir(P, dir);
    md_support<T, NC, DUAL>(w, verts, A, B, dir, v4);
    if (reach_tolerance(P, v4, dir, tol) || it > 50) {
      T wit[3];
      depth = sqrt(origin_tri_dist2(P[1].v, P[2].v, P[3].v, wit));
      if (is_zero(wit[0]) && is_zero(wit[1]) && is_zero(wit[2])) { pdir[0] = dir[0]; pdir[1] = dir[1]; pdir[2] = dir[2]; }
      else { pdir[0] = wit[0]; pdir[1] = wit[1]; pdir[2] = wit[2]; }
      normalize3(pdir);
      find_pos(P, pos);
      return 1;
    }
    expand_portal(P, v4);
  }
}

template <typename T, int NC> DI void mesh_shape(const Ws<T, NC>& w, const DevModel<T>& m, int g, Shape<T>& sh) {
  const int b = m.mesh_body[g];
  sh.kind = 1; sh.body = b; sh.adr = m.mesh_vertadr[g]; sh.num = m.mesh_vertnum[g];
  sh.half[0] = sh.half[1] = sh.half[2] = 0;
  T c[3] = {m.mesh_com[g][0], m.mesh_com[g][1], m.mesh_com[g][2]};
  mat_vec(sh.center, w.xmat[b], c);
#pragma unroll
  for (int k = 0; k < 3; k++) sh.center[k] += w.xpos[b][k];
}
template <typename T, int NC> DI void cube_shape(const Ws<T, NC>& w, const DevModel<T>& m, int c, Shape<T>& sh) {
  const int b = LCR_NABODY + c;
  sh.kind = 0; sh.body = b; sh.adr = 0; sh.num = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) { sh.half[k] = m.cube_size[c][k]; sh.center[k] = w.xpos[b][k]; }
}

// oriented-box overlap test (15 axes) between the body-frame bounding boxes of two mesh geoms; conservative
template <typename T, int NC>
DI bool obb_apart(const Ws<T, NC>& w, const DevModel<T>& m, int g1, int g2) {
  const int b1 = m.mesh_body[g1], b2 = m.mesh_body[g2];
  const T* RA = w.xmat[b1];
  const T* RB = w.xmat[b2];
  const T eps = (T)1e-6;
  T a[3] = {m.mesh_half[g1][0] + eps, m.mesh_half[g1][1] + eps, m.mesh_half[g1][2] + eps};
  T b[3] = {m.mesh_half[g2][0] + eps, m.mesh_half[g2][1] + eps, m.mesh_half[g2][2] + eps};
  T tw[3] = {w.gc[g2][0] - w.gc[g1][0], w.gc[g2][1] - w.gc[g1][1], w.gc[g2][2] - w.gc[g1][2]};
  T R[3][3], AR[3][3], t[3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    t[i] = tw[0] * RA[i] + tw[1] * RA[3 + i] + tw[2] * RA[6 + i];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      R[i][j] = RA[i] * RB[j] + RA[3 + i] * RB[3 + j] + RA[6 + i] * RB[6 + j];
      AR[i][j] = fabs(R[i][j]) + eps;
    }
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (fabs(t[i]) > a[i] + b[0] * AR[i][0] + b[1] * AR[i][1] + b[2] * AR[i][2]) return true;
#pragma unroll
  for (int j = 0; j < 3; j++)
    if (fabs(t[0] * R[0][j] + t[1] * R[1][j] + t[2] * R[2][j]) > a[0] * AR[0][j] + a[1] * AR[1][j] + a[2] * AR[2][j] + b[j]) return true;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const T ra = a[i1] * AR[i2][j] + a[i2] * AR[i1][j], rb = b[j1] * AR[i][j2] + b[j2] * AR[i][j1];
      if (fabs(t[i2] * R[i1][j] - t[i1] * R[i2][j]) > ra + rb) return true;
    }
  }
  return false;
}

// ---------------------------------------------------------------- box-box (boxes cA, cB: cube or static wall): SAT + clipping
template <typename T, int NC>
__device__ __noinline__ void collide_box_box(Ws<T, NC>& w, const DevModel<T>& m, int& ncon, int& nefc, int cA, int cB, const CPar<T>* par) {
  const int bA = LCR_NABODY + cA, bB = LCR_NABODY + cB;
  const T* pA = w.xpos[bA];
  const T* pB = w.xpos[bB];
  const T* RA = w.xmat[bA];
  const T* RB = w.xmat[bB];
  T hA[3] = {m.cube_size[cA][0], m.cube_size[cA][1], m.cube_size[cA][2]}, hB[3] = {m.cube_size[cB][0], m.cube_size[cB][1], m.cube_size[cB][2]};
  const int idA = box_body<NC>(cA), idB = box_body<NC>(cB);  // body ids of the contact records (-1: world)
  T t[3] = {pB[0] - pA[0], pB[1] - pA[1], pB[2] - pA[2]};
  {
    const T r = sqrt(dot3(hA, hA)) + sqrt(dot3(hB, hB));
    if (dot3(t, t) > r * r) return;
  }
  T axA[3][3], axB[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++) { axA[i][k] = RA[3 * k + i]; axB[i][k] = RB[3 * k + i]; }
  T best = (T)1e30, bestn[3] = {0, 0, 0};
  int code = -1;
#pragma unroll 1
  for (int a = 0; a < 15; a++) {
    T L[3];
    if (a < 3) { L[0] = axA[a][0]; L[1] = axA[a][1]; L[2] = axA[a][2]; }
    else if (a < 6) { L[0] = axB[a - 3][0]; L[1] = axB[a - 3][1]; L[2] = axB[a - 3][2]; }
    else {
      cross3(L, axA[(a - 6) / 3], axB[(a - 6) % 3]);
      const T n = sqrt(dot3(L, L));
      if (n < (T)1e-8) continue;
      const T inv = 1 / n;
      L[0] *= inv; L[1] *= inv; L[2] *= inv;
    }
    T rA = 0, rB = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) { rA += hA[i] * fabs(dot3(axA[i], L)); rB += hB[i] * fabs(dot3(axB[i], L)); }
    const T dist = dot3(t, L), ov = rA + rB - fabs(dist);
    if (ov < 0) return;
    const T pen = a < 6 ? ov : ov * (T)1.05 + (T)1e-9;
    if (pen < best) {
      best = pen; code = a;
#pragma unroll
      for (int k = 0; k < 3; k++) bestn[k] = dist < 0 ? -L[k] : L[k];
    }
  }
  if (code < 0) return;
  if (code >= 6) {
    const int ia = (code - 6) / 3, ib = (code - 6) % 3;
    T ea[3] = {pA[0], pA[1], pA[2]}, eb[3] = {pB[0], pB[1], pB[2]};
    for (int i = 0; i < 3; i++) {
      if (i != ia) { const T sg = dot3(axA[i], bestn) > 0 ? (T)1 : (T)-1; for (int k = 0; k < 3; k++) ea[k] += sg * hA[i] * axA[i][k]; }
      if (i != ib) { const T sg = dot3(axB[i], bestn) > 0 ? (T)-1 : (T)1; for (int k = 0; k < 3; k++) eb[k] += sg * hB[i] * axB[i][k]; }
    }
    const T* ua = axA[ia];
    const T* ub = axB[ib];
    T ww[3] = {ea[0] - eb[0], ea[1] - eb[1], ea[2] - eb[2]};
    const T uaub = dot3(ua, ub), q1 = dot3(ua, ww), q2 = dot3(ub, ww), den = 1 - uaub * uaub;
    T sa = 0, sb = 0;
    if (den > (T)1e-12) { sa = (uaub * q2 - q1) / den; sb = (q2 - uaub * q1) / den; }
    sa = clampT(sa, -hA[ia], hA[ia]); sb = clampT(sb, -hB[ib], hB[ib]);
    T pos[3];
    for (int k = 0; k < 3; k++) pos[k] = (T)0.5 * (ea[k] + sa * ua[k] + eb[k] + sb * ub[k]);
    add_contact(w, m, ncon, nefc, par, idA, idB, pos, bestn, -(best / (T)1.05));
    return;
  }
  const bool refA = code < 3;
  // explicit copies of the reference / incident box (no pointer selects into local arrays)
  T pR[3], pI[3], hR[3], hI[3], axR[3][3], axI[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    pR[i] = refA ? pA[i] : pB[i]; pI[i] = refA ? pB[i] : pA[i];
    hR[i] = refA ? hA[i] : hB[i]; hI[i] = refA ? hB[i] : hA[i];
#pragma unroll
    for (int k = 0; k < 3; k++) { axR[i][k] = refA ? axA[i][k] : axB[i][k]; axI[i][k] = refA ? axB[i][k] : axA[i][k]; }
  }
  const int ir = refA ? code : code - 3;
  T nr[3] = {refA ? bestn[0] : -bestn[0], refA ? bestn[1] : -bestn[1], refA ? bestn[2] : -bestn[2]};
  // incident face: the face of I most anti-parallel to nr
  const T dI0 = axI[0][0] * nr[0] + axI[0][1] * nr[1] + axI[0][2] * nr[2];
  const T dI1 = axI[1][0] * nr[0] + axI[1][1] * nr[1] + axI[1][2] * nr[2];
  const T dI2 = axI[2][0] * nr[0] + axI[2][1] * nr[1] + axI[2][2] * nr[2];
  // (the running |max| is tracked in its own variable: nvcc 12.9 drops the |.| of `fabs(mn)` when mn is a
  //  select of the previous candidates, which picked the wrong face)
  const T aI0 = fabs(dI0), aI1 = fabs(dI1), aI2 = fabs(dI2);
  int ii = 0;
  T mn = dI0, ma = aI0;
  if (aI1 > ma) { ma = aI1; mn = dI1; ii = 1; }
  if (aI2 > ma) { ma = aI2; mn = dI2; ii = 2; }
  const T sgI = mn > 0 ? (T)-1 : (T)1;
  T fc[3];
  for (int k = 0; k < 3; k++) fc[k] = pI[k] + sgI * hI[ii] * axI[ii][k];
  const int i1 = (ii + 1) % 3, i2 = (ii + 2) % 3;
  // the clip polygon lives in the (cold) e_jv / e_force scratch rows of the workspace: 2 x 16 x 3 values
  T (*poly)[3] = reinterpret_cast<T(*)[3]>(w.e_jv);
  T (*tmp)[3] = reinterpret_cast<T(*)[3]>(w.e_force);
  int np = 4;
  const T sx[4] = {1, -1, -1, 1}, sy[4] = {1, 1, -1, -1};
  __syncwarp();
  if (LANE == 0)
    for (int v = 0; v < 4; v++)
      for (int k = 0; k < 3; k++) poly[v][k] = fc[k] + sx[v] * hI[i1] * axI[i1][k] + sy[v] * hI[i2] * axI[i2][k];
  __syncwarp();
  const int r1 = (ir + 1) % 3, r2 = (ir + 2) % 3;
#pragma unroll 1
  for (int side = 0; side < 4; side++) {
    const T* ax = axR[side < 2 ? r1 : r2];
    const T sg = (side & 1) ? (T)-1 : (T)1, lim = hR[side < 2 ? r1 : r2];
    int nn = 0;
    for (int v = 0; v < np; v++) {  // warp-uniform: every lane walks the polygon, lane 0 stores
      const T* p = poly[v];
      const T* q = poly[(v + 1) % np];
      T rp[3] = {p[0] - pR[0], p[1] - pR[1], p[2] - pR[2]}, rq[3] = {q[0] - pR[0], q[1] - pR[1], q[2] - pR[2]};
      const T dp = sg * dot3(rp, ax) - lim, dq = sg * dot3(rq, ax) - lim;
      if (dp <= 0) { if (LANE == 0) { tmp[nn][0] = p[0]; tmp[nn][1] = p[1]; tmp[nn][2] = p[2]; } nn++; }
      if ((dp < 0 && dq > 0) || (dp > 0 && dq < 0)) {
        const T u = dp / (dp - dq);
        if (LANE == 0) for (int k = 0; k < 3; k++) tmp[nn][k] = p[k] + u * (q[k] - p[k]);
        nn++;
      }
    }
    __syncwarp();
    np = nn;
    if (LANE == 0) for (int v = 0; v < np; v++) { poly[v][0] = tmp[v][0]; poly[v][1] = tmp[v][1]; poly[v][2] = tmp[v][2]; }
    __syncwarp();
    if (np == 0) return;
  }
  int cnt = 0;
  for (int v = 0; v < np && cnt < 8; v++) {
    T rp[3] = {poly[v][0] - pR[0], poly[v][1] - pR[1], poly[v][2] - pR[2]};
    const T depth = hR[ir] - dot3(rp, nr);
    if (depth <= 0) continue;
    T pos[3] = {poly[v][0] + (T)0.5 * depth * nr[0], poly[v][1] + (T)0.5 * depth * nr[1], poly[v][2] + (T)0.5 * depth * nr[2]};
    if (add_contact(w, m, ncon, nefc, par, idA, idB, pos, bestn, -depth)) cnt++;
  }
}

// ---------------------------------------------------------------- separating-axis cache (performance only)
// A pair that was separated along direction d in the previous substep is first tested along d again; if
// max over A-B of x.d < -tol the pair is still disjoint and MPR is skipped (2 support evaluations instead of ~12).
// Entries live for one API call; round-robin replacement.  The oracle keeps the identical cache.
template <typename T, int NC> DI void sa_clear(Ws<T, NC>& w) {
  if (LANE < LCR_NSA) w.sa_key[LANE] = -1;
  if (LANE == 0) w.sa_next = 0;
  __syncwarp();
}
// ---------------------------------------------------------------- convex pairs: candidates -> jobs -> contacts
// The narrowphase of the box-mesh and mesh-mesh pairs is organised as three steps so that it can run either inline
// (fused kernel) or as its own launch with one warp per (env, slot) (phased mode: MPR calls are the most
// unevenly distributed work of mj_step -- a folded arm yields nine penetrating hull pairs -- and a per-env warp
// would hold its whole launch back):
//   1. collect_candidates : sphere + oriented-box broadphase, candidate keys in canonical (contact generation) order
//   2. narrowphase_job    : cached separating-axis test or full MPR for ONE candidate; reads the workspace only
//   3. consume_candidates : separating-axis cache updates and add_contact, in candidate order
// key: mesh-mesh pair p -> p; box c (cube, or static wall behind the cubes) vs mesh g -> LCR_KEY_CUBE + LCR_MAXMESH * c + g
#define LCR_KEY_CUBE 200

template <typename T, int NC> DI T (*cand_res(Ws<T, NC>& w))[8] { return reinterpret_cast<T(*)[8]>(w.e_w); }

// Upper bound of max over the hull of mesh g of x . dw (dw a world direction, slot/side = its cache entry) WITHOUT touching
// the vertices: with c the world centre of the hull's bounding sphere, r its radius, S the exact support value about c
// along the body-frame direction u0 recorded at the last exact evaluation, and u1 = R^T dw the direction now,
//     max x . dw  =  c . dw + max y . u1  <=  c . dw + S + min(|u1 - u0| r, sum_k |u1 - u0|_k half_k)
// (y = body-frame vertex - centre lies in the bounding sphere and in the bounding box, which share the centre).
// Translation is followed exactly, only the relative rotation since the last exact evaluation costs slack.
template <typename T, int NC>
DI T hull_support_bound(const Ws<T, NC>& w, const DevModel<T>& m, int g, const T* dw, int slot, int side) {
  const T* R = w.xmat[m.mesh_body[g]];
  T u1[3], e2 = 0, eb = 0;
  matT_vec(u1, R, dw);
#pragma unroll
  for (int k = 0; k < 3; k++) { const T e = u1[k] - w.sa_u[slot][side][k]; e2 += e * e; eb += fabs(e) * m.mesh_half[g][k]; }
  const T es = sqrt(e2) * m.mesh_rbound[g];  // y lies in the bounding sphere and in the bounding box: use the smaller slack
  return w.gc[g][0] * dw[0] + w.gc[g][1] * dw[1] + w.gc[g][2] * dw[2] + w.sa_S[slot][side] + (eb < es ? eb : es);
}

// Broadphase use of the cache (lane-local, no warp cooperation): true if the pair `key` has a cached axis that provably
// still separates it -- exact support of the box for cube pairs, hull_support_bound for hulls.  Such a pair produces no
// candidate at all, so a resting or slowly moving arm costs no narrowphase work.
template <typename T, int NC>
DI bool cached_axis_separates(const Ws<T, NC>& w, const DevModel<T>& m, int key) {
  int slot = -1;
#pragma unroll
  for (int k = 0; k < LCR_NSA; k++) if (w.sa_key[k] == key) slot = k;  // at most one entry per key
  if (slot < 0) return false;
  const T d[3] = {w.sa_dir[slot][0], w.sa_dir[slot][1], w.sa_dir[slot][2]}, nd[3] = {-d[0], -d[1], -d[2]};
  T supA;
  int gB;
  if (key >= LCR_KEY_CUBE) {
    const int c = (key - LCR_KEY_CUBE) / LCR_MAXMESH, b = LCR_NABODY + c;
    gB = (key - LCR_KEY_CUBE) % LCR_MAXMESH;
    T dl[3], pl[3], pw[3];
    matT_vec(dl, w.xmat[b], d);
#pragma unroll
    for (int k = 0; k < 3; k++) pl[k] = dl[k] >= 0 ? m.cube_size[c][k] : -m.cube_size[c][k];
    mat_vec(pw, w.xmat[b], pl);
#pragma unroll
    for (int k = 0; k < 3; k++) pw[k] += w.xpos[b][k];
    supA = dot3(pw, d);
  } else {
    supA = hull_support_bound(w, m, m.pair_g1[key], d, slot, 0);
    gB = m.pair_g2[key];
  }
  return supA + hull_support_bound(w, m, gB, nd, slot, 1) < (T)-2e-6;
}

template <typename T, int NC>
__device__ __noinline__ void collect_candidates(Ws<T, NC>& w, const DevModel<T>& m) {
  const int lane = LANE, cmask = m.collision_mask;
  if (lane < m.nmesh) {  // world centres of the mesh bounding spheres / boxes
    const int b = m.mesh_body[lane];
    T c[3] = {m.mesh_center[lane][0], m.mesh_center[lane][1], m.mesh_center[lane][2]}, t[3];
    mat_vec(t, w.xmat[b], c);
    w.gc[lane][0] = w.xpos[b][0] + t[0]; w.gc[lane][1] = w.xpos[b][1] + t[1]; w.gc[lane][2] = w.xpos[b][2] + t[2];
  }
  __syncwarp();
  int n = 0;
  if (cmask & LCR_COLLIDE_CUBE_MESH)
    for (int c = 0; c < Scene<NC>::NCUBE; c++) {
      const int bc = LCR_NABODY + c;
      bool cand = false;
      if (lane < m.nmesh) {
        T hc[3] = {m.cube_size[c][0], m.cube_size[c][1], m.cube_size[c][2]};
        const T r = m.mesh_rbound[lane] + sqrt(dot3(hc, hc));
        T d[3] = {w.gc[lane][0] - w.xpos[bc][0], w.gc[lane][1] - w.xpos[bc][1], w.gc[lane][2] - w.xpos[bc][2]};
        cand = !(dot3(d, d) > r * r);
        if (cand) cand = !cached_axis_separates(w, m, LCR_KEY_CUBE + LCR_MAXMESH * c + lane);
      }
      const unsigned mask = __ballot_sync(FULLMASK, cand);
      if (cand) {
        const int k = n + __popc(mask & ((1u << lane) - 1));
        if (k < Ws<T, NC>::MAXCAND) w.cand_key[k] = (short)(LCR_KEY_CUBE + LCR_MAXMESH * c + lane);
      }
      n += __popc(mask);
    }
  if (Scene<NC>::NWALL > 0 && (cmask & LCR_COLLIDE_WALL_MESH))
    // static boxes vs the meshes of the moving arm bodies (base_link is welded to the world like the walls): bounding
    // sphere against the axis-aligned box, then the cached axis
    for (int c = Scene<NC>::NCUBE; c < Scene<NC>::NBOX; c++) {
      const int bc = LCR_NABODY + c;
      bool cand = false;
      if (lane < m.nmesh && m.mesh_body[lane] != 0) {
        T d2 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) { const T e = fabs(w.gc[lane][k] - w.xpos[bc][k]) - m.cube_size[c][k]; if (e > 0) d2 += e * e; }
        cand = !(d2 > m.mesh_rbound[lane] * m.mesh_rbound[lane]);
        if (cand) cand = !cached_axis_separates(w, m, LCR_KEY_CUBE + LCR_MAXMESH * c + lane);
      }
      const unsigned mask = __ballot_sync(FULLMASK, cand);
      if (cand) {
        const int k = n + __popc(mask & ((1u << lane) - 1));
        if (k < Ws<T, NC>::MAXCAND) w.cand_key[k] = (short)(LCR_KEY_CUBE + LCR_MAXMESH * c + lane);
      }
      n += __popc(mask);
    }
  if (cmask & LCR_COLLIDE_MESH_MESH) {
    // pass 1: bounding spheres over all pairs, survivors compacted (in pair order) into scratch that is dead here
    // (the solver's e_jv row / the RNE temporaries); pass 2: oriented boxes and the cached-axis bound on the short list
    short* list = reinterpret_cast<short*>(w.e_jv);
    int n1 = 0;
    for (int base = 0; base < m.npair; base += 32) {
      const int p = base + lane;
      bool pass = false;
      if (p < m.npair) {
        const int g1 = m.pair_g1[p], g2 = m.pair_g2[p];
        const T r = m.mesh_rbound[g1] + m.mesh_rbound[g2];
        T d[3] = {w.gc[g1][0] - w.gc[g2][0], w.gc[g1][1] - w.gc[g2][1], w.gc[g1][2] - w.gc[g2][2]};
        pass = !(dot3(d, d) > r * r);
      }
      const unsigned mask = __ballot_sync(FULLMASK, pass);
      if (pass) list[n1 + __popc(mask & ((1u << lane) - 1))] = (short)p;
      n1 += __popc(mask);
    }
    __syncwarp();
    for (int base = 0; base < n1; base += 32) {
      bool cand = false;
      int p = 0;
      if (base + lane < n1) {
        p = list[base + lane];
        cand = !obb_apart(w, m, m.pair_g1[p], m.pair_g2[p]);
        if (cand) cand = !cached_axis_separates(w, m, p);
      }
      const unsigned mask = __ballot_sync(FULLMASK, cand);
      if (cand) {
        const int k = n + __popc(mask & ((1u << lane) - 1));
        if (k < Ws<T, NC>::MAXCAND) w.cand_key[k] = (short)p;
      }
      n += __popc(mask);
    }
  }
  if (lane == 0) w.ncand = n;  // n > Ws<T, NC>::MAXCAND: the tail is recomputed inline by consume_candidates (rare)
  __syncwarp();
}

template <typename T, int NC> DI void key_shapes(const Ws<T, NC>& w, const DevModel<T>& m, int key, Shape<T>& A, Shape<T>& B) {
  if (key >= LCR_KEY_CUBE) {
    const int c = (key - LCR_KEY_CUBE) / LCR_MAXMESH, g = (key - LCR_KEY_CUBE) % LCR_MAXMESH;
    cube_shape(w, m, c, A);
    mesh_shape(w, m, g, B);
  } else {
    mesh_shape(w, m, m.pair_g1[key], A);
    mesh_shape(w, m, m.pair_g2[key], B);
  }
}

// One candidate.  res = {code, SA, dir[3], SB | depth, dir[3], pos[3]}; code 1: penetrating (depth, dir, pos);
// 0: separated, new axis in dir; 3: still separated along the cached axis by the exact test, axis and supports
// refreshed; -1: separated without a usable axis.  For codes 0 and 3, SA = res[1] and SB = res[5] are the support
// values of the two shapes about their bounding-sphere centres along +dir / -dir (meshes only), which the cache keeps
// for hull_support_bound.  `w` is only read (it may live in HBM, or belong to another warp of the CTA).
template <typename T, int NC, bool DUAL>
__device__ __noinline__ void narrowphase_job(const Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts, int key, T* res) {
  Shape<T> A, B;
  key_shapes(w, m, key, A, B);
  const bool cube = key >= LCR_KEY_CUBE;
  const int gA = cube ? -1 : m.pair_g1[key], gB = cube ? (key - LCR_KEY_CUBE) % LCR_MAXMESH : m.pair_g2[key];
  const unsigned hit = __ballot_sync(FULLMASK, LANE < LCR_NSA && w.sa_key[LANE] == key);
  T d[3] = {0, 0, 0};
  SPoint<T> p;
  int code = -2;
  if (hit) {
    const int slot = __ffs(hit) - 1;
    d[0] = w.sa_dir[slot][0]; d[1] = w.sa_dir[slot][1]; d[2] = w.sa_dir[slot][2];
    // (the vertex-free bound along this axis already failed in the broadphase, see cached_axis_separates)
    md_support<T, NC, DUAL>(w, verts, A, B, d, p);  // exact test: two hull scans
    if (dot3(p.v, d) < (T)-1e-6) code = 3;
  }
  if (code == -2) {
    T depth = 0, pos[3] = {0, 0, 0};
    code = mpr_penetration<T, NC, DUAL>(w, verts, A, B, depth, d, pos);
    if (code != 0) {
      res[0] = (T)code; res[1] = depth;
      res[2] = d[0]; res[3] = d[1]; res[4] = d[2];
      res[5] = pos[0]; res[6] = pos[1]; res[7] = pos[2];
      return;
    }
    md_support<T, NC, DUAL>(w, verts, A, B, d, p);  // supports along the new axis, kept with it
  }
  // code 0 or 3: MPR stops at the first direction that separates, usually a grazing one.  The parts of this arm are
  // box-like and face each other across millimetre gaps, so the widest gap is (nearly) along a face normal of one of
  // the two oriented bounding boxes: take the normal whose box-vs-box gap estimate is largest, evaluate it exactly
  // (one more support pair) and keep it if it separates better -- the axis then survives many substeps of relative
  // rotation before hull_support_bound needs the vertices again.
  {
    const T* RA = w.xmat[A.body];
    const T* RB = w.xmat[B.body];
    T cA[3], cB[3], hA[3], hB[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      cA[k] = cube ? w.xpos[A.body][k] : w.gc[gA][k]; hA[k] = cube ? A.half[k] : m.mesh_half[gA][k];
      cB[k] = w.gc[gB][k]; hB[k] = m.mesh_half[gB][k];
    }
    const T t[3] = {cB[0] - cA[0], cB[1] - cA[1], cB[2] - cA[2]};
    T best = (T)-1e30, dn[3] = {0, 0, 0};
#pragma unroll 1
    for (int c = 0; c < 6; c++) {
      const T* R = c < 3 ? RA : RB;
      const int k = c % 3;
      T ax[3] = {R[k], R[3 + k], R[6 + k]};
      const T sg = dot3(t, ax) < 0 ? (T)-1 : (T)1;
      ax[0] *= sg; ax[1] *= sg; ax[2] *= sg;
      T eA = 0, eB = 0;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        eA += fabs(RA[j] * ax[0] + RA[3 + j] * ax[1] + RA[6 + j] * ax[2]) * hA[j];
        eB += fabs(RB[j] * ax[0] + RB[3 + j] * ax[1] + RB[6 + j] * ax[2]) * hB[j];
      }
      const T est = dot3(t, ax) - eA - eB;
      if (est > best) { best = est; dn[0] = ax[0]; dn[1] = ax[1]; dn[2] = ax[2]; }
    }
    SPoint<T> q;
    md_support<T, NC, DUAL>(w, verts, A, B, dn, q);
    if (-dot3(q.v, dn) > -dot3(p.v, d)) { p = q; d[0] = dn[0]; d[1] = dn[1]; d[2] = dn[2]; }
  }
  res[0] = (T)code;
  res[2] = d[0]; res[3] = d[1]; res[4] = d[2];
  T sA = 0;
  if (!cube) sA = (p.v1[0] - w.gc[gA][0]) * d[0] + (p.v1[1] - w.gc[gA][1]) * d[1] + (p.v1[2] - w.gc[gA][2]) * d[2];
  res[1] = sA;
  res[5] = -((p.v2[0] - w.gc[gB][0]) * d[0] + (p.v2[1] - w.gc[gB][1]) * d[1] + (p.v2[2] - w.gc[gB][2]) * d[2]);
  res[6] = 0; res[7] = 0;
}

template <typename T, int NC>
__device__ __noinline__ void run_jobs_inline(Ws<T, NC>& w, const DevModel<T>& m, const T* __restrict__ verts) {
  const int n = w.ncand < Ws<T, NC>::MAXCAND ? w.ncand : Ws<T, NC>::MAXCAND;
  T (*res)[8] = cand_res(w);
  for (int k = 0; k < n; k++) {
    T r[8];
    narrowphase_job<T, NC, true>(w, m, verts, w.cand_key[k], r);
    __syncwarp();
    if (LANE < 8) res[k][LANE] = r[LANE];
    __syncwarp();
  }
}

template <typename T, int NC>
DI void apply_result(Ws<T, NC>& w, const DevModel<T>& m, int& ncon, int& nefc, int key, const T* r) {
  const int code = (int)r[0];
  if (code == 2) return;
  const unsigned hit = __ballot_sync(FULLMASK, LANE < LCR_NSA && w.sa_key[LANE] == key);
  int slot = hit ? __ffs(hit) - 1 : -1;
  __syncwarp();
  if (code == 0 || code == 3) {
    if (code == 0 && slot < 0) { slot = w.sa_next; __syncwarp(); if (LANE == 0) w.sa_next = (slot + 1) % LCR_NSA; }
    if (slot >= 0 && LANE == 0) {  // (code 3 without an entry cannot happen: the job found the entry it refreshed)
      const bool cube = key >= LCR_KEY_CUBE;
      const int gB = cube ? (key - LCR_KEY_CUBE) % LCR_MAXMESH : m.pair_g2[key];
      const T d[3] = {r[2], r[3], r[4]}, nd[3] = {-r[2], -r[3], -r[4]};
      w.sa_key[slot] = (short)key;
      w.sa_dir[slot][0] = d[0]; w.sa_dir[slot][1] = d[1]; w.sa_dir[slot][2] = d[2];
      w.sa_S[slot][0] = r[1]; w.sa_S[slot][1] = r[5];
      if (!cube) matT_vec(w.sa_u[slot][0], w.xmat[m.mesh_body[m.pair_g1[key]]], d);
      matT_vec(w.sa_u[slot][1], w.xmat[m.mesh_body[gB]], nd);
    }
  } else if (slot >= 0) {
    if (LANE == 0) w.sa_key[slot] = -1;
  }
  __syncwarp();
  if (code != 1 || (r[2] == 0 && r[3] == 0 && r[4] == 0)) return;
  T dir[3] = {r[2], r[3], r[4]}, pos[3] = {r[5], r[6], r[7]};
  if (key >= LCR_KEY_CUBE) {
    const int c = (key - LCR_KEY_CUBE) / LCR_MAXMESH, g = (key - LCR_KEY_CUBE) % LCR_MAXMESH;
    add_contact(w, m, ncon, nefc, &m.par_cube_mesh[c][g], box_body<NC>(c), m.mesh_body[g], pos, dir, -r[1]);
  } else {
    add_contact(w, m, ncon, nefc, &m.par_mesh_mesh[key], m.mesh_body[m.pair_g1[key]], m.mesh_body[m.pair_g2[key]], pos, dir, -r[1]);
  }
}

// contacts of the candidates [k0, k1) whose keys satisfy cube == (key >= LCR_KEY_CUBE); candidates past Ws<T, NC>::MAXCAND
// have no stored key / result and are not processed (counted as overflow)
template <typename T, int NC>
__device__ __noinline__ void consume_candidates(Ws<T, NC>& w, const DevModel<T>& m, int& ncon, int& nefc, bool cube) {
  const int n = w.ncand < Ws<T, NC>::MAXCAND ? w.ncand : Ws<T, NC>::MAXCAND;
  T (*res)[8] = cand_res(w);
  for (int k = 0; k < n; k++) {
    const int key = w.cand_key[k];
    if ((key >= LCR_KEY_CUBE) != cube) continue;
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = res[k][j];
    apply_result(w, m, ncon, nefc, key, r);
  }
  if (!cube && w.ncand > Ws<T, NC>::MAXCAND && LANE == 0) { w.diag[4] += w.ncand - Ws<T, NC>::MAXCAND; w.ovf = 1; }
}

}  // namespace lcr
