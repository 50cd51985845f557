// lcr_render.cu -- batched image observations (reach_cube_env.py:109-112,288-292: mujoco.Renderer, cameras camera_front /
// camera_top, 240 x 320 RGB uint8) as a ray caster over the scene's convex geometry: one thread per pixel, one CTA per
// 16 x 16 tile of one camera of one env.
//
// Geometry = the floor plane, the convex hulls of the arm's visual meshes (half-space lists in the body frame), and the boxes
// (cubes, PushCubeLoop's rails).  A CTA first culls the geoms whose bounding sphere misses the cone around its tile (one warp,
// ballot -> a bit mask, so every pixel walks the survivors in the same order: images are deterministic); a pixel then clips
// its ray against the half-spaces of each survivor (entry = the latest crossing into a half-space, exit = the earliest crossing
// out), keeps the nearest entry and shades it with the scene's headlight (ambient 0.3, diffuse 0.6) and its point light
// (diffuse 0.7 at (0, 0, 3)), no specular term, no shadows; the floor carries the 0.1 m checker of the groundplane material and
// rays that leave the scene the skybox gradient.  What MuJoCo's OpenGL renderer adds on top (the concave detail of the original
// visual meshes, the translucent target / goal markers, reflectance, shadows, anti-aliasing, the edge marks of the checker) is
// not reproduced: the images are geometrically faithful (camera model, poses, hull silhouettes, colours), not pixel-identical.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/lcrsim.h"

namespace {

constexpr int kTile = 16, kMaxGeom = 32, kMaxSlot = 16;

struct Cam {
  float pos[3], rot[9], tan_half;  // rot: columns = camera x (right), y (up), z (backwards: the camera looks along -z)
};
struct RenderArgs {
  const float* poses;   // [n][nslot][12]
  const float* geoms;   // [ngeom][LCR_RENDER_GEOM_WORDS]
  const float* planes;  // [P][4]: n . x + d <= 0 inside, body frame
  uint8_t* out;         // [n][ncam][H][W][3]
  int nslot, ngeom, ncam, H, W;
  Cam cam[LCR_RENDER_MAXCAM];
};

__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__global__ void __launch_bounds__(kTile* kTile) k_render(RenderArgs A) {
  __shared__ float pose[kMaxSlot][12];
  __shared__ unsigned mask_s;
  const int env = blockIdx.z, ci = blockIdx.y, tiles_x = (A.W + kTile - 1) / kTile;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x, tid = threadIdx.y * kTile + threadIdx.x;
  const Cam& cam = A.cam[ci];
  for (int i = tid; i < A.nslot * 12; i += kTile * kTile) pose[i / 12][i % 12] = A.poses[(size_t)env * A.nslot * 12 + i];
  __syncthreads();
  const float aspect = (float)A.W / (float)A.H;
  auto ray_dir = [&](float px, float py, float* d) {  // pixel centre (px, py) in pixels -> unit world direction
    const float x = (2.0f * px / A.W - 1.0f) * cam.tan_half * aspect, y = (1.0f - 2.0f * py / A.H) * cam.tan_half;
    const float c[3] = {x, y, -1.0f};
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; k++) v[k] = cam.rot[3 * k] * c[0] + cam.rot[3 * k + 1] * c[1] + cam.rot[3 * k + 2] * c[2];
    const float inv = rsqrtf(dot3(v, v));
    d[0] = v[0] * inv; d[1] = v[1] * inv; d[2] = v[2] * inv;
  };
  // ---- tile culling: bounding sphere of every geom against the cone around the tile
  if (tid < 32) {
    float dc[3], dk[3];
    const float x0 = tx * kTile, y0 = ty * kTile, x1 = fminf(x0 + kTile, (float)A.W), y1 = fminf(y0 + kTile, (float)A.H);
    ray_dir(0.5f * (x0 + x1), 0.5f * (y0 + y1), dc);
    float cmin = 1.0f;
    const float cx[4] = {x0, x1, x0, x1}, cy[4] = {y0, y0, y1, y1};
#pragma unroll
    for (int k = 0; k < 4; k++) { ray_dir(cx[k], cy[k], dk); cmin = fminf(cmin, dot3(dc, dk)); }
    cmin = fminf(1.0f, fmaxf(cmin, 0.0f));
    const float smax = sqrtf(fmaxf(0.0f, 1.0f - cmin * cmin)), tmax = smax / fmaxf(cmin, 1e-6f);
    bool keep = false;
    if (tid < A.ngeom) {
      const float* g = A.geoms + tid * LCR_RENDER_GEOM_WORDS;
      const int slot = (int)g[1];
      const float* P = pose[slot];
      float c[3];
#pragma unroll
      for (int k = 0; k < 3; k++) c[k] = P[k] + P[3 + 3 * k] * g[4] + P[4 + 3 * k] * g[5] + P[5 + 3 * k] * g[6] - cam.pos[k];
      const float r = g[7], along = dot3(c, dc), perp = sqrtf(fmaxf(0.0f, dot3(c, c) - along * along));
      keep = along > -r && perp <= fmaxf(along, 0.0f) * tmax + r / fmaxf(cmin, 1e-6f);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (tid == 0) mask_s = m;
  }
  __syncthreads();
  const int px = tx * kTile + threadIdx.x, py = ty * kTile + threadIdx.y;
  if (px >= A.W || py >= A.H) return;
  float d[3];
  ray_dir(px + 0.5f, py + 0.5f, d);
  const float* o = cam.pos;
  float tbest = 1e30f, nrm[3] = {0, 0, 1}, rgb[3];
  int hit = -1;  // -1 sky, -2 floor, >= 0 geom
  if (d[2] < -1e-9f && o[2] > 0) { tbest = -o[2] / d[2]; hit = -2; }
  for (unsigned m = mask_s; m; m &= m - 1) {
    const int gi = __ffs((int)m) - 1;
    const float* g = A.geoms + gi * LCR_RENDER_GEOM_WORDS;
    const float* P = pose[(int)g[1]];
    float ol[3], dl[3];  // ray in the body frame
    {
      const float rel[3] = {o[0] - P[0], o[1] - P[1], o[2] - P[2]};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        ol[k] = P[3 + k] * rel[0] + P[6 + k] * rel[1] + P[9 + k] * rel[2];
        dl[k] = P[3 + k] * d[0] + P[6 + k] * d[1] + P[9 + k] * d[2];
      }
    }
    {  // bounding sphere
      const float c[3] = {g[4] - ol[0], g[5] - ol[1], g[6] - ol[2]};
      const float along = dot3(c, dl), perp2 = dot3(c, c) - along * along;
      if (perp2 > g[7] * g[7] || along + g[7] < 0 || along - g[7] > tbest) continue;
    }
    float t0 = 0.0f, t1 = tbest, n0[3] = {0, 0, 0};
    bool ok = true;
    if ((int)g[0] == 1) {  // box: three slabs
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float h = g[8 + k];
        if (fabsf(dl[k]) < 1e-12f) { ok = ok && fabsf(ol[k]) <= h; continue; }
        const float inv = 1.0f / dl[k], ta = (-h - ol[k]) * inv, tb = (h - ol[k]) * inv;
        const float tn = fminf(ta, tb), tf = fmaxf(ta, tb);
        if (tn > t0) { t0 = tn; n0[0] = n0[1] = n0[2] = 0; n0[k] = dl[k] > 0 ? -1.0f : 1.0f; }
        t1 = fminf(t1, tf);
      }
    } else {  // hull: half-space list, behind a slab test against the hull's bounding box in the body frame (rejection only)
      {
        float ta = 0.0f, tb = tbest;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const float h = g[8 + k] + 1e-5f, c = ol[k] - g[4 + k];
          if (fabsf(dl[k]) < 1e-12f) { if (fabsf(c) > h) tb = -1.0f; continue; }
          const float inv = 1.0f / dl[k], u = (-h - c) * inv, v = (h - c) * inv;
          ta = fmaxf(ta, fminf(u, v)); tb = fminf(tb, fmaxf(u, v));
        }
        if (ta > tb) continue;
      }
      const int p0 = (int)g[2], pn = (int)g[3];
      for (int p = p0; p < p0 + pn && t0 <= t1; p++) {
        const float4 pl = __ldg(reinterpret_cast<const float4*>(A.planes) + p);
        const float den = pl.x * dl[0] + pl.y * dl[1] + pl.z * dl[2], dist = pl.x * ol[0] + pl.y * ol[1] + pl.z * ol[2] + pl.w;
        if (fabsf(den) < 1e-12f) { if (dist > 0) { ok = false; break; } continue; }
        const float t = -dist / den;
        if (den < 0) { if (t > t0) { t0 = t; n0[0] = pl.x; n0[1] = pl.y; n0[2] = pl.z; } }
        else t1 = fminf(t1, t);
      }
    }
    if (!ok || t0 > t1 || t0 <= 0.0f || t0 >= tbest) continue;
    tbest = t0; hit = gi;
#pragma unroll
    for (int k = 0; k < 3; k++) nrm[k] = P[3 + 3 * k] * n0[0] + P[4 + 3 * k] * n0[1] + P[5 + 3 * k] * n0[2];
  }
  if (hit == -1) {  // skybox gradient: rgb1 = (0.3, 0.5, 0.7) at the zenith to rgb2 = 0 at and below the horizon
    const float e = fmaxf(0.0f, d[2]);
    rgb[0] = 0.3f * e; rgb[1] = 0.5f * e; rgb[2] = 0.7f * e;
  } else {
    const float hp[3] = {o[0] + tbest * d[0], o[1] + tbest * d[1], o[2] + tbest * d[2]};
    float base[3];
    if (hit == -2) {
      nrm[0] = 0; nrm[1] = 0; nrm[2] = 1;
      const int cxi = (int)floorf(hp[0] * 10.0f), cyi = (int)floorf(hp[1] * 10.0f);
      const bool odd = ((cxi + cyi) & 1) != 0;
      base[0] = odd ? 0.1f : 0.2f; base[1] = odd ? 0.2f : 0.3f; base[2] = odd ? 0.3f : 0.4f;
    } else {
      const float* g = A.geoms + hit * LCR_RENDER_GEOM_WORDS;
      base[0] = g[11]; base[1] = g[12]; base[2] = g[13];
    }
    float l2[3] = {-hp[0], -hp[1], 3.0f - hp[2]};
    const float inv = rsqrtf(dot3(l2, l2));
    const float head = fmaxf(0.0f, -dot3(nrm, d)), point = fmaxf(0.0f, dot3(nrm, l2) * inv);
    const float lum = 0.3f + 0.6f * head + 0.7f * point;
#pragma unroll
    for (int k = 0; k < 3; k++) rgb[k] = fminf(1.0f, base[k] * lum);
  }
  uint8_t* dst = A.out + ((((size_t)env * A.ncam + ci) * A.H + py) * A.W + px) * 3;
#pragma unroll
  for (int k = 0; k < 3; k++) dst[k] = (uint8_t)(rgb[k] * 255.0f + 0.5f);
}

}  // namespace

extern "C" int lcr_render(const float* d_poses, int n_envs, int n_slots, const float* d_geoms, int n_geoms, const float* d_planes,
                          const float* h_cameras, int n_cams, int height, int width, uint8_t* d_images, void* stream) {
  if (!d_poses || !d_geoms || !d_planes || !h_cameras || !d_images) return 1;
  if (n_envs <= 0 || n_slots <= 0 || n_slots > kMaxSlot || n_geoms < 0 || n_geoms > kMaxGeom || n_cams <= 0 || n_cams > LCR_RENDER_MAXCAM ||
      height <= 0 || width <= 0)
    return 1;
  RenderArgs A;
  A.poses = d_poses; A.geoms = d_geoms; A.planes = d_planes; A.out = d_images;
  A.nslot = n_slots; A.ngeom = n_geoms; A.ncam = n_cams; A.H = height; A.W = width;
  for (int c = 0; c < n_cams; c++) {
    const float* h = h_cameras + 13 * c;
    for (int k = 0; k < 3; k++) A.cam[c].pos[k] = h[k];
    for (int k = 0; k < 9; k++) A.cam[c].rot[k] = h[3 + k];
    A.cam[c].tan_half = tanf(0.5f * h[12] * 3.14159265358979f / 180.0f);
  }
  const int tiles = ((width + kTile - 1) / kTile) * ((height + kTile - 1) / kTile);
  for (int e0 = 0; e0 < n_envs; e0 += 65535) {  // gridDim.z limit
    const int ne = n_envs - e0 < 65535 ? n_envs - e0 : 65535;
    A.poses = d_poses + (size_t)e0 * n_slots * 12;
    A.out = d_images + (size_t)e0 * n_cams * height * width * 3;
    k_render<<<dim3(tiles, n_cams, ne), dim3(kTile, kTile), 0, (cudaStream_t)stream>>>(A);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
