// (T) allsight tactile renderer, batched over env x fingertip sensor frames.
//
//   K0  tac_gel_raster / tac_gel_shade   static gel: depth0 + bg_sim (once)
//   K1a tac_geom       per frame: f64 pose chain, cluster + triangle cull against the gel
//                      interior, triangle setup records, contact worklist
//   K2f tac_fill       every live frame: color = bg_real, gel_depth = 0, obs = obs_empty
//                      (pure 128-bit streaming stores; the no-contact result is exact)
//   K1b/K2/K3 tac_contact  frames with surviving triangles: tile z-buffer raster in shared
//                      memory, PBR shade, (c - bg_sim)*s, 7x7 Gaussian, + bg_real, clip ->
//                      color; depth0 - depth; then the obs pixels the dirty window touches
//
// Raster arithmetic follows the contract in oracle/raster.c / DESIGN.md "raster spec":
// f32, every operation rounded separately (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn), same
// order, so coverage and depth are bit-identical to the oracle.
#include "igi_common.cuh"
#include "../../include/igi_b200.h"

namespace {

constexpr int TW = 224, TH = 224;       // tactile image
constexpr int OBS_W = 64, OBS_H = 32;   // encoder image after crop
constexpr int MAX_LIGHTS = 8;
constexpr int HALO = 3;  // blur radius
constexpr uint64_t ZEMPTY = 0xffffffffffffffffull;

struct TacConst {
  float znear;
  int n_lights;
  float light_pos[MAX_LIGHTS][3], light_dir[MAX_LIGHTS][3], light_col[MAX_LIGHTS][3];
  float light_int[MAX_LIGHTS], light_las[MAX_LIGHTS], light_lao[MAX_LIGHTS];
  int inverse_square;
  float base[3], metallic, roughness;
  // shading constants derived from the material / lights at upload time
  float sh_a2, sh_f90, sh_f0[3], sh_cdiff_pi[3], sh_rad[MAX_LIGHTS][3];
  int gray;  // 1: material and lights are colourless -> the three channels are identical
  double cam_R[9], cam_p[3];  // camera zero pose in the sensor frame
  float gel_camx;
  double max_force, max_deformation;
  float calib_scale, clip_lo, clip_hi;
  float gauss[7];
  // conservative gel-interior distance grid (camera frame, metres)
  float grid_org[3], grid_h, grid_slack;
  int grid_n[3];
  float depth0_max;
  int hiz_levels, hiz_off[10], hiz_w[10];  // max-pyramid of depth0 (level 0 = full resolution)
  float sx0, sx1, sy0, sy1;  // dxp[0], dxp[W-1], dyp[0], dyp[H-1]: ends of the ray-slope tables
};
// The sensor constants travel with every launch as a __grid_constant__ kernel parameter (constant bank, same
// access cost as a __constant__ symbol), so two engines with different sensor yamls can share a process and
// a device and the library keeps no sensor state.  The two 224-entry ray-slope tables are device arrays owned
// by the caller (IgiTactileStatic.dxp / dyp); the kernels stage them in shared memory.
// cv2 INTER_AREA 3.5x taps as cv2 stores them (float32 of 1/3.5 and 0.5/3.5): even dst: w0 w0 w0 w1, odd: w1 w0 w0 w0
constexpr float AREA_W0 = (float)(1.0 / 3.5), AREA_W1 = (float)(0.5 / 3.5);

// ---- spec arithmetic ---------------------------------------------------------------
struct V3 { float x, y, z; };
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dot3s(V3 a, V3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }
__device__ __forceinline__ V3 cross3s(V3 a, V3 b) {
  return V3{sub(mul(a.y, b.z), mul(a.z, b.y)), sub(mul(a.z, b.x), mul(a.x, b.z)), sub(mul(a.x, b.y), mul(a.y, b.x))};
}
__device__ __forceinline__ V3 xform(const float* M, V3 v) {
  V3 o;
  o.x = add(add(add(mul(M[0], v.x), mul(M[1], v.y)), mul(M[2], v.z)), M[3]);
  o.y = add(add(add(mul(M[4], v.x), mul(M[5], v.y)), mul(M[6], v.z)), M[7]);
  o.z = add(add(add(mul(M[8], v.x), mul(M[9], v.y)), mul(M[10], v.z)), M[11]);
  return o;
}
__device__ __forceinline__ V3 gel_to_cam(const float* v, float camx) { return V3{-v[1], v[2], -sub(v[0], camx)}; }
__device__ __forceinline__ bool owns_zero(V3 n) {
  if (n.x != 0.0f) return n.x > 0.0f;
  if (n.y != 0.0f) return n.y > 0.0f;
  return n.z > 0.0f;
}
__device__ __forceinline__ float edge_fn(float dx, float dy, V3 n) { return sub(add(mul(dx, n.x), mul(dy, n.y)), n.z); }
__device__ __forceinline__ bool edge_in(float e, V3 n) { return e < 0.0f || (e == 0.0f && owns_zero(n)); }

struct __align__(16) Setup {       // 64 B record
  V3 n0, n1, n2, N;  // edge-plane normals BxC, CxA, AxB and face normal (B-A)x(C-A)
  float det;         // N.A  (< 0: front-facing)
  uint32_t bbox;     // x0 | y0<<8 | x1<<16 | y1<<24 (pixels, inclusive, conservative)
  uint32_t tri;      // face index in the (cluster-ordered) mesh face table
  uint32_t orig;     // original face index (depth-tie order)
};
static_assert(sizeof(Setup) == 64, "setup record must be 64 B");

// Camera-frame vertex normals of the triangle in the same slot of the setup list (written by tac_geom for the
// triangles it emits): shading reads them with three independent 128-bit loads instead of walking
// slot -> face -> vertex ids -> normals and rotating the interpolated normal per pixel.
struct __align__(16) NRec { float n[12]; };   // n[3*v + c] for vertex v = 0..2, n[9..11] unused
static_assert(sizeof(NRec) == 48, "normal record must be 48 B");

__device__ __forceinline__ bool make_setup(V3 A, V3 B, V3 C, Setup& s) {
  V3 E1{sub(B.x, A.x), sub(B.y, A.y), sub(B.z, A.z)};
  V3 E2{sub(C.x, A.x), sub(C.y, A.y), sub(C.z, A.z)};
  s.N = cross3s(E1, E2);
  s.det = dot3s(s.N, A);
  if (!(s.det < 0.0f)) return false;
  s.n0 = cross3s(B, C);
  s.n1 = cross3s(C, A);
  s.n2 = cross3s(A, B);
  return true;
}

// Conservative pixel bbox of the part of the triangle at depth >= znear. Returns false if empty.
__device__ bool tri_bbox(const TacConst& kc, V3 A, V3 B, V3 C, int& x0, int& y0, int& x1, int& y1) {
  const float zn = kc.znear;
  float zA = -A.z, zB = -B.z, zC = -C.z;
  if (fmaxf(zA, fmaxf(zB, zC)) < zn) return false;
  float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
  V3 P[3] = {A, B, C};
  float Z[3] = {zA, zB, zC};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int j = (i + 1) % 3;
    if (Z[i] >= zn) {
      float sx = P[i].x / Z[i], sy = P[i].y / Z[i];
      mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
    }
    if ((Z[i] >= zn) != (Z[j] >= zn)) {  // edge crosses the near plane: add the crossing point
      float t = (zn - Z[i]) / (Z[j] - Z[i]);
      float sx = (P[i].x + t * (P[j].x - P[i].x)) / zn, sy = (P[i].y + t * (P[j].y - P[i].y)) / zn;
      mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
    }
  }
  const float sx0 = kc.sx0, sx1 = kc.sx1, sy0 = kc.sy0, sy1 = kc.sy1;
  const float kx = (float)(TW - 1) / (sx1 - sx0), ky = (float)(TH - 1) / (sy1 - sy0);
  float fx0 = (mnx - sx0) * kx, fx1 = (mxx - sx0) * kx;
  float fy0 = (mxy - sy0) * ky, fy1 = (mny - sy0) * ky;
  if (fx1 < -2.0f || fy1 < -2.0f || fx0 > (float)(TW + 1) || fy0 > (float)(TH + 1)) return false;
  x0 = (int)fmaxf(floorf(fx0) - 1.0f, 0.0f);
  y0 = (int)fmaxf(floorf(fy0) - 1.0f, 0.0f);
  x1 = (int)fminf(ceilf(fx1) + 1.0f, (float)(TW - 1));
  y1 = (int)fminf(ceilf(fy1) + 1.0f, (float)(TH - 1));
  return x0 <= x1 && y0 <= y1;
}

#ifndef BB_MARGIN
#define BB_MARGIN 0.125f
#endif
// Tight box with reciprocal multiplies (geometry kernel).
__device__ __forceinline__ bool tri_bbox_fast(const TacConst& kc, V3 A, V3 B, V3 C, int& x0, int& y0, int& x1, int& y1) {
  const float zn = kc.znear;
  const float zA = -A.z, zB = -B.z, zC = -C.z;
  if (fminf(zA, fminf(zB, zC)) < zn) return tri_bbox(kc, A, B, C, x0, y0, x1, y1);  // near-plane cases: general path
  const float ra = __fdividef(1.0f, zA), rb = __fdividef(1.0f, zB), rc = __fdividef(1.0f, zC);
  const float ax = A.x * ra, ay = A.y * ra, bx = B.x * rb, by = B.y * rb, cx = C.x * rc, cy = C.y * rc;
  const float mnx = fminf(ax, fminf(bx, cx)), mxx = fmaxf(ax, fmaxf(bx, cx));
  const float mny = fminf(ay, fminf(by, cy)), mxy = fmaxf(ay, fmaxf(by, cy));
  const float sx0 = kc.sx0, sx1 = kc.sx1, sy0 = kc.sy0, sy1 = kc.sy1;
  const float kx = __fdividef((float)(TW - 1), sx1 - sx0), ky = __fdividef((float)(TH - 1), sy1 - sy0);
  const float fx0 = (mnx - sx0) * kx, fx1 = (mxx - sx0) * kx;
  const float fy0 = (mxy - sy0) * ky, fy1 = (mny - sy0) * ky;
  if (fx1 < -2.0f || fy1 < -2.0f || fx0 > (float)(TW + 1) || fy0 > (float)(TH + 1)) return false;
  // Pixel centres sit at integer coordinates here.  A centre can only be covered when it lies inside the
  // projected triangle up to the rounding of cover()'s separately-rounded edge functions and of the fast
  // reciprocals above (together < 0.01 pixel for the thin and short triangles of these meshes, see DESIGN.md
  // "tight boxes"); BB_MARGIN = 1/8 pixel keeps the box conservative with an order of magnitude to spare.  Most
  // mesh triangles are 1-3 pixels wide, so a box without the old +-1 pixel dilation has a third of the pixels,
  // and triangles that fall between pixel centres are dropped here.
  x0 = (int)fmaxf(ceilf(fx0 - BB_MARGIN), 0.0f);
  y0 = (int)fmaxf(ceilf(fy0 - BB_MARGIN), 0.0f);
  x1 = (int)fminf(floorf(fx1 + BB_MARGIN), (float)(TW - 1));
  y1 = (int)fminf(floorf(fy1 + BB_MARGIN), (float)(TH - 1));
  return x0 <= x1 && y0 <= y1;
}

// coverage + depth of one pixel; returns t (depth) or -1 when not covered / clipped
__device__ __forceinline__ float cover(float znear, const Setup& s, float dx, float dy, float& e1, float& e2, float& esum) {
  const float e0 = edge_fn(dx, dy, s.n0);
  if (!edge_in(e0, s.n0)) return -1.0f;
  e1 = edge_fn(dx, dy, s.n1);
  if (!edge_in(e1, s.n1)) return -1.0f;
  e2 = edge_fn(dx, dy, s.n2);
  if (!edge_in(e2, s.n2)) return -1.0f;
  const float den = edge_fn(dx, dy, s.N);
  const float t = __fdiv_rn(s.det, den);
  if (!(t >= znear)) return -1.0f;
  esum = add(add(e0, e1), e2);
  return t;
}


// Conservative test of one image-aligned 8x8 block (clipped to [x0,x1]x[y0,y1]) against a
// triangle: false when the block lies outside an edge plane, or when the nearest depth of the
// triangle's plane over the block is behind the farthest gel depth of the block (hzmax).  Edge
// functions and the depth denominator are affine in the ray slopes, so their extreme over the
// block sits at a corner; a relative margin keeps the bound conservative against the
// separately-rounded per-pixel arithmetic of cover().
__device__ __forceinline__ bool block_may_hit(const Setup& s, const float* __restrict__ tdx, const float* __restrict__ tdy,
                                              int x0, int x1, int y0, int y1, float hzmax) {
  const float dxa = tdx[x0], dxb = tdx[x1], dya = tdy[y0], dyb = tdy[y1];
  auto lo_bound = [&](V3 n) {  // lower bound of dx*n.x + dy*n.y - n.z over the block
    const float ax = fminf(dxa * n.x, dxb * n.x), ay = fminf(dya * n.y, dyb * n.y);
    const float m = 1e-5f * (fmaxf(fabsf(dxa * n.x), fabsf(dxb * n.x)) +
                             fmaxf(fabsf(dya * n.y), fabsf(dyb * n.y)) + fabsf(n.z));
    return ax + ay - n.z - m;
  };
  if (lo_bound(s.n0) > 0.0f || lo_bound(s.n1) > 0.0f || lo_bound(s.n2) > 0.0f) return false;
  const float den_lo = lo_bound(s.N);  // most negative denominator -> nearest depth
  if (!(den_lo < 0.0f)) return false;  // plane never faces these rays
  const float tmin = (s.det / den_lo) * (1.0f - 1e-5f);
  return tmin < hzmax;
}

__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ V3 normalize(V3 v) {
  float l = sqrtf(v.x * v.x + v.y * v.y + v.z * v.z);
  if (l > 0.0f) { float r = 1.0f / l; v.x *= r; v.y *= r; v.z *= r; }
  return v;
}

// single-MUFU approximations (flush-to-zero): the shading result feeds an 8-bit quantiser
__device__ __forceinline__ float fast_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// pyrender mesh.frag (metallic-roughness, spot lights) -> 8-bit UNORM rgb.
// NCH = 1 when material and lights are colourless (r == g == b), else 3.  Uses the fast
// reciprocal / rsqrt / exp2-log2 paths: the result feeds an 8-bit quantiser and is compared with
// the oracle within 1/255.
// NL = number of lights when known at compile time (the loop unrolls and the light constants become
// immediate constant-bank operands), 0 = kc.n_lights.
// EARLY (callers whose 32 lanes are converged): the light loop stops once every lane's sum has reached 1.0 in every
// channel.  Each light adds a non-negative term (n.l >= 0.001, attenuation >= 0, BRDF terms >= 0) and the output
// clamps at 1.0 -> 255, so the result is bit-identical; it only skips work for fragments that clip (every
// fragment under the inverse-square light model with the shipped intensities).
// `voting` = false: this lane only keeps the warp converged (its result is discarded) and votes "clipped".
template <int NCH, int NL, bool EARLY = false>
__device__ __forceinline__ void shade_t(const TacConst& kc, V3 p, V3 n, uint8_t* rgb, bool voting = true) {
  const float A2_PI = kc.sh_a2 * 0.31830988618379067f;
  const float ipl = fast_rsqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  const V3 v{-p.x * ipl, -p.y * ipl, -p.z * ipl};
  const float a2 = kc.sh_a2;
  const float nv_raw = n.x * v.x + n.y * v.y + n.z * v.z;
  const float nv = clampf(nv_raw, 0.001f, 1.0f);
  const float gv = nv + fast_sqrt(a2 + (1.0f - a2) * (nv * nv));   // 2 nv / gv = Smith term of the view direction
  float col[NCH];
#pragma unroll
  for (int k = 0; k < NCH; ++k) col[k] = 0.f;
  const int nl_count = NL ? NL : kc.n_lights;
#pragma unroll
  for (int i = 0; i < nl_count; ++i) {
    const V3 L{kc.light_pos[i][0] - p.x, kc.light_pos[i][1] - p.y, kc.light_pos[i][2] - p.z};
    const float il = fast_rsqrt(L.x * L.x + L.y * L.y + L.z * L.z);
    const V3 l{L.x * il, L.y * il, L.z * il};
    // half vector h = (l + v)/|l + v| with |l + v|^2 = 2 + 2 v.l, so n.h and v.h need no vector h
    const float vl = v.x * l.x + v.y * l.y + v.z * l.z;
    const float ih = fast_rsqrt(fmaxf(2.0f + 2.0f * vl, 1e-20f));
    const float nl_raw = n.x * l.x + n.y * l.y + n.z * l.z;
    const float nl = clampf(nl_raw, 0.001f, 1.0f);
    const float nh = clampf((nl_raw + nv_raw) * ih, 0.001f, 1.0f);
    const float vh = clampf((vl + 1.0f) * ih, 0.001f, 1.0f);
    const float cd = -(kc.light_dir[i][0] * l.x + kc.light_dir[i][1] * l.y + kc.light_dir[i][2] * l.z);
    float att = __saturatef(cd * kc.light_las[i] + kc.light_lao[i]);
    att = att * att;
    if (kc.inverse_square) att = att * il * il;
    const float w = 1.0f - vh;
    const float w2 = w * w, fw = w2 * w2 * w;
    // specular = G_l G_v D / (4 nl nv) with G_x = 2 x / g_x  ->  D / (g_l g_v)
    const float gl = nl + fast_sqrt(a2 + (1.0f - a2) * (nl * nl));
    const float f = (nh * a2 - nh) * nh + 1.0f;
    const float sp = A2_PI * fast_rcp(f * f * gl * gv);
    const float na = nl * att;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const float F = kc.sh_f0[k] + (kc.sh_f90 - kc.sh_f0[k]) * fw;
      col[k] += na * kc.sh_rad[i][k] * ((1.0f - F) * kc.sh_cdiff_pi[k] + F * sp);
    }
    if (EARLY) {
      bool clipped = true;
#pragma unroll
      for (int k = 0; k < NCH; ++k) clipped = clipped && (col[k] >= 1.0f || !voting);
      if (__all_sync(0xffffffffu, clipped)) break;
    }
  }
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const float o = __saturatef(fast_ex2(fast_lg2(col[k]) * (1.0f / 2.2f)));
    rgb[k] = (uint8_t)(o * 255.0f + 0.5f);
  }
  if (NCH == 1) rgb[1] = rgb[2] = rgb[0];
}
__device__ __forceinline__ void shade(const TacConst& kc, V3 p, V3 n, uint8_t* rgb) {
  if (kc.gray) shade_t<1, 0>(kc, p, n, rgb); else shade_t<3, 0>(kc, p, n, rgb);
}

// ---- K0: static gel -------------------------------------------------------------------
__global__ void tac_gel_raster(const __grid_constant__ TacConst kc, const float* __restrict__ dxp, const float* __restrict__ dyp,
                               const float* __restrict__ gel_tris, int G, unsigned long long* __restrict__ zbuf) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const float* t = gel_tris + (size_t)g * 9;
  V3 A = gel_to_cam(t, kc.gel_camx), B = gel_to_cam(t + 3, kc.gel_camx), C = gel_to_cam(t + 6, kc.gel_camx);
  Setup s;
  if (!make_setup(A, B, C, s)) return;
  int x0, y0, x1, y1;
  if (!tri_bbox(kc, A, B, C, x0, y0, x1, y1)) return;
  for (int py = y0; py <= y1; ++py)
    for (int px = x0; px <= x1; ++px) {
      float e1, e2, es;
      const float tt = cover(kc.znear, s, __ldg(dxp + px), __ldg(dyp + py), e1, e2, es);
      if (tt < 0.0f) continue;
      const unsigned long long key = ((unsigned long long)__float_as_uint(tt) << 32) | (uint32_t)g;
      atomicMin(zbuf + (size_t)py * TW + px, key);
    }
}

__global__ void tac_gel_shade(const __grid_constant__ TacConst kc, const float* __restrict__ dxp, const float* __restrict__ dyp,
                              const float* __restrict__ gel_tris, const unsigned long long* __restrict__ zbuf,
                              float* __restrict__ depth0, uint8_t* __restrict__ bg_sim) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= TW * TH) return;
  const unsigned long long key = zbuf[i];
  if (key == ZEMPTY) {
    depth0[i] = 0.0f;
    bg_sim[3 * i] = bg_sim[3 * i + 1] = bg_sim[3 * i + 2] = 255;
    return;
  }
  const float t = __uint_as_float((uint32_t)(key >> 32));
  const int g = (int)(key & 0xffffffffu);
  depth0[i] = t;
  const float* tp = gel_tris + (size_t)g * 9;
  V3 A = gel_to_cam(tp, kc.gel_camx), B = gel_to_cam(tp + 3, kc.gel_camx), C = gel_to_cam(tp + 6, kc.gel_camx);
  V3 E1{sub(B.x, A.x), sub(B.y, A.y), sub(B.z, A.z)}, E2{sub(C.x, A.x), sub(C.y, A.y), sub(C.z, A.z)};
  V3 n = normalize(cross3s(E1, E2));
  const int px = i % TW, py = i / TW;
  V3 p{mul(__ldg(dxp + px), t), mul(__ldg(dyp + py), t), -t};
  shade(kc, p, n, bg_sim + 3 * i);
}

// ---- K2f: fill -----------------------------------------------------------------------------
struct FillArgs {
  const uint8_t* bg_real;    // (n_bg, TH, TW, 3)
  const int32_t* bg_id;      // (F) index into bg_real
  const float* obs_empty;    // (OBS_H*OBS_W)
  const int32_t* counts;     // (F) -1 => frame not updated
  uint8_t* color;            // (F, TH, TW, 3) or null
  float* gel_depth;          // (F, TH, TW) or null
  float* obs;                // frame (e,n) at obs + e*obs_env_stride + n*obs_sensor_stride
  int64_t obs_env_stride, obs_sensor_stride;
  int sensors_per_env;
  int n_frames;
  int parts;                 // bit 0 colour, bit 1 gel_depth, bit 2 obs
};
constexpr int FILL_BLOCK = 256;
constexpr int FILL_PARTS = 4;  // CTAs per frame

// The no-contact result of frame f (exact: diff = 0 => colour = bg_real, gel_depth = 0, obs = obs_empty),
// written with 128-bit streaming stores by threads [first, first + stride, ...).
__device__ __forceinline__ void fill_frame(const FillArgs& a, int f, int first, int stride) {
  if (a.color && (a.parts & 1)) {
    constexpr int NV = TW * TH * 3 / 16;  // 9408 uint4
    const uint4* src = reinterpret_cast<const uint4*>(a.bg_real + (size_t)a.bg_id[f] * TW * TH * 3);
    uint4* dst = reinterpret_cast<uint4*>(a.color + (size_t)f * TW * TH * 3);
#pragma unroll 4
    for (int i = first; i < NV; i += stride) __stcs(dst + i, __ldg(src + i));
  }
  if (a.gel_depth && (a.parts & 2)) {
    constexpr int NV = TW * TH * 4 / 16;  // 12544 uint4
    uint4* dst = reinterpret_cast<uint4*>(a.gel_depth + (size_t)f * TW * TH);
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll 4
    for (int i = first; i < NV; i += stride) __stcs(dst + i, z);
  }
  if (a.parts & 4) {
    constexpr int NV = OBS_W * OBS_H * 4 / 16;  // 512 float4
    const float4* src = reinterpret_cast<const float4*>(a.obs_empty);
    float4* dst = reinterpret_cast<float4*>(a.obs + (size_t)(f / a.sensors_per_env) * a.obs_env_stride +
                                            (size_t)(f % a.sensors_per_env) * a.obs_sensor_stride);
    for (int i = first; i < NV; i += stride) dst[i] = __ldg(src + i);
  }
}

__global__ void __launch_bounds__(FILL_BLOCK) tac_fill(FillArgs a) {
  const int f = blockIdx.x / FILL_PARTS, part = blockIdx.x % FILL_PARTS;
  if (a.counts[f] < 0) return;
  fill_frame(a, f, part * FILL_BLOCK + threadIdx.x, FILL_PARTS * FILL_BLOCK);
}

// ---- K1a: geometry ----------------------------------------------------------------------
struct MeshInfo { int face_off, n_faces, cl_off, n_cl; };
struct Cluster { float cx, cy, cz, r; int first, count, pad0, pad1; };

struct GeomArgs {
  const float* fpos_n[8];    // sensor n's positions at fpos_n[n] + env*fpos_stride (frame f = env*S + sensor)
  const float* fquat_n[8];   // sensor n's quaternions (xyzw) at fquat_n[n] + env*fquat_stride
  int64_t fpos_stride, fquat_stride;
  const float* plug_pos;     // (N,3)
  const float* plug_quat;    // (N,4)
  const float* force;        // (F) or null -> force_const
  const uint8_t* update;     // (N) or null
  const uint8_t* update2;    // (N) or null, ANDed with update
  const int32_t* mesh_id;    // (N)
  const MeshInfo* meshes;
  const Cluster* clusters;
  const float* verts;        // (nv,3) all meshes
  const int32_t* faces;      // (nf,3) global vertex ids, cluster order
  const int32_t* face_orig;  // (nf)
  const float* grid;         // distance grid
  const float* hiz;          // depth0 max-pyramid
  const float* depth0;       // (TH,TW)
  const float* dxp;          // (TW) ray slopes per column
  const float* dyp;          // (TH) ray slopes per row
  float* M_out;              // (F,12)
  Setup* setups;             // (F, kmax)
  const float* vnorm;        // (nv,3) vertex normals, object frame
  NRec* nrecs;               // (F, kmax) camera-frame vertex normals of the emitted triangles
  int32_t* counts;           // (F)   surviving triangles (may exceed kmax -> overflow)
  int32_t* bbox;             // (F,4) dirty window x0,y0,x1,y1
  int32_t* worklist;         // (F)
  int32_t* work_n;           // (1)
  int32_t* overflow;         // (2) sticky overflow flag, largest per-frame triangle count seen
  int sensors_per_env, kmax;
  float force_const;
  int fused_fill;            // 1: this kernel also writes (fill.parts of) the frame's no-contact result
  int list_all;              // 1: every live frame goes to the worklist (tac_contact fills the other parts)
  FillArgs fill;
};

__device__ __forceinline__ float grid_lower_bound(const TacConst& kc, const float* __restrict__ grid, float x, float y, float z) {
  // lower bound of the distance from (x,y,z) (camera frame, depth = -z) to the gel interior
  const float gx = x, gy = y, gz = -z;
  const float lo0 = kc.grid_org[0], lo1 = kc.grid_org[1], lo2 = kc.grid_org[2];
  const float h = kc.grid_h;
  const float hi0 = lo0 + h * kc.grid_n[0], hi1 = lo1 + h * kc.grid_n[1], hi2 = lo2 + h * kc.grid_n[2];
  const float cx = clampf(gx, lo0, hi0), cy = clampf(gy, lo1, hi1), cz = clampf(gz, lo2, hi2);
  const float ddx = gx - cx, ddy = gy - cy, ddz = gz - cz;
  const float dbox = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
  // cell index by reciprocal multiply: a point within an ulp of a cell face may land in the
  // neighbouring cell, which the slack (> one cell diagonal) already covers
  const float ih = __fdividef(1.0f, h);
  int ix = min(max((int)((cx - lo0) * ih), 0), kc.grid_n[0] - 1);
  int iy = min(max((int)((cy - lo1) * ih), 0), kc.grid_n[1] - 1);
  int iz = min(max((int)((cz - lo2) * ih), 0), kc.grid_n[2] - 1);
  const float d = grid[((size_t)iz * kc.grid_n[1] + iy) * kc.grid_n[0] + ix];
  return fmaxf(dbox, d - kc.grid_slack - dbox);
}

// Optional per-phase cycle counters (build with -DCT_PROFILE; read with igi_debug_read_prof).
#ifdef CT_PROFILE
__device__ unsigned long long g_ct_prof[32];
#define CT_T(slot)                                                     \
  do {                                                                 \
    if (threadIdx.x == 0) {                                            \
      const long long now__ = clock64();                               \
      atomicAdd(&g_ct_prof[slot], (unsigned long long)(now__ - t_prof)); \
      t_prof = now__;                                                  \
    }                                                                  \
  } while (0)
#define CT_COUNT(slot, v) do { if ((v) != 0) atomicAdd(&g_ct_prof[slot], (unsigned long long)(v)); } while (0)
#else
#define CT_T(slot) do { } while (0)
#define CT_COUNT(slot, v) do { } while (0)
#endif

constexpr int GEOM_BLOCK = 128;
constexpr int GEOM_MAX_CL = 512;
constexpr int GEOM_ROUND = 256;        // faces prepared per round (2 per thread); survivors are queued in shared memory
constexpr int GEOM_WARPS = GEOM_BLOCK / 32;

#ifndef GEOM_MIN_CTAS
#define GEOM_MIN_CTAS 9   // 56 registers: 9 CTAs per SM (no bound: 88 regs, 5 CTAs, 0.82 ms; 8: 0.68 ms; 9: 0.65 ms)
#endif
__global__ void __launch_bounds__(GEOM_BLOCK, GEOM_MIN_CTAS) tac_geom(const __grid_constant__ TacConst kc, GeomArgs a) {
  __shared__ float sM[12];
  __shared__ int s_cl[GEOM_MAX_CL];
  __shared__ __align__(16) Setup s_q[GEOM_ROUND];   // survivors of the cheap culls of this round
  __shared__ int s_qoff[GEOM_ROUND];                // their 8x8-block counts -> exclusive prefix
  __shared__ uint32_t s_vmask[GEOM_ROUND][2];       // block columns / rows in which the triangle may be visible
  __shared__ int s_wsum[GEOM_WARPS];
  __shared__ float s_dxp[TW], s_dyp[TH];  // ray-slope tables (per-lane indexing would serialise in the constant cache)
  __shared__ int s_ncl, s_count, s_nq[2];  // queue length, double-buffered by round parity
  __shared__ int s_bb[4];
  const int f = blockIdx.x;
  const int env = f / a.sensors_per_env;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((a.update && !a.update[env]) || (a.update2 && !a.update2[env])) {
    if (tid == 0) a.counts[f] = -1;  // frame not rendered this step
    return;
  }
  for (int i = tid; i < TW; i += GEOM_BLOCK) { s_dxp[i] = __ldg(a.dxp + i); s_dyp[i] = __ldg(a.dyp + i); }
  // Fused fill: warps 1.. stream the frame's no-contact result out while lane 0 of warp 0 runs the serial
  // f64 pose chain; the stores drain in the background of the culling work below.  tac_contact, the next
  // kernel on the stream, rewrites the dirty window.
  if (a.fused_fill && warp > 0) fill_frame(a.fill, f, tid - 32, GEOM_BLOCK - 32);
  if (tid == 0) {
    // ---- pose chain in f64 (xyzquat_to_tf_numpy, update_camera_pose_from_matrix, adjust_with_force)
    double q[4], R[9], Ro[9];
    const int sensor = f - env * a.sensors_per_env;
    const float* fq = a.fquat_n[sensor] + (size_t)env * a.fquat_stride;
    const float* fp = a.fpos_n[sensor] + (size_t)env * a.fpos_stride;
    const float* oq = a.plug_quat + (size_t)env * 4;
    const float* op = a.plug_pos + (size_t)env * 3;
    auto q2m = [](const float* qq, double* m) {
      double x = qq[0], y = qq[1], z = qq[2], w = qq[3];
      const double n = sqrt(x * x + y * y + z * z + w * w);
      x /= n; y /= n; z /= n; w /= n;
      const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
      const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
      m[0] = x2 - y2 - z2 + w2; m[1] = 2 * (xy - zw);        m[2] = 2 * (xz + yw);
      m[3] = 2 * (xy + zw);     m[4] = -x2 + y2 - z2 + w2;   m[5] = 2 * (yz - xw);
      m[6] = 2 * (xz - yw);     m[7] = 2 * (yz + xw);        m[8] = -x2 - y2 + z2 + w2;
    };
    (void)q;
    q2m(fq, R);
    q2m(oq, Ro);
    // camera world pose = T_finger * cam_zero
    double Rc[9], pc[3];
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j)
        Rc[3 * i + j] = R[3 * i] * kc.cam_R[j] + R[3 * i + 1] * kc.cam_R[3 + j] + R[3 * i + 2] * kc.cam_R[6 + j];
      pc[i] = R[3 * i] * kc.cam_p[0] + R[3 * i + 1] * kc.cam_p[1] + R[3 * i + 2] * kc.cam_p[2] + (double)fp[i];
    }
    const double force = a.force ? (double)a.force[f] : (double)a.force_const;
    const double offset = fmin(kc.max_force, force) / kc.max_force;
    double dir[3] = {pc[0] - (double)op[0], pc[1] - (double)op[1], pc[2] - (double)op[2]};
    const double nrm = sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]) + 1e-6;
    double po[3];
    for (int i = 0; i < 3; ++i) po[i] = (double)op[i] + offset * kc.max_deformation * (dir[i] / nrm);
    // M = inv(cam_world) * [Ro, po]
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j)
        sM[4 * i + j] = (float)(Rc[i] * Ro[j] + Rc[3 + i] * Ro[3 + j] + Rc[6 + i] * Ro[6 + j]);
      sM[4 * i + 3] = (float)(Rc[i] * (po[0] - pc[0]) + Rc[3 + i] * (po[1] - pc[1]) + Rc[6 + i] * (po[2] - pc[2]));
    }
    s_ncl = 0;
    s_count = 0;
    s_nq[0] = 0; s_nq[1] = 0;
    s_bb[0] = TW; s_bb[1] = TH; s_bb[2] = -1; s_bb[3] = -1;
  }
  __syncthreads();
  if (tid < 12) a.M_out[(size_t)f * 12 + tid] = sM[tid];
  const MeshInfo mi = a.meshes[a.mesh_id[env]];
  // ---- cluster cull: bounding sphere against the gel-interior distance grid
  for (int c = tid; c < mi.n_cl; c += GEOM_BLOCK) {
    const Cluster cl = a.clusters[mi.cl_off + c];
    V3 cc = xform(sM, V3{cl.cx, cl.cy, cl.cz});
    if (grid_lower_bound(kc, a.grid, cc.x, cc.y, cc.z) <= cl.r) {
      const int slot = atomicAdd(&s_ncl, 1);
      if (slot < GEOM_MAX_CL) s_cl[slot] = mi.cl_off + c;
    }
  }
  __syncthreads();
  const int ncl = min(s_ncl, GEOM_MAX_CL);
  if (tid == 0) { CT_COUNT(16, ncl); CT_COUNT(17, mi.n_cl); }
  Setup* out = a.setups + (size_t)f * a.kmax;
  NRec* nout = a.nrecs + (size_t)f * a.kmax;
  const float* hz3 = a.hiz + kc.hiz_off[3];
  const int hw3 = kc.hiz_w[3];
  auto emit = [&](Setup& s, int face, int x0, int y0, int x1, int y1) {
    s.bbox = (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
    s.tri = (uint32_t)face;
    s.orig = (uint32_t)a.face_orig[face];
    const int slot = atomicAdd(&s_count, 1);
    if (slot < a.kmax) {
      out[slot] = s;
      // rotate the three vertex normals into the camera frame once per emitted triangle
      const int vi[3] = {a.faces[3 * face], a.faces[3 * face + 1], a.faces[3 * face + 2]};
      float r[12];
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const float nx = __ldg(a.vnorm + 3 * vi[v]), ny = __ldg(a.vnorm + 3 * vi[v] + 1), nz = __ldg(a.vnorm + 3 * vi[v] + 2);
        r[3 * v + 0] = sM[0] * nx + sM[1] * ny + sM[2] * nz;
        r[3 * v + 1] = sM[4] * nx + sM[5] * ny + sM[6] * nz;
        r[3 * v + 2] = sM[8] * nx + sM[9] * ny + sM[10] * nz;
      }
      float4* dst = reinterpret_cast<float4*>(nout + slot);
      dst[0] = make_float4(r[0], r[1], r[2], r[3]);
      dst[1] = make_float4(r[4], r[5], r[6], r[7]);
      dst[2] = make_float4(r[8], 0.f, 0.f, 0.f);
    }
    atomicMin(&s_bb[0], x0); atomicMin(&s_bb[1], y0); atomicMax(&s_bb[2], x1); atomicMax(&s_bb[3], y1);
  };
  // transform + setup + the cheap culls of one face; false = culled
  auto prepare = [&](int face, Setup& s, int& x0, int& y0, int& x1, int& y1) {
    const int i0 = a.faces[3 * face], i1 = a.faces[3 * face + 1], i2 = a.faces[3 * face + 2];
    V3 A = xform(sM, V3{a.verts[3 * i0], a.verts[3 * i0 + 1], a.verts[3 * i0 + 2]});
    V3 B = xform(sM, V3{a.verts[3 * i1], a.verts[3 * i1 + 1], a.verts[3 * i1 + 2]});
    V3 C = xform(sM, V3{a.verts[3 * i2], a.verts[3 * i2 + 1], a.verts[3 * i2 + 2]});
    // face normal + facing first; the three edge planes only for triangles that survive the culls
    V3 E1{sub(B.x, A.x), sub(B.y, A.y), sub(B.z, A.z)};
    V3 E2{sub(C.x, A.x), sub(C.y, A.y), sub(C.z, A.z)};
    s.N = cross3s(E1, E2);
    s.det = dot3s(s.N, A);
    CT_COUNT(18, 1);
    if (!(s.det < 0.0f)) return false;
    CT_COUNT(19, 1);
    // bounding sphere of the triangle around its centroid against the gel interior
    const float gx = (A.x + B.x + C.x) * (1.0f / 3.0f), gy = (A.y + B.y + C.y) * (1.0f / 3.0f),
                gz = (A.z + B.z + C.z) * (1.0f / 3.0f);
    auto d2 = [&](V3 P) { return (P.x - gx) * (P.x - gx) + (P.y - gy) * (P.y - gy) + (P.z - gz) * (P.z - gz); };
    const float rr = sqrtf(fmaxf(d2(A), fmaxf(d2(B), d2(C)))) * 1.0001f + 1e-7f;
    if (grid_lower_bound(kc, a.grid, gx, gy, gz) > rr) return false;
    CT_COUNT(20, 1);
    if (!tri_bbox_fast(kc, A, B, C, x0, y0, x1, y1)) return false;
    CT_COUNT(21, 1);
    // hierarchical-Z: nearest possible fragment vs the farthest gel depth under the box
    int L = 0;
    while (L < kc.hiz_levels - 1 && (((x1 >> L) - (x0 >> L)) > 1 || ((y1 >> L) - (y0 >> L)) > 1)) ++L;
    const float* hz = a.hiz + kc.hiz_off[L];
    const int hw = kc.hiz_w[L];
    const int ax = x0 >> L, bx = x1 >> L, ay = y0 >> L, by = y1 >> L;
    const float m = fmaxf(fmaxf(hz[ay * hw + ax], hz[ay * hw + bx]), fmaxf(hz[by * hw + ax], hz[by * hw + bx]));
    const float zmin = fmaxf(fminf(-A.z, fminf(-B.z, -C.z)), kc.znear);
    if (zmin > m) return false;
    CT_COUNT(22, 1);
    s.n0 = cross3s(B, C);
    s.n1 = cross3s(C, A);
    s.n2 = cross3s(A, B);
    return true;
  };

  // ---- faces of the surviving clusters, GEOM_ROUND at a time.  Per round: (1) every thread prepares two
  // faces (transform, facing, gel distance, screen box, hi-z) and queues the survivors' setup records;
  // (2) the queued triangles' 8x8 image blocks are flattened into one item list (block scan of the block
  // counts) so that every thread tests one (triangle, block) pair against the gel whatever the triangle
  // sizes are; (3) triangles with a block in which they may be visible are emitted with the box of
  // those blocks.
  const int n_faces_padded = ncl * 64;
  for (int r0 = 0, par = 0; r0 < n_faces_padded; r0 += GEOM_ROUND, par ^= 1) {
    for (int it = r0 + tid; it < min(r0 + GEOM_ROUND, n_faces_padded); it += GEOM_BLOCK) {
      const int2 cl = *reinterpret_cast<const int2*>(&a.clusters[s_cl[it >> 6]].first);  // first, count
      const int k = it & 63;
      if (k >= cl.y) continue;
      const int face = cl.x + k;
      Setup s;
      int x0, y0, x1, y1;
      if (!prepare(face, s, x0, y0, x1, y1)) continue;
      s.bbox = (uint32_t)x0 | ((uint32_t)y0 << 8) | ((uint32_t)x1 << 16) | ((uint32_t)y1 << 24);
      s.tri = (uint32_t)face;
      const int slot = atomicAdd(&s_nq[par], 1);
      s_q[slot] = s;
      s_qoff[slot] = ((x1 >> 3) - (x0 >> 3) + 1) * ((y1 >> 3) - (y0 >> 3) + 1);
      s_vmask[slot][0] = 0u; s_vmask[slot][1] = 0u;
    }
    __syncthreads();
    // The counter of the NEXT round is the other one, so a thread that runs ahead cannot disturb this read;
    // this one is reused two rounds later, after at least one more barrier.
    const int nq = s_nq[par];
    if (nq == 0) continue;   // uniform: nothing was queued, the counter is still 0
    // exclusive scan of the block counts (two queue entries per thread)
    int total;
    {
      const int e0 = 2 * tid, e1 = 2 * tid + 1;
      const int c0 = e0 < nq ? s_qoff[e0] : 0, c1 = e1 < nq ? s_qoff[e1] : 0;
      int incl = c0 + c1;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (lane == 31) s_wsum[warp] = incl;
      __syncthreads();
      int woff = 0;
      total = 0;
#pragma unroll
      for (int w = 0; w < GEOM_WARPS; ++w) {
        const int t = s_wsum[w];
        if (w < warp) woff += t;
        total += t;
      }
      const int base = woff + incl - c0 - c1;
      if (e0 < nq) s_qoff[e0] = base;
      if (e1 < nq) s_qoff[e1] = base + c0;
    }
    __syncthreads();
    for (int it = tid; it < total; it += GEOM_BLOCK) {
      int lo = 0, hi = nq - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_qoff[mid] <= it) lo = mid; else hi = mid - 1;
      }
      const Setup& s = s_q[lo];
      const int b = it - s_qoff[lo];
      const uint32_t bb = s.bbox;
      const int x0 = (int)(bb & 255u), y0 = (int)((bb >> 8) & 255u), x1 = (int)((bb >> 16) & 255u), y1 = (int)(bb >> 24);
      const int gbx0 = x0 >> 3, nbx = (x1 >> 3) - gbx0 + 1;
      const int by = b / nbx;
      const int gy = (y0 >> 3) + by, gx = gbx0 + b - by * nbx;
      const int cx0 = max(gx * 8, x0), cx1 = min(gx * 8 + 7, x1), cy0 = max(gy * 8, y0), cy1 = min(gy * 8 + 7, y1);
      if (block_may_hit(s, s_dxp, s_dyp, cx0, cx1, cy0, cy1, hz3[gy * hw3 + gx])) {
        atomicOr(&s_vmask[lo][0], 1u << gx);
        atomicOr(&s_vmask[lo][1], 1u << gy);
      }
    }
    __syncthreads();
    for (int q = tid; q < nq; q += GEOM_BLOCK) {
      const uint32_t mx = s_vmask[q][0], my = s_vmask[q][1];
      if (mx == 0u) continue;
      Setup s = s_q[q];
      const uint32_t bb = s.bbox;
      const int x0 = (int)(bb & 255u), y0 = (int)((bb >> 8) & 255u), x1 = (int)((bb >> 16) & 255u), y1 = (int)(bb >> 24);
      const int vx0 = max((__ffs(mx) - 1) * 8, x0), vx1 = min((31 - __clz(mx)) * 8 + 7, x1);
      const int vy0 = max((__ffs(my) - 1) * 8, y0), vy1 = min((31 - __clz(my)) * 8 + 7, y1);
      emit(s, (int)s.tri, vx0, vy0, vx1, vy1);
    }
    if (tid == 0) s_nq[par] = 0;
    __syncthreads();   // the queue and its counter are free again
  }
  __syncthreads();
  if (tid == 0) {
    const int n = s_count;
    a.counts[f] = min(n, a.kmax);
    if (n > a.kmax || s_ncl > GEOM_MAX_CL) atomicExch(a.overflow, 1);
    if (n > 0) atomicMax(a.overflow + 1, n);   // high-water mark of the triangle lists: the host sizes `kmax` from it
    if (n > 0) {
      a.bbox[4 * f + 0] = s_bb[0]; a.bbox[4 * f + 1] = s_bb[1];
      a.bbox[4 * f + 2] = s_bb[2]; a.bbox[4 * f + 3] = s_bb[3];
    }
    if (n > 0 || a.list_all) a.worklist[atomicAdd(a.work_n, 1)] = f;
  }
}

// ---- contact frames ----------------------------------------------------------------------
struct ContactArgs {
  const float* M;            // (F,12)
  const Setup* setups;
  const NRec* nrecs;         // (F, kmax) camera-frame vertex normals per setup slot
  const int32_t* counts;
  const int32_t* bbox;
  const int32_t* worklist;
  const int32_t* work_n;
  int32_t* cursor;           // (1) work-stealing cursor, zeroed by the launcher
  const float* verts;
  const float* vnorm;
  const int32_t* faces;
  const float* depth0;       // (TH,TW)
  const float* hiz;          // depth0 max-pyramid
  const float* dxp;          // (TW)
  const float* dyp;          // (TH)
  const uint8_t* bg_sim;     // (TH,TW,3)
  const uint8_t* bg_real;
  const int32_t* bg_id;
  uint8_t* color;            // (F,TH,TW,3)  required
  float* gel_depth;          // (F,TH,TW)    required
  float* obs;
  int64_t obs_env_stride, obs_sensor_stride;
  int sensors_per_env;
  int kmax;
  int budget;                // region pixel budget (<= the compiled one)
  FillArgs fill;             // fill.parts != 0: this kernel writes those parts of every listed frame first
};
#ifndef CT_BLOCK_N
#define CT_BLOCK_N 512
#endif
#ifndef CT_CTAS
#define CT_CTAS 2
#endif
#ifndef CT_BUD_N
#define CT_BUD_N 12288
#endif
#ifndef CT_ZERO_BYTES
#define CT_ZERO_BYTES 4096
#endif
#ifndef CT_EARLY_OUT
#define CT_EARLY_OUT 1   // 1: the two exact early-outs for clipped fragments (light loop, zero-difference regions)
#endif
constexpr int CT_BLOCK = CT_BLOCK_N;
constexpr int CT_CHUNK = 1024;       // triangles whose scan rows are enumerated together
static_assert(CT_CHUNK % CT_BLOCK_N == 0, "CT_CHUNK must be a multiple of the block size");
constexpr int CT_BUD_GRAY = CT_BUD_N;    // region pixels (interior + halo) held in shared memory, 8 B each
constexpr int CT_BUD_RGB = 4096;     // three-channel path: + 24 B per pixel of difference / blur planes
// Empty z-buffer entry: depth +inf in the high word, 0 in the low word.  Any fragment key (finite depth, low word
// >= 1) is smaller; a pixel is a hit iff its low word is non-zero; and once shading reuses the two words of a pixel
// as (difference, horizontal blur), an untouched pixel already READS as difference 0.0f - no clearing pass.
constexpr unsigned long long CT_EMPTY = 0x7f80000000000000ull;

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ Setup load_setup(const Setup* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  union { uint4 v[4]; Setup s; } u;
  u.v[0] = __ldg(q); u.v[1] = __ldg(q + 1); u.v[2] = __ldg(q + 2); u.v[3] = __ldg(q + 3);
  return u.s;
}

// Pixel-column interval [xlo, xhi] of row `dy` that can lie inside the three edge planes, and whether it
// is exact.  An edge value is affine in the ray slope dx: e = dx*n.x + (dy*n.y - n.z).  Each edge value
// is also a monotone function of the column even with its roundings (products and sums round
// monotonically), so the covered columns of a row form ONE interval whose ends are the columns where
// the edges switch.  A switching point is known up to +-m (worst-case rounding of the exactly-rounded
// per-pixel evaluation in cover()); when no pixel centre of the row lies inside an uncertainty band that
// matters, the interval is exact and its pixels need no edge test at all, only depth.  Otherwise the
// conservative interval is returned with exact = false and every pixel runs cover().
__device__ __forceinline__ void row_span(const Setup& s, float dy, float sx0, float kx, int bx0, int bx1, int& xlo,
                                         int& xhi, bool& exact) {
  float Llo = -1e30f, Lhi = -1e30f, Ulo = 1e30f, Uhi = 1e30f;
  bool empty = false, sure = true;
  auto edge = [&](V3 n) {
    const float cy = dy * n.y;
    const float c = cy - n.z;
    const float eps = 2e-6f * (1.7f * fabsf(n.x) + fabsf(cy) + fabsf(n.z));
    if (n.x == 0.0f) {
      empty = empty || (c > eps);
      sure = sure && (fabsf(c) > eps);
    } else {
      const float inv = __fdividef(1.0f, n.x);
      const float b = -c * inv, m = eps * fabsf(inv) + 1e-7f * fabsf(b);
      if (n.x > 0.0f) { Uhi = fminf(Uhi, b + m); Ulo = fminf(Ulo, b - m); }
      else { Llo = fmaxf(Llo, b - m); Lhi = fmaxf(Lhi, b + m); }
    }
  };
  edge(s.n0); edge(s.n1); edge(s.n2);
  const float lim_lo = -1.0f, lim_hi = (float)TW;
  const int loA = (int)ceilf(fminf(fmaxf((Llo - sx0) * kx - 0.002f, lim_lo), lim_hi));
  const int loB = (int)ceilf(fminf(fmaxf((Lhi - sx0) * kx + 0.002f, lim_lo), lim_hi));
  const int hiA = (int)floorf(fminf(fmaxf((Uhi - sx0) * kx + 0.002f, lim_lo), lim_hi));
  const int hiB = (int)floorf(fminf(fmaxf((Ulo - sx0) * kx - 0.002f, lim_lo), lim_hi));
  xlo = max(bx0, loA);
  xhi = empty ? -1 : min(bx1, hiA);
  exact = sure && loB <= xlo && hiB >= xhi;
}

// Bounds of the pixels this lane has put a peg fragment on (image coordinates).
struct HitBox {
  int x0, y0, x1, y1;
  __device__ __forceinline__ void add(int px, int py) {
    x0 = min(x0, px); x1 = max(x1, px); y0 = min(y0, py); y1 = max(y1, py);
  }
};

// The z test of one fragment of depth t at the shared z-buffer entry *zp: GL_LESS against the gel (drawn first,
// d0 = its depth at this pixel, 0 = no gel: ties keep the gel), then against the peg fragments already there.
// key = depth bits << 32 | (orig face << 12 | slot) + 1: equal depths resolve to the lower original face index.
__device__ __forceinline__ void z_test(float t, float d0, uint32_t orig, int k, unsigned long long* zp, HitBox& hb, int px,
                                       int py) {
  if (d0 != 0.0f && !(t < d0)) return;
  const unsigned long long key =
      ((unsigned long long)__float_as_uint(t) << 32) | (unsigned long long)(((orig << 12) | (uint32_t)k) + 1u);
  if (key < *zp) {
    atomicMin(zp, key);
    hb.add(px, py);      // a pixel that ever received a peg fragment in front of the gel stays a hit
  }
}

// One fragment: exact coverage + depth.
__device__ __forceinline__ void raster_frag(float znear, const Setup& s, int k, float dx, float dy, unsigned long long* zp,
                                            const float* __restrict__ d0p, HitBox& hb, int px, int py) {
  const float d0 = __ldg(d0p);   // issued first: its latency hides behind the edge functions and the division
  float e1, e2, es;
  const float t = cover(znear, s, dx, dy, e1, e2, es);
  if (t < 0.0f) return;
  z_test(t, d0, s.orig, k, zp, hb, px, py);
}

// Fragment of an exact span: inside by construction, only depth (N, det) and the z test.
__device__ __forceinline__ void raster_frag_depth(float znear, V3 N, float det, uint32_t orig, int k, float dx, float dy,
                                                  unsigned long long* zp, const float* __restrict__ d0p, HitBox& hb,
                                                  int px, int py) {
  const float d0 = __ldg(d0p);
  const float t = __fdiv_rn(det, edge_fn(dx, dy, N));
  if (!(t >= znear)) return;
  z_test(t, d0, orig, k, zp, hb, px, py);
}

#ifdef CT_MAXNREG   // experiment hook: cap the registers instead of asking for CT_CTAS resident CTAs
template <int NCH>
__global__ void __maxnreg__(CT_MAXNREG) tac_contact(const __grid_constant__ TacConst kc, ContactArgs a) {
#else
template <int NCH>
__global__ void __launch_bounds__(CT_BLOCK, NCH == 1 ? CT_CTAS : 1) tac_contact(const __grid_constant__ TacConst kc, ContactArgs a) {
#endif
#ifdef CT_PROFILE
  long long t_prof = clock64();
#endif
  constexpr int BUD = NCH == 1 ? CT_BUD_GRAY : CT_BUD_RGB;
  constexpr int DS = NCH == 1 ? 2 : 3;  // float stride between pixels of the difference / blur planes
  extern __shared__ __align__(16) unsigned char ct_smem[];
  // z-buffer (8 B / pixel).  Gray path: once a pixel is shaded its key is dead, and the two words
  // are reused as (difference, horizontal blur) so the whole pipeline lives in 8 B per pixel.
  unsigned long long* s_z = reinterpret_cast<unsigned long long*>(ct_smem);
  float* s_diff = NCH == 1 ? reinterpret_cast<float*>(ct_smem) : reinterpret_cast<float*>(ct_smem + (size_t)BUD * 8);
  float* s_h = NCH == 1 ? s_diff + 1 : s_diff + (size_t)BUD * 3;
  __shared__ int s_off[CT_CHUNK];
  __shared__ int s_wsum[CT_BLOCK / 32];
  __shared__ int s_anydiff;
  // Frame pipeline.  Frame ids come from the worklist TWO frames ahead, a frame's header (triangle count, dirty box,
  // background id - written by tac_geom and long evicted from L2 by the 4 GB of fill traffic in between) ONE frame
  // ahead: loaded into registers when a frame starts, parked in shared memory after its raster phase, and the next
  // frame's setup / normal records are prefetched into L2 behind that.  No DRAM round trip stays on the critical
  // path between two frames.
  __shared__ int s_ring[3];     // frame ids: [it % 3] current, [(it + 1) % 3] next, [(it + 2) % 3] the one after
  __shared__ int s_hdr[2][8];   // header of the frame of parity it & 1: K, box x0 y0 x1 y1, background id
  __shared__ int s_rctr, s_sctr;   // next raster item / next hit-box pixel to hand out
  __shared__ int s_hb[4];  // bounds of the frame's hit pixels (all sub-windows whose colour changed)
  __shared__ int s_sb[4];  // bounds of the hit pixels of the current sub-window's region
  __shared__ unsigned short s_q[CT_BLOCK / 32][64];  // per-warp queue of hit pixels waiting to be shaded
  __shared__ double s_rb[511];            // remove_bg: d / 255.0 + 0.5 for d = -255..255 (f64 divide once)
  __shared__ float s_dxp[TW], s_dyp[TH];  // ray-slope tables (per-lane indexing would serialise in the constant cache)
  __shared__ __align__(16) float s_zero[CT_ZERO_BYTES / 4];  // source of the asynchronous gel_depth = 0 fill
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CT_BLOCK / 32;
  for (int i = tid; i < TW; i += CT_BLOCK) { s_dxp[i] = __ldg(a.dxp + i); s_dyp[i] = __ldg(a.dyp + i); }
  for (int i = tid; i < 511; i += CT_BLOCK) s_rb[i] = (double)(i - 255) / 255.0 + 0.5;
  for (int i = tid; i < CT_ZERO_BYTES / 4; i += CT_BLOCK) s_zero[i] = 0.0f;
  igi_fence_proxy_async();   // the zeros must be visible to the bulk-copy (async) proxy
  const float span_x0 = kc.sx0, span_kx = (float)(TW - 1) / (kc.sx1 - kc.sx0);
  // gel_depth = 0 of one frame (200 704 B of zeros): a few lanes of every warp hand 4 KB pieces of the zero
  // buffer to the bulk-copy engine (shared -> global), which costs this issue-bound kernel no store
  // instructions.  Every thread commits one (possibly empty) bulk group per call.
  auto zero_fill_async = [&](int frame) {
    if ((a.fill.parts & 2) && a.fill.gel_depth && frame >= 0) {
      constexpr int NB = TW * TH * 4 / CT_ZERO_BYTES;
      static_assert(NB * CT_ZERO_BYTES == TW * TH * 4, "zero buffer must divide the gel_depth frame");
      constexpr int PER_WARP = (NB + NW - 1) / NW;
      const int i = warp * PER_WARP + lane;
      if (lane < PER_WARP && i < NB) {
        char* dst = reinterpret_cast<char*>(a.fill.gel_depth + (size_t)frame * TW * TH);
        igi_bulk_s2g(dst + (size_t)i * CT_ZERO_BYTES, s_zero, CT_ZERO_BYTES);
      }
    }
    igi_bulk_commit();
  };
  constexpr int HDR_N = 6;
  auto fetch_id = [&]() {
    const int w = atomicAdd(a.cursor, 1);
    return (w < *a.work_n) ? a.worklist[w] : -1;
  };
  auto load_hdr = [&](int frame) -> int {   // thread tid < HDR_N loads word tid of the frame's header
    if (frame < 0 || tid >= HDR_N) return 0;
    if (tid == 0) return a.counts[frame];
    if (tid < 5) return a.bbox[4 * frame + tid - 1];
    return a.bg_id[frame];
  };
  if (tid == 0) {
    s_ring[0] = fetch_id();
    s_ring[1] = s_ring[0] >= 0 ? fetch_id() : -1;
  }
  __syncthreads();
  {
    const int v = load_hdr(s_ring[0]);
    if (tid < HDR_N) s_hdr[0][tid] = v;
  }
  // The zero fill of a frame's gel_depth is issued when the frame BEFORE it starts, so it has a whole frame time to
  // land before that frame's first gel_depth write.
  zero_fill_async(s_ring[0]);
  for (int it = 0;; ++it) {
    __syncthreads();   // the previous frame is finished (its obs pixels read shared memory); ring and header slots are published
    const int f = s_ring[it % 3];
    CT_T(0);
    if (f < 0) {
      igi_bulk_wait0();   // s_zero must outlive the copies that read it
      return;
    }
    zero_fill_async(s_ring[(it + 1) % 3]);
    const int* hdr = s_hdr[it & 1];
    const int K = hdr[0];
    // Once per frame, at a point where a late warp 0 costs nothing (the raster batches are handed out dynamically):
    // take the id of the frame after next from the worklist and bring the next frame's header into shared memory.
    auto advance = [&]() {
      if (tid < HDR_N) {
        const int nf = s_ring[(it + 1) % 3];
        const int v = load_hdr(nf);
        if (tid == 0) s_ring[(it + 2) % 3] = nf >= 0 ? fetch_id() : -1;
        s_hdr[(it + 1) & 1][tid] = v;
      }
    };
    if (tid == 0) { s_hb[0] = TW; s_hb[1] = TH; s_hb[2] = -1; s_hb[3] = -1; }
    bool frame_changed = false;
    // Fused fill: the other parts of the frame's no-contact result that tac_geom left to this kernel go out
    // with plain stores while the raster / shading work of the frame runs; the barriers below order them
    // before the rewrite of the changed box.
    if (a.fill.parts & ~2) {
      FillArgs fa = a.fill;
      fa.parts &= ~2;
      fill_frame(fa, f, tid, CT_BLOCK);
    }
    if (K <= 0) {   // listed only to be filled
      advance();
      continue;
    }
    const Setup* list = a.setups + (size_t)f * a.kmax;
    const NRec* nlist = a.nrecs + (size_t)f * a.kmax;
    // window that can change: union of the triangle boxes, dilated by the blur radius
    const int wx0 = max(hdr[1] - HALO, 0), wy0 = max(hdr[2] - HALO, 0);
    const int wx1 = min(hdr[3] + HALO, TW - 1), wy1 = min(hdr[4] + HALO, TH - 1);
    const uint8_t* bgr = a.bg_real + (size_t)hdr[5] * TW * TH * 3;
    uint8_t* col = a.color + (size_t)f * TW * TH * 3;
    float* gdep = a.gel_depth + (size_t)f * TW * TH;
    // cut the window into the fewest sub-windows whose region (interior + halo) fits the budget
    const int ww = wx1 - wx0 + 1, wh = wy1 - wy0 + 1;
    int nsx = 1, nsy = 1;
    while (((ww + nsx - 1) / nsx + 2 * HALO) * ((wh + nsy - 1) / nsy + 2 * HALO) > a.budget) {
      if ((ww + nsx - 1) / nsx >= (wh + nsy - 1) / nsy) ++nsx; else ++nsy;
    }
    const int tw = (ww + nsx - 1) / nsx, th = (wh + nsy - 1) / nsy;
    const int RW = tw + 2 * HALO, RH = th + 2 * HALO;

    for (int ty = wy0; ty <= wy1; ty += th)
      for (int tx = wx0; tx <= wx1; tx += tw) {
        // interior [tx, ix1] x [ty, iy1]; region = interior + halo at image coordinates rx0.., ry0..
        const int ix1 = min(tx + tw - 1, wx1), iy1 = min(ty + th - 1, wy1);
        const int rx0 = tx - HALO, ry0 = ty - HALO;
        const int cx0 = max(rx0, 0), cy0 = max(ry0, 0);
        const int cx1 = min(ix1 + HALO, TW - 1), cy1 = min(iy1 + HALO, TH - 1);
        if (tx != wx0 || ty != wy0) __syncthreads();   // (the frame's first region starts behind the barrier at the top of the frame loop)
        CT_T(1);
        // --- raster: work items are (triangle, image row) pairs, enumerated with a block scan so that
        // every thread gets the same number of rows whatever the triangle sizes are
        HitBox hb{TW, TH, -1, -1};
        for (int c0 = 0; c0 < K; c0 += CT_CHUNK) {
          const int kn = min(CT_CHUNK, K - c0);
          int rows[CT_CHUNK / CT_BLOCK], sum = 0;
#pragma unroll
          for (int j = 0; j < CT_CHUNK / CT_BLOCK; ++j) {
            const int k = tid * (CT_CHUNK / CT_BLOCK) + j;
            rows[j] = 0;
            if (k < kn) {
              const uint32_t bb = __ldg(&list[c0 + k].bbox);
              const int bx0 = max((int)(bb & 255u), cx0), by0 = max((int)((bb >> 8) & 255u), cy0);
              const int bx1 = min((int)((bb >> 16) & 255u), cx1), by1 = min((int)(bb >> 24), cy1);
              if (bx0 <= bx1 && by0 <= by1) rows[j] = by1 - by0 + 1;
            }
            sum += rows[j];
          }
          if (c0 == 0) {
            // --- z-buffer starts EMPTY (two pixels per 128-bit store, no index arithmetic, no global loads; the gel's
            // depth is looked up per fragment instead).  Placed here so that the stores run under the latency of the
            // box loads above.
            uint4* z4 = reinterpret_cast<uint4*>(s_z);
            const int n4 = (RW * RH + 1) >> 1;
            const uint32_t hi = (uint32_t)(CT_EMPTY >> 32);
            for (int i = tid; i < n4; i += CT_BLOCK) z4[i] = make_uint4(0u, hi, 0u, hi);
            if (NCH != 1) {   // three-channel path: the difference plane is separate from the keys
              for (int i = tid; i < RW * RH * 3; i += CT_BLOCK) s_diff[i] = 0.0f;
            }
            if (tid == 0) { s_anydiff = 0; s_sctr = 0; s_sb[0] = TW; s_sb[1] = TH; s_sb[2] = -1; s_sb[3] = -1; }
          }
          int incl = sum;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
          }
          if (c0 > 0) __syncthreads();  // the previous chunk's readers of s_off / s_wsum are done (chunk 0 starts behind a barrier anyway)
          if (lane == 31) s_wsum[warp] = incl;
          __syncthreads();
          int woff = 0, total = 0;
#pragma unroll
          for (int w = 0; w < NW; ++w) {
            const int t = s_wsum[w];
            if (w < warp) woff += t;
            total += t;
          }
          int base = woff + incl - sum;
#pragma unroll
          for (int j = 0; j < CT_CHUNK / CT_BLOCK; ++j) {
            s_off[tid * (CT_CHUNK / CT_BLOCK) + j] = base;
            base += rows[j];
          }
          if (tid == 0) s_rctr = 0;
          __syncthreads();
          CT_T(2);
          if (c0 == 0 && tx == wx0 && ty == wy0) advance();   // first chunk of the frame's first region
          if (tid == 0) { CT_COUNT(9, total); CT_COUNT(12, 1); CT_COUNT(13, RW * RH); CT_COUNT(14, kn); }
          for (;;) {
            int i0 = 0;
            if (lane == 0) i0 = atomicAdd(&s_rctr, 32);
            i0 = __shfl_sync(0xffffffffu, i0, 0);
            if (i0 >= total) break;
            const int i = i0 + lane;
            int k = 0, py = 0, xlo = 0, xhi = -1, exact = 1;
            Setup s;
            if (i < total) {
              int lo = 0, hi = kn - 1;
              while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
              }
              k = c0 + lo;
              s = load_setup(list + k);
              const int bx0 = max((int)(s.bbox & 255u), cx0), by0 = max((int)((s.bbox >> 8) & 255u), cy0);
              const int bx1 = min((int)((s.bbox >> 16) & 255u), cx1);
              py = by0 + (i - s_off[lo]);
              const float dy = s_dyp[py];
              bool ex;
              row_span(s, dy, span_x0, span_kx, bx0, bx1, xlo, xhi, ex);
              exact = ex ? 1 : 0;
              // most spans are one or two pixels: those are done right here by the lane that owns the row
              unsigned long long* zrow = s_z + (py - ry0) * RW - rx0;
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int px = xlo + u;
                if (px <= xhi) {
                  if (ex) raster_frag_depth(kc.znear, s.N, s.det, s.orig, k, s_dxp[px], dy, zrow + px, a.depth0 + py * TW + px, hb, px, py);
                  else raster_frag(kc.znear, s, k, s_dxp[px], dy, zrow + px, a.depth0 + py * TW + px, hb, px, py);
                }
              }
              xlo += 2;
            }
            // What is left of the longer spans is flattened into one pixel list (warp scan of the
            // lengths) and dealt out one pixel per lane, so lanes stay busy whatever the span lengths.
            const int len = max(xhi - xlo + 1, 0);
            int excl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, excl, d);
              if (lane >= d) excl += t;
            }
            const int npix = __shfl_sync(0xffffffffu, excl, 31);
            excl -= len;
            if (lane == 0) CT_COUNT(10, npix);
            for (int p0 = 0; p0 < npix; p0 += 32) {
              const int p = p0 + lane;
              int src = 0;  // last lane whose span starts at or before pixel p
#pragma unroll
              for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(0xffffffffu, excl, src + step);   // src + step <= 31
                if (v <= p) src += step;
              }
              const int kk = __shfl_sync(0xffffffffu, k, src);
              const int yx = __shfl_sync(0xffffffffu, py | (exact << 8) | (xlo << 16), src);
              const int yy = yx & 255, px = (yx >> 16) + (p - __shfl_sync(0xffffffffu, excl, src));
              if (p < npix) {
                unsigned long long* zp = s_z + (yy - ry0) * RW + (px - rx0);
                const float dx = s_dxp[px], dy = s_dyp[yy];
                if (yx & 256) {
                  // exact span: depth only (second half of the setup record: N, det, orig)
                  const uint4* q4 = reinterpret_cast<const uint4*>(list + kk);
                  const uint4 u2 = __ldg(q4 + 2), u3 = __ldg(q4 + 3);
                  const V3 N{__uint_as_float(u2.y), __uint_as_float(u2.z), __uint_as_float(u2.w)};
                  raster_frag_depth(kc.znear, N, __uint_as_float(u3.x), u3.w, kk, dx, dy, zp, a.depth0 + yy * TW + px, hb, px, yy);
                } else {
                  const Setup ss = load_setup(list + kk);
                  raster_frag(kc.znear, ss, kk, dx, dy, zp, a.depth0 + yy * TW + px, hb, px, yy);
                }
              }
            }
          }
        }
        // bounds of the hit pixels of this region (one shared-memory atomic per warp and bound)
        if (__any_sync(0xffffffffu, hb.x1 >= 0)) {
          const int x0 = __reduce_min_sync(0xffffffffu, hb.x0), x1 = __reduce_max_sync(0xffffffffu, hb.x1);
          const int y0 = __reduce_min_sync(0xffffffffu, hb.y0), y1 = __reduce_max_sync(0xffffffffu, hb.y1);
          if (lane == 0) {
            atomicMin(&s_sb[0], x0); atomicMax(&s_sb[2], x1);
            atomicMin(&s_sb[1], y0); atomicMax(&s_sb[3], y1);
          }
        }
        igi_bulk_wait1();   // this frame's zero fill has landed; only the next frame's group may still be in flight
        __syncthreads();
        CT_T(3);
        if (tx == wx0 && ty == wy0) {
          // the next frame's triangle records (written by tac_geom, in DRAM by now) -> L2, one 128-byte line per thread
          const int Kn = s_hdr[(it + 1) & 1][0], fn = s_ring[(it + 1) % 3];
          if (fn >= 0 && Kn > 0) {
            const char* p0 = reinterpret_cast<const char*>(a.setups + (size_t)fn * a.kmax);
            const char* p1 = reinterpret_cast<const char*>(a.nrecs + (size_t)fn * a.kmax);
            for (int o = tid * 128; o < Kn * (int)sizeof(Setup); o += CT_BLOCK * 128) igi_prefetch_l2(p0 + o);
            for (int o = tid * 128; o < Kn * (int)sizeof(NRec); o += CT_BLOCK * 128) igi_prefetch_l2(p1 + o);
          }
        }
        if (s_sb[2] < 0) continue;  // nothing of the peg is visible here: fill already wrote the result
        // --- shade hits, build the scaled difference image (0 where the gel is visible).  Only the box of the hit
        // pixels is scanned; each warp takes 32 box pixels at a time, queues the hit ones and shades a full warp of
        // them whenever 32 are waiting, so the long shading path runs with all lanes active.
        const int hbx0 = s_sb[0], hby0 = s_sb[1], hbx1 = s_sb[2], hby1 = s_sb[3];
        {
          const int bw = hbx1 - hbx0 + 1, npx = bw * (hby1 - hby0 + 1);
          const uint32_t inv_bw = 0xffffffffu / (uint32_t)bw + 1u;   // wraps to 0 for bw == 1
          const uint32_t inv_rw = 0xffffffffu / (uint32_t)RW + 1u;   // exact floor(i / RW) for i < 2^16
          unsigned short* q = s_q[warp];
          int qn = 0;
          bool nz_any = false;
          // all 32 lanes run the arithmetic (warp-uniform early-out votes inside shade_t); `act` gates the writes
          auto shade_px = [&](int i, bool act) {
            // lanes without a pixel (the last, partial batch of a warp) run the shader on a dummy fragment: they
            // touch no shared memory, so no lane reads a key that another lane of the warp is about to overwrite
            V3 pp{0.f, 0.f, -1.f}, n{0.f, 0.f, 1.f};
            float t = 1.0f;
            int px = 0, py = 0;
            if (act) {
              const int ry = (int)__umulhi((uint32_t)i, inv_rw), rx = i - ry * RW;
              px = rx0 + rx; py = ry0 + ry;
              const unsigned long long key = s_z[i];
              const uint32_t low = (uint32_t)key;
              t = __uint_as_float((uint32_t)(key >> 32));
              const int slot = (int)((low - 1u) & 0xfffu);
              const Setup s = load_setup(list + slot);
              const uint4* nq = reinterpret_cast<const uint4*>(nlist + slot);
              const uint4 na = __ldg(nq), nb = __ldg(nq + 1), nc = __ldg(nq + 2);
              const float dx = s_dxp[px], dy = s_dyp[py];
              const float e0 = edge_fn(dx, dy, s.n0), e1 = edge_fn(dx, dy, s.n1), e2 = edge_fn(dx, dy, s.n2);
              const float es = add(add(e0, e1), e2);
              const float ies = __fdividef(1.0f, es);
              const float l1 = e1 * ies, l2 = e2 * ies;
              const float l0 = 1.0f - l1 - l2;
              // barycentric blend of the camera-frame vertex normals (rotation and blend commute)
              n.x = l0 * __uint_as_float(na.x) + l1 * __uint_as_float(na.w) + l2 * __uint_as_float(nb.z);
              n.y = l0 * __uint_as_float(na.y) + l1 * __uint_as_float(nb.x) + l2 * __uint_as_float(nb.w);
              n.z = l0 * __uint_as_float(na.z) + l1 * __uint_as_float(nb.y) + l2 * __uint_as_float(nc.x);
              const float r = rsqrtf(fmaxf(n.x * n.x + n.y * n.y + n.z * n.z, 1e-30f));
              n.x *= r; n.y *= r; n.z *= r;
              pp = V3{mul(dx, t), mul(dy, t), -t};
            }
            uint8_t rgb[3];
            shade_t<NCH, 0, CT_EARLY_OUT != 0>(kc, pp, n, rgb, act);
            if (!act) return;
            const uint8_t* bs = a.bg_sim + (py * TW + px) * 3;
            // gel_depth = depth0 - depth (allsight_render.py:193-197), interior pixels only
            if (px >= tx && px <= ix1 && py >= ty && py <= iy1)
              gdep[py * TW + px] = sub(__ldg(a.depth0 + py * TW + px), t);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              const int d = (int)rgb[c] - (int)bs[c];
              nz_any = nz_any || d != 0;
              s_diff[DS * i + c] = (float)d * kc.calib_scale;
            }
            CT_COUNT(11, 1);
          };
          for (;;) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_sctr, 32);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= npx) break;
            const int j = base + lane;
            bool hit = false;
            int i = 0;
            if (j < npx) {
              const int by = bw == 1 ? j : (int)__umulhi((uint32_t)j, inv_bw), bx = j - by * bw;
              i = (hby0 + by - ry0) * RW + (hbx0 + bx - rx0);
              hit = (uint32_t)s_z[i] != 0u;     // low word: 0 = no peg fragment here (reads as difference 0)
            }
            const unsigned bm = __ballot_sync(0xffffffffu, hit);
            if (hit) q[qn + __popc(bm & ((1u << lane) - 1u))] = (unsigned short)i;
            qn += __popc(bm);
            __syncwarp();
            if (qn >= 32) {
              qn -= 32;
              const int idx = q[qn + lane];
              __syncwarp();
              shade_px(idx, true);
            }
          }
          if (qn > 0) shade_px(lane < qn ? q[lane] : 0, lane < qn);
          if (__any_sync(0xffffffffu, nz_any) && lane == 0) s_anydiff = 1;
        }
        __syncthreads();
        CT_T(4);
#if CT_EARLY_OUT
        // Every shaded pixel equals the simulated background (e.g. both clip at 255 under the inverse-square light
        // model): the difference image is identically 0, so the calibrated colour is the real background frame the
        // fill already wrote and the observation is the empty one.  Exact, not an approximation.
        if (!s_anydiff) continue;
#endif
        frame_changed = true;   // uniform: every thread read the same flag behind the same barrier
        // Only pixels within the blur radius of a hit can differ from what tac_fill wrote (the difference
        // image is 0 elsewhere): the changed box = hit box dilated by HALO, inside the interior.
        const int bx0 = max(hbx0 - HALO, tx), bx1 = min(hbx1 + HALO, ix1);
        const int by0 = max(hby0 - HALO, ty), by1 = min(hby1 + HALO, iy1);
        if (tid == 0) {
          s_hb[0] = min(s_hb[0], hbx0); s_hb[1] = min(s_hb[1], hby0);
          s_hb[2] = max(s_hb[2], hbx1); s_hb[3] = max(s_hb[3], hby1);
        }
        if (bx0 > bx1 || by0 > by1) continue;   // hits only in this region's halo: they belong to a neighbour
        // --- 7-tap horizontal pass over the changed columns (BORDER_REFLECT_101 at the image edge), for the
        // rows the vertical pass will read.  A thread produces 4 neighbouring outputs from one 10-value window.
        const int iw = bx1 - bx0 + 1, ih = by1 - by0 + 1;
        {
          const int nqx = (iw + 3) >> 2;
          const uint32_t inv_q = 0xffffffffu / (uint32_t)nqx + 1u;   // wraps to 0 for nqx == 1
          const int hr0 = by0 - HALO - ry0, nhr = ih + 2 * HALO;     // region rows [hr0, hr0 + nhr)
          for (int it = tid; it < nhr * nqx; it += CT_BLOCK) {
            const int rq = nqx == 1 ? it : (int)__umulhi((uint32_t)it, inv_q), qx = it - rq * nqx;
            const int ry = hr0 + rq;
            const int py = ry0 + ry;
            if (py < 0 || py >= TH) continue;
            const int px0 = bx0 + 4 * qx;
            float win[10][NCH];
            const bool inner = px0 >= HALO && px0 + 3 + HALO < TW;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const int sx = (inner ? px0 + j - HALO : reflect101(px0 + j - HALO, TW)) - rx0;
              const float* dp = &s_diff[(ry * RW + min(sx, RW - 1)) * DS];
#pragma unroll
              for (int c = 0; c < NCH; ++c) win[j][c] = dp[c];
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (px0 + o > bx1) break;
              float acc[NCH];
#pragma unroll
              for (int c = 0; c < NCH; ++c) acc[c] = 0.f;
#pragma unroll
              for (int k = 0; k < 7; ++k) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) acc[c] = fmaf(kc.gauss[k], win[o + k][c], acc[c]);
              }
              float* hp = &s_h[(ry * RW + (px0 + o - rx0)) * DS];
#pragma unroll
              for (int c = 0; c < NCH; ++c) hp[c] = acc[c];
            }
          }
        }
        __syncthreads();
        CT_T(5);
        // --- vertical pass, + bg_real, clip, truncate (numpy astype(uint8)); 4 rows of a column per
        // thread.  The per-channel change (colour - bg_real, 10 bits each) is left in the pixel's dead
        // low word for the obs stage.
        {
          const int nqy = (ih + 3) >> 2;
          const uint32_t inv_w = 0xffffffffu / (uint32_t)iw + 1u;    // wraps to 0 for iw == 1
          for (int it = tid; it < nqy * iw; it += CT_BLOCK) {
            const int qy = iw == 1 ? it : (int)__umulhi((uint32_t)it, inv_w), lx = it - qy * iw;
            const int px = bx0 + lx, py0 = by0 + 4 * qy;
            float win[10][NCH];
            const bool inner = py0 >= HALO && py0 + 3 + HALO < TH;
#pragma unroll
            for (int j = 0; j < 10; ++j) {
              const int sy = (inner ? py0 + j - HALO : reflect101(py0 + j - HALO, TH)) - ry0;
              const float* hp = &s_h[(min(sy, RH - 1) * RW + (px - rx0)) * DS];
#pragma unroll
              for (int c = 0; c < NCH; ++c) win[j][c] = hp[c];
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const int py = py0 + o;
              if (py > by1) break;
              float acc[NCH];
#pragma unroll
              for (int c = 0; c < NCH; ++c) acc[c] = 0.f;
#pragma unroll
              for (int k = 0; k < 7; ++k) {
#pragma unroll
                for (int c = 0; c < NCH; ++c) acc[c] = fmaf(kc.gauss[k], win[o + k][c], acc[c]);
              }
              const size_t off = ((size_t)py * TW + px) * 3;
              uint32_t packed = 0;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const int bb = (int)__ldg(bgr + off + c);
                const float v = clampf(acc[NCH == 1 ? 0 : c] + (float)bb, kc.clip_lo, kc.clip_hi);
                const int q = (int)(uint8_t)v;
                col[off + c] = (uint8_t)q;
                packed |= (uint32_t)(q - bb + 256) << (10 * c);
              }
              if (NCH == 1) reinterpret_cast<uint32_t*>(s_z)[2 * ((py - ry0) * RW + (px - rx0))] = packed;
            }
          }
        }
        CT_T(6);
      }
    if (!frame_changed) continue;   // no region changed a colour: the obs row stays the empty one the fill wrote
    __syncthreads();
    CT_T(7);
    // --- obs pixels whose 3.5x3.5 source window meets a pixel that changed ---------------------
    // flipud + crop: obs row r reads flipped rows [3.5r, 3.5r+3.5) = original rows 223 - that.
    if (s_hb[2] >= 0) {
      const int hx0 = max(s_hb[0] - HALO, 0), hy0 = max(s_hb[1] - HALO, 0);
      const int hx1 = min(s_hb[2] + HALO, TW - 1), hy1 = min(s_hb[3] + HALO, TH - 1);
      const int fy_lo = TH - 1 - hy1, fy_hi = TH - 1 - hy0;  // changed rows in flipped coordinates
      const int oy0 = max((2 * fy_lo) / 7 - 1, 0), oy1 = min((2 * fy_hi) / 7 + 1, OBS_H - 1);
      const int ox0 = max((2 * hx0) / 7 - 1, 0), ox1 = min((2 * hx1) / 7 + 1, OBS_W - 1);
      const int nw = ox1 - ox0 + 1, nh = oy1 - oy0 + 1;
      if (tid == 0) { CT_COUNT(15, 1); CT_COUNT(23, (hx1 - hx0 + 1) * (hy1 - hy0 + 1)); CT_COUNT(24, ww * wh); }
      float* ob = a.obs + (size_t)(f / a.sensors_per_env) * a.obs_env_stride +
                  (size_t)(f % a.sensors_per_env) * a.obs_sensor_stride;
      // One region held the whole window: every changed pixel's (colour - bg_real) sits in shared
      // memory and everything outside the window is unchanged (difference 0), so no global reads.
      const bool from_smem = NCH == 1 && nsx == 1 && nsy == 1;
      const uint32_t* s_delta = reinterpret_cast<const uint32_t*>(s_z);
      const int rx0 = wx0 - HALO, ry0 = wy0 - HALO;
      // changed box of the (single) region: the vertical pass left colour - bg_real there, it is 0 elsewhere
      const int cbx0 = max(hx0, wx0), cbx1 = min(hx1, wx1), cby0 = max(hy0, wy0), cby1 = min(hy1, wy1);
      for (int i = tid; i < nw * nh; i += CT_BLOCK) {
        const int oy = oy0 + i / nw, ox = ox0 + i % nw;
        // cv2 INTER_AREA, scale 3.5: taps for even/odd destination index
        const int sx0 = (ox >> 1) * 7 + ((ox & 1) ? 3 : 0), sy0 = (oy >> 1) * 7 + ((oy & 1) ? 3 : 0);
        double acc[3] = {0.0, 0.0, 0.0};
        for (int j = 0; j < 4; ++j) {
          const float wy = (oy & 1) ? (j == 0 ? AREA_W1 : AREA_W0) : (j == 3 ? AREA_W1 : AREA_W0);
          const int fy = sy0 + j;           // flipped row
          const int py = TH - 1 - fy;       // original row
          double racc[3] = {0.0, 0.0, 0.0};
          for (int k = 0; k < 4; ++k) {
            const float wx = (ox & 1) ? (k == 0 ? AREA_W1 : AREA_W0) : (k == 3 ? AREA_W1 : AREA_W0);
            const int px = sx0 + k;
            const int ddx = px - TW / 2, ddy = py - TH / 2;
            const double m = (ddx * ddx + ddy * ddy <= (TW / 2) * (TW / 2)) ? 1.0 : 0.0;  // circle_mask
            int dl[3];
            if (from_smem) {
              uint32_t pk = (256u << 20) | (256u << 10) | 256u;
              if (px >= cbx0 && px <= cbx1 && py >= cby0 && py <= cby1) pk = s_delta[2 * ((py - ry0) * RW + (px - rx0))];
#pragma unroll
              for (int c = 0; c < 3; ++c) dl[c] = (int)((pk >> (10 * c)) & 1023u) - 256;
            } else {
              const size_t o = ((size_t)py * TW + px) * 3;
#pragma unroll
              for (int c = 0; c < 3; ++c) dl[c] = (int)col[o + c] - (int)bgr[o + c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const double v = s_rb[dl[c] + 255] * m;  // remove_bg * mask
              racc[c] += v * (double)wx;
            }
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) acc[c] += racc[c] * (double)wy;
        }
        // cvtColor(float32, BGR2GRAY) applied to an RGB-ordered array
        const float g = 0.114f * (float)acc[0] + 0.587f * (float)acc[1] + 0.299f * (float)acc[2];
        ob[oy * OBS_W + ox] = g;
      }
    }
    CT_T(8);
  }
}

// ---- standalone obs kernel (K3) for externally produced color images ------------------------
__global__ void tac_obs_kernel(const uint8_t* __restrict__ color, const uint8_t* __restrict__ bg_real,
                               const int32_t* __restrict__ bg_id, float* __restrict__ obs, int64_t obs_stride,
                               int n_frames) {
  const int f = blockIdx.x;
  const uint8_t* col = color + (size_t)f * TW * TH * 3;
  const uint8_t* bgr = bg_real + (size_t)bg_id[f] * TW * TH * 3;
  float* ob = obs + (size_t)f * obs_stride;
  for (int i = threadIdx.x; i < OBS_W * OBS_H; i += blockDim.x) {
    const int oy = i / OBS_W, ox = i % OBS_W;
    const int sx0 = (ox >> 1) * 7 + ((ox & 1) ? 3 : 0), sy0 = (oy >> 1) * 7 + ((oy & 1) ? 3 : 0);
    double acc[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 4; ++j) {
      const float wy = (oy & 1) ? (j == 0 ? AREA_W1 : AREA_W0) : (j == 3 ? AREA_W1 : AREA_W0);
      const int py = TH - 1 - (sy0 + j);
      double racc[3] = {0.0, 0.0, 0.0};
      for (int k = 0; k < 4; ++k) {
        const float wx = (ox & 1) ? (k == 0 ? AREA_W1 : AREA_W0) : (k == 3 ? AREA_W1 : AREA_W0);
        const int px = sx0 + k;
        const int ddx = px - TW / 2, ddy = py - TH / 2;
        const double m = (ddx * ddx + ddy * ddy <= (TW / 2) * (TW / 2)) ? 1.0 : 0.0;
        const size_t o = ((size_t)py * TW + px) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          racc[c] += (((double)((int)col[o + c] - (int)bgr[o + c]) / 255.0 + 0.5) * m) * (double)wx;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += racc[c] * (double)wy;
    }
    ob[i] = 0.114f * (float)acc[0] + 0.587f * (float)acc[1] + 0.299f * (float)acc[2];
  }
}

}  // namespace

// =============================================================================================
// C-ABI
// =============================================================================================
#ifndef FILL_GEOM_PARTS
#define FILL_GEOM_PARTS 5   // fill parts (1 colour, 2 gel_depth, 4 obs) written by tac_geom; the rest by tac_contact
#endif

// IgiSensorParams (host) -> the constant block every launch carries.  No library state: two engines with
// different sensor yamls, or on different devices, never see each other's constants.
static int make_const(const IgiSensorParams* p, TacConst* out) {
  IGI_REQUIRE(p != nullptr, "sensor params: null");
  IGI_REQUIRE(p->width == TW && p->height == TH, "sensor params: only 224x224 is built");
  IGI_REQUIRE(p->n_lights >= 0 && p->n_lights <= MAX_LIGHTS, "sensor params: at most 8 lights");
  IGI_REQUIRE(p->blur_ksize == 7, "sensor params: blur kernel size must be 7");
  IGI_REQUIRE(p->n_lights == 0 || (p->light_pos && p->light_dir && p->light_col && p->light_int && p->light_las && p->light_lao),
              "sensor params: null light array");
  TacConst c{};
  c.znear = p->znear;
  c.n_lights = p->n_lights;
  for (int i = 0; i < p->n_lights; ++i) {
    for (int k = 0; k < 3; ++k) {
      c.light_pos[i][k] = p->light_pos[3 * i + k];
      c.light_dir[i][k] = p->light_dir[3 * i + k];
      c.light_col[i][k] = p->light_col[3 * i + k];
    }
    c.light_int[i] = p->light_int[i];
    c.light_las[i] = p->light_las[i];
    c.light_lao[i] = p->light_lao[i];
  }
  c.inverse_square = p->inverse_square;
  for (int k = 0; k < 3; ++k) c.base[k] = p->base_color[k];
  c.metallic = p->metallic;
  c.roughness = p->roughness;
  {
    const float alpha = p->roughness * p->roughness;
    c.sh_a2 = alpha * alpha;
    float fmx = 0.f;
    bool gray = true;
    for (int k = 0; k < 3; ++k) {
      c.sh_f0[k] = 0.04f * (1.0f - p->metallic) + p->base_color[k] * p->metallic;
      c.sh_cdiff_pi[k] = p->base_color[k] * (1.0f - 0.04f) * (1.0f - p->metallic) * 0.31830988618379067f;
      fmx = c.sh_f0[k] > fmx ? c.sh_f0[k] : fmx;
      gray = gray && p->base_color[k] == p->base_color[0];
    }
    c.sh_f90 = fmx * 25.0f > 1.0f ? 1.0f : fmx * 25.0f;
    for (int i = 0; i < p->n_lights; ++i)
      for (int k = 0; k < 3; ++k) {
        c.sh_rad[i][k] = p->light_col[3 * i + k] * p->light_int[i];
        gray = gray && p->light_col[3 * i + k] == p->light_col[3 * i];
      }
    c.gray = gray ? 1 : 0;
  }
  for (int k = 0; k < 9; ++k) c.cam_R[k] = p->cam_R[k];
  for (int k = 0; k < 3; ++k) c.cam_p[k] = p->cam_p[k];
  c.gel_camx = (float)p->cam_p[0];
  c.max_force = p->max_force;
  c.max_deformation = p->max_deformation;
  c.calib_scale = p->calib_scale;
  c.clip_lo = p->clip_lo;
  c.clip_hi = p->clip_hi;
  for (int k = 0; k < 7; ++k) c.gauss[k] = p->gauss[k];
  for (int k = 0; k < 3; ++k) { c.grid_org[k] = p->grid_org[k]; c.grid_n[k] = p->grid_n[k]; }
  c.grid_h = p->grid_h;
  c.grid_slack = p->grid_slack;
  c.depth0_max = p->depth0_max;
  IGI_REQUIRE(p->hiz_levels >= 1 && p->hiz_levels <= 10, "sensor params: hiz_levels must be 1..10");
  c.hiz_levels = p->hiz_levels;
  for (int k = 0; k < p->hiz_levels; ++k) { c.hiz_off[k] = p->hiz_off[k]; c.hiz_w[k] = p->hiz_w[k]; }
  c.sx0 = p->dxp_first; c.sx1 = p->dxp_last; c.sy0 = p->dyp_first; c.sy1 = p->dyp_last;
  IGI_REQUIRE(c.sx1 != c.sx0 && c.sy1 != c.sy0, "sensor params: degenerate ray-slope range");
  *out = c;
  return IGI_OK;
}

extern "C" int igi_tactile_gel_precompute(const IgiSensorParams* sensor, const float* dxp, const float* dyp,
                                          const float* gel_tris, int n_tris, uint64_t* scratch_zbuf, float* depth0,
                                          uint8_t* bg_sim, void* stream) {
  IGI_REQUIRE(dxp && dyp && gel_tris && scratch_zbuf && depth0 && bg_sim && n_tris > 0, "igi_tactile_gel_precompute: bad args");
  TacConst kc;
  const int rc = make_const(sensor, &kc);
  if (rc != IGI_OK) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  IGI_CUDA(cudaMemsetAsync(scratch_zbuf, 0xff, sizeof(uint64_t) * TW * TH, s));
  tac_gel_raster<<<(n_tris + 127) / 128, 128, 0, s>>>(kc, dxp, dyp, gel_tris, n_tris, (unsigned long long*)scratch_zbuf);
  IGI_CHECK_LAUNCH("tac_gel_raster");
  tac_gel_shade<<<(TW * TH + 255) / 256, 256, 0, s>>>(kc, dxp, dyp, gel_tris, (const unsigned long long*)scratch_zbuf, depth0, bg_sim);
  IGI_CHECK_LAUNCH("tac_gel_shade");
  return IGI_OK;
}

extern "C" int igi_tactile_render(const IgiSensorParams* sensor, const IgiTactileMeshes* m, const IgiTactileStatic* st,
                                  const IgiTactileFrames* fr, const IgiTactileScratch* sc, const IgiTactileOut* out,
                                  void* stream) {
  IGI_REQUIRE(sensor && m && st && fr && sc && out, "igi_tactile_render: null struct");
  IGI_REQUIRE(fr->n_envs >= 0 && fr->sensors_per_env >= 1, "igi_tactile_render: bad frame counts");
  IGI_REQUIRE(fr->sensors_per_env <= 8, "igi_tactile_render: at most 8 sensors per env");
  IGI_REQUIRE(fr->plug_pos && fr->plug_quat && fr->mesh_id && fr->bg_id, "igi_tactile_render: null pose pointer");
  for (int n = 0; n < fr->sensors_per_env; ++n)
    IGI_REQUIRE((fr->finger_pos || fr->finger_pos_n[n]) && (fr->finger_quat || fr->finger_quat_n[n]),
                "igi_tactile_render: null fingertip pose pointer");
  IGI_REQUIRE((fr->finger_pos || fr->finger_pos_stride >= 3) && (fr->finger_quat || fr->finger_quat_stride >= 4),
              "igi_tactile_render: per-sensor fingertip views need their env strides");
  IGI_REQUIRE(m->verts && m->vnorm && m->faces && m->face_orig && m->meshes && m->clusters,
              "igi_tactile_render: null mesh pointer");
  IGI_REQUIRE(st->depth0 && st->bg_sim && st->bg_real && st->obs_empty && st->grid && st->hiz && st->dxp && st->dyp,
              "igi_tactile_render: null static pointer");
  IGI_REQUIRE(sc->normals, "igi_tactile_render: scratch.normals is null ((F,kmax) 48-byte records)");
  IGI_REQUIRE(sc->M && sc->setups && sc->counts && sc->bbox && sc->worklist && sc->counters && sc->kmax > 0 &&
                  sc->kmax <= 4096,
              "igi_tactile_render: bad scratch (kmax must be 1..4096)");
  IGI_REQUIRE(out->obs && out->obs_sensor_stride >= OBS_W * OBS_H &&
                  out->obs_env_stride >= out->obs_sensor_stride * fr->sensors_per_env &&
                  out->obs_env_stride % 4 == 0 && out->obs_sensor_stride % 4 == 0,
              "igi_tactile_render: bad obs output strides");
  IGI_REQUIRE(out->color && out->gel_depth, "igi_tactile_render: color and gel_depth outputs are required");
  IGI_REQUIRE(fr->fill_split >= 0 && fr->fill_split <= 8, "igi_tactile_render: fill_split must be 0 (default) or parts mask + 1");
  IGI_REQUIRE(fr->region_budget == 0 || fr->region_budget >= (2 * HALO + 1) * (2 * HALO + 1),
              "igi_tactile_render: region_budget must be 0 or >= 49 pixels");
  TacConst kc;
  {
    const int rc = make_const(sensor, &kc);
    if (rc != IGI_OK) return rc;
  }
  const int F = fr->n_envs * fr->sensors_per_env;
  if (F == 0) return IGI_OK;
  cudaStream_t s = (cudaStream_t)stream;
  // stage bits: 1 geometry alone, 2 standalone fill, 4 contact, 8 geometry with the fill fused in
  // (the product path: 0 = 8 | 4).  1/2 exist for per-stage timing and for callers that want one stage.
  const int stages = fr->stage_mask ? fr->stage_mask : (8 | 4);
  IGI_REQUIRE((stages & ~15) == 0 && !((stages & 8) && (stages & 3)),
              "igi_tactile_render: stage_mask 8 (fused geometry+fill) excludes bits 1 and 2");
  // In the fused modes (no bit 1 / 2) the fill is split between the two kernels: tac_geom writes
  // fill_geom_parts, tac_contact the rest (it then visits every live frame, not only those with candidates).
  const bool fused = (stages & 3) == 0;
  const int fill_geom_parts = fr->fill_split ? fr->fill_split - 1 : FILL_GEOM_PARTS;
  const int parts_geom = fused ? fill_geom_parts : 0, parts_contact = fused ? (7 & ~fill_geom_parts) : 0;
  // counters: [0] work_n, [1] cursor, [2] overflow (sticky), [3] largest triangle count of a frame (sticky);
  // the caller reads [2], [3] whenever it likes (e.g. an asynchronous copy per step) and clears them
  if (stages & 9) IGI_CUDA(cudaMemsetAsync(sc->counters, 0, 2 * sizeof(int32_t), s));
  else IGI_CUDA(cudaMemsetAsync(sc->counters + 1, 0, sizeof(int32_t), s));
  FillArgs fa{};
  fa.bg_real = st->bg_real; fa.bg_id = fr->bg_id; fa.obs_empty = st->obs_empty; fa.counts = sc->counts;
  fa.color = out->color; fa.gel_depth = out->gel_depth; fa.obs = out->obs;
  fa.obs_env_stride = out->obs_env_stride; fa.obs_sensor_stride = out->obs_sensor_stride;
  fa.sensors_per_env = fr->sensors_per_env;
  fa.n_frames = F;
  fa.parts = 7;
  GeomArgs g{};
  g.plug_pos = fr->plug_pos; g.plug_quat = fr->plug_quat;
  for (int n = 0; n < fr->sensors_per_env; ++n) {
    g.fpos_n[n] = fr->finger_pos ? fr->finger_pos + 3 * n : fr->finger_pos_n[n];
    g.fquat_n[n] = fr->finger_quat ? fr->finger_quat + 4 * n : fr->finger_quat_n[n];
  }
  g.fpos_stride = fr->finger_pos ? 3 * fr->sensors_per_env : fr->finger_pos_stride;
  g.fquat_stride = fr->finger_quat ? 4 * fr->sensors_per_env : fr->finger_quat_stride;
  g.force = fr->force; g.update = fr->update; g.update2 = fr->update2; g.mesh_id = fr->mesh_id;
  g.meshes = (const MeshInfo*)m->meshes; g.clusters = (const Cluster*)m->clusters;
  g.verts = m->verts; g.faces = m->faces; g.face_orig = m->face_orig; g.grid = st->grid; g.hiz = st->hiz; g.depth0 = st->depth0;
  g.dxp = st->dxp; g.dyp = st->dyp;
  g.M_out = sc->M; g.setups = (Setup*)sc->setups; g.vnorm = m->vnorm; g.nrecs = (NRec*)sc->normals; g.counts = sc->counts; g.bbox = sc->bbox;
  g.worklist = sc->worklist; g.work_n = sc->counters; g.overflow = sc->counters + 2;
  g.sensors_per_env = fr->sensors_per_env; g.kmax = sc->kmax; g.force_const = fr->force_const;
  g.fused_fill = (stages & 8) && parts_geom ? 1 : 0;
  g.list_all = parts_contact ? 1 : 0;
  g.fill = fa;
  g.fill.parts = parts_geom;
  if (stages & 9) {
    tac_geom<<<F, GEOM_BLOCK, 0, s>>>(kc, g);
    IGI_CHECK_LAUNCH("tac_geom");
  }
  if (stages & 2) {
    tac_fill<<<F * FILL_PARTS, FILL_BLOCK, 0, s>>>(fa);
    IGI_CHECK_LAUNCH("tac_fill");
  }
  ContactArgs ca{};
  ca.M = sc->M; ca.setups = (const Setup*)sc->setups; ca.nrecs = (const NRec*)sc->normals; ca.counts = sc->counts; ca.bbox = sc->bbox;
  ca.worklist = sc->worklist; ca.work_n = sc->counters; ca.cursor = sc->counters + 1;
  ca.verts = m->verts; ca.vnorm = m->vnorm; ca.faces = m->faces;
  ca.depth0 = st->depth0; ca.hiz = st->hiz; ca.dxp = st->dxp; ca.dyp = st->dyp;
  ca.bg_sim = st->bg_sim; ca.bg_real = st->bg_real; ca.bg_id = fr->bg_id;
  ca.color = out->color; ca.gel_depth = out->gel_depth; ca.obs = out->obs;
  ca.obs_env_stride = out->obs_env_stride; ca.obs_sensor_stride = out->obs_sensor_stride;
  ca.sensors_per_env = fr->sensors_per_env;
  ca.kmax = sc->kmax;
  ca.fill = fa;
  ca.fill.parts = parts_contact;
  const bool gray = kc.gray != 0;
  {
    const int compiled = gray ? CT_BUD_GRAY : CT_BUD_RGB;
    ca.budget = fr->region_budget > 0 && fr->region_budget < compiled ? fr->region_budget : compiled;
  }
  int dev = 0, sms = 148;
  IGI_CUDA(cudaGetDevice(&dev));
  IGI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (stages & 4) {
    constexpr size_t smem_gray = (size_t)CT_BUD_GRAY * 8, smem_rgb = (size_t)CT_BUD_RGB * (8 + 24);
    // the opt-in shared-memory size is a per-DEVICE function attribute
    static bool attr_set[64] = {false};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      IGI_CUDA(cudaFuncSetAttribute(tac_contact<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_gray));
      IGI_CUDA(cudaFuncSetAttribute(tac_contact<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rgb));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (gray) tac_contact<1><<<min(F, sms * CT_CTAS), CT_BLOCK, smem_gray, s>>>(kc, ca);
    else tac_contact<3><<<min(F, sms), CT_BLOCK, smem_rgb, s>>>(kc, ca);
    IGI_CHECK_LAUNCH("tac_contact");
  }
  return IGI_OK;
}

#ifdef CT_PROFILE
extern "C" int igi_debug_read_prof(unsigned long long* out32, int reset) {
  IGI_CUDA(cudaDeviceSynchronize());
  IGI_CUDA(cudaMemcpyFromSymbol(out32, g_ct_prof, sizeof(unsigned long long) * 32));
  if (reset) {
    unsigned long long z[32] = {0};
    IGI_CUDA(cudaMemcpyToSymbol(g_ct_prof, z, sizeof(z)));
  }
  return IGI_OK;
}
#endif

extern "C" int igi_tactile_obs(const uint8_t* color, const uint8_t* bg_real, const int32_t* bg_id, int n_frames,
                               float* obs, int64_t obs_stride, void* stream) {
  IGI_REQUIRE(color && bg_real && bg_id && obs && n_frames >= 0 && obs_stride >= OBS_W * OBS_H, "igi_tactile_obs: bad args");
  if (n_frames == 0) return IGI_OK;
  tac_obs_kernel<<<n_frames, 256, 0, (cudaStream_t)stream>>>(color, bg_real, bg_id, obs, obs_stride, n_frames);
  IGI_CHECK_LAUNCH("tac_obs_kernel");
  return IGI_OK;
}
