/*
 * CPU ORACLE (test infrastructure, not product code): single-sample software
 * rasteriser + PBR shader standing in for the pyrender/OpenGL draw
 *   color, depth = self.r.render(self.scene, flags=0)
 * at isaacgyminsertion/allsight/tacto/renderer.py:642 (scene built at :137-183,
 * :186-237, :291-325 and tacto_allsight_wrapper/allsight_wrapper.py:100-174).
 *
 * PARITY UNPINNED for this stage: pyrender (>=0.1.43, unpinned, requirements.txt:6),
 * PyOpenGL and a GL driver are third-party and absent from /root/reference and from
 * this image; the reference ships no golden image.  This file restates pyrender's
 * published behaviour (perspective camera with infinite far plane, back-face culling,
 * GL_LESS depth test, metallic-roughness shader of shaders/mesh.frag, sRGB-ish
 * pow(1/2.2) output, 8-bit UNORM framebuffer, depth read-back in metres with 0 for
 * background) with ONE sample per pixel centre (the OSMesa path has no MSAA).
 *
 * Arithmetic contract shared with the CUDA kernels (DESIGN.md "raster spec"): all
 * coverage / depth maths is IEEE f32 with every operation rounded separately (compile
 * with -ffp-contract=off; the kernels use __fmul_rn/__fadd_rn), in exactly the order
 * written here, so coverage masks and depths can be compared bit-for-bit.  Shading
 * uses libm (powf, sqrtf) and is compared within 1/255.
 *
 * Like the reference, the whole scene (static gel mesh + posed peg) is rasterised for
 * every frame.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int W, H;
  float znear;
  const float* dxp; /* (W) ray slope x per column:  ((px+.5)/W*2-1) * tan(yfov/2) * aspect */
  const float* dyp; /* (H) ray slope y per row:     (1-(py+.5)/H*2) * tan(yfov/2)          */
  int n_lights;
  const float* light_pos; /* (L,3) camera frame */
  const float* light_dir; /* (L,3) camera frame, unit */
  const float* light_col; /* (L,3) */
  const float* light_int; /* (L)   */
  const float* light_las; /* (L) spot angle scale  */
  const float* light_lao; /* (L) spot angle offset */
  float base[3];
  float metallic, roughness;
  int inverse_square; /* 1: attenuate by 1/d^2 (pyrender's documented punctual-light model, the default); 0: no distance falloff (sensor yaml `lights.falloff: none`, documented deviation) */
} OracleCam;

typedef struct {
  float t;      /* depth along the view axis, INFINITY = empty */
  int kind;     /* 0 gel (flat), 1 peg (smooth) */
  int tri;
  float l1, l2; /* barycentrics of vertices B and C */
} Frag;

static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

static inline void cross3(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

/* tie rule for a pixel centre exactly on an edge plane: the edge belongs to the
 * triangle for which the first non-zero component of its edge normal is positive
 * (the neighbour across the edge has the exactly negated normal). */
static inline int edge_owns_zero(const float* n) {
  if (n[0] != 0.0f) return n[0] > 0.0f;
  if (n[1] != 0.0f) return n[1] > 0.0f;
  return n[2] > 0.0f;
}

static inline float clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void normalize3(float* v) {
  float l = sqrtf(dot3(v, v));
  if (l > 0.0f) {
    v[0] /= l; v[1] /= l; v[2] /= l;
  }
}

/* pyrender shaders/mesh.frag, metallic-roughness branch, spot lights, no shadows,
 * ambient 0.  p, n in camera frame (camera at the origin). */
static void shade(const OracleCam* c, const float* p, const float* n, uint8_t* rgb) {
  const float PI = 3.14159265358979323846f;
  float v[3] = {-p[0], -p[1], -p[2]};
  normalize3(v);
  float f0[3], cdiff[3], col[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < 3; ++k) {
    f0[k] = 0.04f * (1.0f - c->metallic) + c->base[k] * c->metallic;
    cdiff[k] = c->base[k] * (1.0f - 0.04f) * (1.0f - c->metallic);
  }
  float refl = fmaxf(fmaxf(f0[0], f0[1]), f0[2]);
  float f90 = clampf(refl * 25.0f, 0.0f, 1.0f);
  float alpha = c->roughness * c->roughness;
  float a2 = alpha * alpha;
  for (int i = 0; i < c->n_lights; ++i) {
    const float* lp = c->light_pos + 3 * i;
    const float* ld = c->light_dir + 3 * i;
    float L[3] = {lp[0] - p[0], lp[1] - p[1], lp[2] - p[2]};
    float d2 = dot3(L, L);
    float l[3] = {L[0], L[1], L[2]};
    normalize3(l);
    float h[3] = {l[0] + v[0], l[1] + v[1], l[2] + v[2]};
    normalize3(h);
    float nl = clampf(dot3(n, l), 0.001f, 1.0f);
    float nv = clampf(dot3(n, v), 0.001f, 1.0f);
    float nh = clampf(dot3(n, h), 0.001f, 1.0f);
    float vh = clampf(dot3(v, h), 0.001f, 1.0f);
    float ml[3] = {-l[0], -l[1], -l[2]};
    float cd = dot3(ld, ml);
    float att = clampf(cd * c->light_las[i] + c->light_lao[i], 0.0f, 1.0f);
    att = att * att;
    if (c->inverse_square) att = att / d2;
    float fw = powf(clampf(1.0f - vh, 0.0f, 1.0f), 5.0f);
    float al = 2.0f * nl / (nl + sqrtf(a2 + (1.0f - a2) * (nl * nl)));
    float av = 2.0f * nv / (nv + sqrtf(a2 + (1.0f - a2) * (nv * nv)));
    float G = al * av;
    float f = (nh * a2 - nh) * nh + 1.0f;
    float D = a2 / (PI * f * f);
    for (int k = 0; k < 3; ++k) {
      float F = f0[k] + (f90 - f0[k]) * fw;
      float diffuse = (1.0f - F) * cdiff[k] / PI;
      float spec = F * G * D / (4.0f * nl * nv);
      float radiance = att * c->light_col[3 * i + k] * c->light_int[i];
      col[k] += nl * radiance * (diffuse + spec);
    }
  }
  for (int k = 0; k < 3; ++k) {
    float o = clampf(powf(col[k], 1.0f / 2.2f), 0.0f, 1.0f);
    rgb[k] = (uint8_t)floorf(o * 255.0f + 0.5f);
  }
}

/* Rasterise one triangle (camera-frame vertices A,B,C) into the fragment buffer. */
static void raster_tri(const OracleCam* c, const float* A, const float* B, const float* C, int kind, int tri,
                       Frag* fb) {
  float E1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
  float E2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
  float N[3];
  cross3(E1, E2, N);
  float det = dot3(N, A);
  if (!(det < 0.0f)) return; /* back-facing or edge-on: culled (single-sided material) */
  float n0[3], n1[3], n2[3];
  cross3(B, C, n0);
  cross3(C, A, n1);
  cross3(A, B, n2);

  /* conservative pixel bounding box of the part of the triangle at depth >= znear (what GL
   * keeps after near-plane clipping); it need not match the GPU's box, only contain it */
  int x0, x1, y0, y1;
  {
    const float zn = c->znear;
    const float* P[3] = {A, B, C};
    float Z[3] = {-A[2], -B[2], -C[2]};
    if (fmaxf(Z[0], fmaxf(Z[1], Z[2])) < zn) return; /* entirely in front of the near plane */
    float mnx = 1e30f, mxx = -1e30f, mny = 1e30f, mxy = -1e30f;
    for (int i = 0; i < 3; ++i) {
      int j = (i + 1) % 3;
      if (Z[i] >= zn) {
        float sx = P[i][0] / Z[i], sy = P[i][1] / Z[i];
        mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
      }
      if ((Z[i] >= zn) != (Z[j] >= zn)) {
        float t = (zn - Z[i]) / (Z[j] - Z[i]);
        float sx = (P[i][0] + t * (P[j][0] - P[i][0])) / zn, sy = (P[i][1] + t * (P[j][1] - P[i][1])) / zn;
        mnx = fminf(mnx, sx); mxx = fmaxf(mxx, sx); mny = fminf(mny, sy); mxy = fmaxf(mxy, sy);
      }
    }
    float sx0 = c->dxp[0], sx1 = c->dxp[c->W - 1];
    float sy0 = c->dyp[0], sy1 = c->dyp[c->H - 1];
    float kx = (float)(c->W - 1) / (sx1 - sx0);
    float ky = (float)(c->H - 1) / (sy1 - sy0); /* negative: rows grow downwards */
    float fx0 = (mnx - sx0) * kx, fx1 = (mxx - sx0) * kx;
    float fy0 = (mxy - sy0) * ky, fy1 = (mny - sy0) * ky;
    if (fx1 < -2.0f || fy1 < -2.0f || fx0 > (float)(c->W + 1) || fy0 > (float)(c->H + 1)) return;
    x0 = (int)fmaxf(floorf(fx0) - 1.0f, 0.0f);
    y0 = (int)fmaxf(floorf(fy0) - 1.0f, 0.0f);
    x1 = (int)fminf(ceilf(fx1) + 1.0f, (float)(c->W - 1));
    y1 = (int)fminf(ceilf(fy1) + 1.0f, (float)(c->H - 1));
  }
  int own0 = edge_owns_zero(n0), own1 = edge_owns_zero(n1), own2 = edge_owns_zero(n2);
  for (int py = y0; py <= y1; ++py) {
    float dy = c->dyp[py];
    for (int px = x0; px <= x1; ++px) {
      float dx = c->dxp[px];
      float e0 = (dx * n0[0] + dy * n0[1]) - n0[2];
      float e1 = (dx * n1[0] + dy * n1[1]) - n1[2];
      float e2 = (dx * n2[0] + dy * n2[1]) - n2[2];
      if (!(e0 < 0.0f || (e0 == 0.0f && own0))) continue;
      if (!(e1 < 0.0f || (e1 == 0.0f && own1))) continue;
      if (!(e2 < 0.0f || (e2 == 0.0f && own2))) continue;
      float den = (dx * N[0] + dy * N[1]) - N[2];
      float t = det / den;
      if (!(t >= c->znear)) continue;
      Frag* f = fb + (size_t)py * c->W + px;
      if (t < f->t) { /* GL_LESS; equal depth keeps the earlier (lower-index) fragment */
        float s = (e0 + e1) + e2;
        f->t = t;
        f->kind = kind;
        f->tri = tri;
        f->l1 = e1 / s;
        f->l2 = e2 / s;
      }
    }
  }
}

/* sensor-frame gel vertex -> camera frame (camera at (cx,0,0) looking along +x, up +z):
 * R_cam0^T (v - cam_pos) with R_cam0 = euler xyz (90,0,-90) deg, a signed permutation. */
static inline void gel_to_cam(const float* v, float camx, float* o) {
  o[0] = -v[1];
  o[1] = v[2];
  o[2] = -(v[0] - camx);
}

static inline void xform(const float* M, const float* v, float* o) {
  for (int r = 0; r < 3; ++r) o[r] = ((M[4 * r] * v[0] + M[4 * r + 1] * v[1]) + M[4 * r + 2] * v[2]) + M[4 * r + 3];
}

/*
 * Render the scene.  gel_tris: (G,3,3) f32 sensor frame.  Peg: indexed mesh with
 * per-vertex normals in the object frame, M = (3,4) row-major object->camera (f32),
 * peg pointers may be NULL / nf=0 for the background render.
 * Outputs: color (H,W,3) u8, depth (H,W) f32 (0 where empty), kind (H,W) i8 (-1 empty,
 * 0 gel, 1 peg) — the latter is the coverage mask compared bit-for-bit with the GPU.
 */
int igi_oracle_render(const OracleCam* c, const float* gel_tris, int G, float gel_camx, const float* peg_v,
                      const float* peg_vn, const int32_t* peg_f, int nf, const float* M, uint8_t* color,
                      float* depth, int8_t* kind) {
  const size_t npx = (size_t)c->W * c->H;
  Frag* fb = (Frag*)malloc(npx * sizeof(Frag));
  if (!fb) return -1;
  for (size_t i = 0; i < npx; ++i) {
    fb[i].t = INFINITY;
    fb[i].kind = -1;
    fb[i].tri = -1;
    fb[i].l1 = fb[i].l2 = 0.f;
  }
  for (int g = 0; g < G; ++g) {
    float A[3], B[3], C[3];
    gel_to_cam(gel_tris + 9 * (size_t)g, gel_camx, A);
    gel_to_cam(gel_tris + 9 * (size_t)g + 3, gel_camx, B);
    gel_to_cam(gel_tris + 9 * (size_t)g + 6, gel_camx, C);
    raster_tri(c, A, B, C, 0, g, fb);
  }
  for (int t = 0; t < nf; ++t) {
    float A[3], B[3], C[3];
    xform(M, peg_v + 3 * (size_t)peg_f[3 * t], A);
    xform(M, peg_v + 3 * (size_t)peg_f[3 * t + 1], B);
    xform(M, peg_v + 3 * (size_t)peg_f[3 * t + 2], C);
    raster_tri(c, A, B, C, 1, t, fb);
  }
  for (int py = 0; py < c->H; ++py) {
    for (int px = 0; px < c->W; ++px) {
      size_t i = (size_t)py * c->W + px;
      const Frag* f = fb + i;
      if (kind) kind[i] = (int8_t)f->kind;
      if (f->kind < 0) {
        color[3 * i] = color[3 * i + 1] = color[3 * i + 2] = 255; /* scene bg_color white */
        depth[i] = 0.0f;
        continue;
      }
      depth[i] = f->t;
      float p[3] = {c->dxp[px] * f->t, c->dyp[py] * f->t, -f->t};
      float n[3];
      if (f->kind == 0) {
        float A[3], B[3], C[3];
        gel_to_cam(gel_tris + 9 * (size_t)f->tri, gel_camx, A);
        gel_to_cam(gel_tris + 9 * (size_t)f->tri + 3, gel_camx, B);
        gel_to_cam(gel_tris + 9 * (size_t)f->tri + 6, gel_camx, C);
        float E1[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
        float E2[3] = {C[0] - A[0], C[1] - A[1], C[2] - A[2]};
        cross3(E1, E2, n); /* flat shading (smooth=False, renderer.py:177) */
        normalize3(n);
      } else {
        const int32_t* fi = peg_f + 3 * (size_t)f->tri;
        float l0 = (1.0f - f->l1) - f->l2;
        float no[3];
        for (int k = 0; k < 3; ++k)
          no[k] = (l0 * peg_vn[3 * (size_t)fi[0] + k] + f->l1 * peg_vn[3 * (size_t)fi[1] + k]) +
                  f->l2 * peg_vn[3 * (size_t)fi[2] + k];
        for (int r = 0; r < 3; ++r) n[r] = (M[4 * r] * no[0] + M[4 * r + 1] * no[1]) + M[4 * r + 2] * no[2];
        normalize3(n);
      }
      shade(c, p, n, color + 3 * i);
    }
  }
  free(fb);
  return 0;
}
