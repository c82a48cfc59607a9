/*
 * CPU ORACLE (test infrastructure, not product code): farthest-point sampling with the semantics of
 * pointnet2_ops' furthest_point_sampling_kernel (sampling_gpu.cu), the op behind `fps()` in
 * algo/models/transformer/point_mae.py:14-21.  PARITY UNPINNED: the package is third-party and absent; oracle/fps.py
 * states the algorithm and the tie rule, this file is its arithmetic.
 *
 * Floating point: the published source computes
 *     mag = x2*x2 + y2*y2 + z2*z2;   d = (x2-x1)*(x2-x1) + (y2-y1)*(y2-y1) + (z2-z1)*(z2-z1);
 * and is compiled by nvcc with its default -fmad=true, which contracts each expression to
 *     mul(y,y) -> fma(x,x,.) -> fma(z,z,.)
 * (checked in this image: `nvcc -ptx` of exactly these two lines emits mul.f32, fma.rn.f32, fma.rn.f32 in that
 * order; tests/test_fps_contraction.py repeats the check).  The explicit fmaf() calls below restate that; the file is built with -ffp-contract=off so nothing else
 * is contracted.  The CUDA kernels use __fmaf_rn / __fmul_rn in the same order, so indices compare bit-for-bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static uint32_t bitrev(uint32_t v, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

/* pts (n,3) f32 -> idx (m) i32.  Start index 0, temp = 1e10; per pick: points with |p|^2 <= 1e-3 are skipped,
 * temp[k] = min(temp[k], d(k, old)); the next pick is the arg-max of temp over the candidates, ties resolved like
 * the upstream block reduction: smallest bit-reversed (k mod B), then smallest k, B = largest power of two
 * <= min(n, 512).  No candidate -> index 0. */
int igi_oracle_fps(const float* pts, int n, int m, int32_t* idx) {
  for (int j = 0; j < m; ++j) idx[j] = 0;
  if (n <= 0 || m <= 0) return 0;
  int lg = 0;
  while ((2 << lg) <= n && lg < 9) ++lg;   /* floor(log2(n)), capped at 9 */
  const uint32_t B = 1u << lg;
  float* temp = (float*)malloc(sizeof(float) * (size_t)n);
  unsigned char* cand = (unsigned char*)malloc((size_t)n);
  uint64_t* key = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)n);
  if (!temp || !cand || !key) { free(temp); free(cand); free(key); return -1; }
  int any = 0;
  for (int k = 0; k < n; ++k) {
    const float x = pts[3 * k], y = pts[3 * k + 1], z = pts[3 * k + 2];
    const float mag = fmaf(z, z, fmaf(x, x, y * y));
    cand[k] = mag > 1e-3f;
    any |= cand[k];
    temp[k] = 1e10f;
    key[k] = ((uint64_t)bitrev((uint32_t)k & (B - 1u), lg) << 32) | (uint32_t)k;
  }
  int old = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = pts[3 * old], y1 = pts[3 * old + 1], z1 = pts[3 * old + 2];
    float best = -1.0f;
    uint64_t bkey = 0;
    int bi = 0;
    for (int k = 0; k < n; ++k) {
      if (!cand[k]) continue;
      const float dx = pts[3 * k] - x1, dy = pts[3 * k + 1] - y1, dz = pts[3 * k + 2] - z1;
      const float d = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
      const float d2 = d < temp[k] ? d : temp[k];
      temp[k] = d2;
      if (d2 > best || (d2 == best && key[k] < bkey)) { best = d2; bkey = key[k]; bi = k; }
    }
    old = any ? bi : 0;
    idx[j] = old;
  }
  free(temp); free(cand); free(key);
  return 0;
}
