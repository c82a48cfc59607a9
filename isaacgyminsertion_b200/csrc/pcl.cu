// (P) external-camera point-cloud path: K4 unproject/filter/compact, K5A gather, K5B FPS.
// Entry points are declared in include/igi_b200.h.
#include <cooperative_groups.h>

#include "igi_common.cuh"
#include "../../include/igi_b200.h"

namespace cg = cooperative_groups;

namespace {

constexpr int kCompactBlock = 256;
constexpr int kMaxClasses = 4;

struct CompactParams {
  const float* depth;
  const int32_t* seg;
  const float* uvx;
  const float* uvy;
  const float* uvz;
  const float* ext;
  const float* e2g_inv;
  float* out_pts;
  int32_t* out_count;
  int32_t* out_any;
  int n_envs, H, W, n_classes;
  int seg_ids[kMaxClasses];
  float depth_max;  // < 0: disabled
  int has_box;
  float box[6];
  int use_bulk;  // 1: TMA bulk copy of the env's depth/seg rows into smem
};

// One CTA per env.  The env's depth and segmentation rows (H*W*4 B each, 20 736 B
// for 54x96) are staged into shared memory with two 1-D bulk async copies (TMA
// engine) that complete on an mbarrier; every class (plug, socket) is then
// produced from the staged tile, so HBM sees each input byte once.
// Compaction keeps row-major pixel order (pcl_utils.py:71-72 boolean indexing); a thread owns a run of
// consecutive pixels, so thread order = pixel order:
//   pass 1   an ordered list of the pixels of the wanted class (two compares per image pixel, one block scan)
//   pass 2   the listed pixels are unprojected / transformed / box-tested on dense lanes and the kept points
//            stored at their ordered offsets (one block scan per chunk of 512 list entries)
struct CompactEval {
  const float* A;   // ext, row-vector: w = [px,py,pz,1] @ A
  const float* B;   // inverse(env_to_global): o = w @ B^T
  const float* uvx;
  const float* uvy;
  float uvz, depth_max;
  int W, has_box;
  uint32_t inv_w;   // floor(i / W) = umulhi(i, inv_w) for i < 2^16
  float box[6];
  // point of pixel i with (masked) depth d; false when dropped by `d > -depth_max` or the box
  __device__ __forceinline__ bool operator()(int i, float d, float* o) const {
    if (!(depth_max < 0.0f) && !(d > -depth_max)) return false;
    const int v = (int)__umulhi((uint32_t)i, inv_w), u = i - v * W;
    const float px = __fmul_rn(uvx[u], d);
    const float py = __fmul_rn(uvy[v], d);
    const float pz = __fmul_rn(uvz, d);
    // w = [px py pz 1] @ ext   (pcl_utils.py:77-83)
    float w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = fmaf(pz, A[8 + j], fmaf(py, A[4 + j], fmaf(px, A[j], A[12 + j])));
    // o = w @ inverse(env_to_global)^T   (pcl_utils.py:84-85)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[j] = fmaf(w[3], B[4 * j + 3], fmaf(w[2], B[4 * j + 2], fmaf(w[1], B[4 * j + 1], w[0] * B[4 * j])));
    if (has_box)
      return (o[0] >= box[0]) && (o[0] <= box[1]) && (o[1] >= box[2]) && (o[1] <= box[3]) && (o[2] >= box[4]) &&
             (o[2] <= box[5]);
    return true;
  }
};

// Block-wide exclusive scan of one int per thread (kCompactBlock threads, thread order); `total` = the sum.
__device__ __forceinline__ int compact_block_scan(int v, int* s_wsum, int& total) {
  constexpr int NW = kCompactBlock / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int sft = 1; sft < 32; sft <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, sft);
    if (lane >= sft) incl += t;
  }
  __syncthreads();   // the previous scan's readers of s_wsum are done
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int woff = 0;
  total = 0;
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    const int t = s_wsum[w];
    if (w < warp) woff += t;
    total += t;
  }
  return woff + incl - v;
}

__global__ void __launch_bounds__(kCompactBlock) pcl_compact_kernel(CompactParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = kCompactBlock / 32;
  const int npix = p.H * p.W;
  const int seg_off = (npix * 4 + 15) & ~15;
  float* s_depth = reinterpret_cast<float*>(smem_raw);
  int32_t* s_seg = reinterpret_cast<int32_t*>(smem_raw + seg_off);
  uint16_t* s_idx = reinterpret_cast<uint16_t*>(smem_raw + 2 * (size_t)seg_off);   // ordered list of the pixels to evaluate
  __shared__ uint64_t s_bar;
  __shared__ int s_wsum[NW];
  __shared__ float s_m[32];  // ext (16) + e2g_inv (16)

  const int env = blockIdx.x;
  const int tid = threadIdx.x;
  const float* g_depth = p.depth + (size_t)env * npix;
  const int32_t* g_seg = p.seg ? p.seg + (size_t)env * npix : nullptr;

  if (p.use_bulk) {
    if (tid == 0) {
      igi_mbar_init(&s_bar, 1);
      igi_fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)npix * 4u;
      igi_mbar_expect_tx(&s_bar, g_seg ? 2u * bytes : bytes);
      igi_bulk_g2s(s_depth, g_depth, bytes, &s_bar);
      if (g_seg) igi_bulk_g2s(s_seg, g_seg, bytes, &s_bar);
    }
  } else {
    for (int i = tid; i < npix; i += kCompactBlock) {
      s_depth[i] = g_depth[i];
      if (g_seg) s_seg[i] = g_seg[i];
    }
  }
  if (tid < 16) s_m[tid] = p.ext[(size_t)env * 16 + tid];
  else if (tid < 32) s_m[tid] = p.e2g_inv[(size_t)env * 16 + (tid - 16)];
  if (p.use_bulk) igi_mbar_wait(&s_bar, 0);
  __syncthreads();

  CompactEval ev;
  ev.A = s_m; ev.B = s_m + 16;
  ev.uvx = p.uvx + (size_t)env * p.W;
  ev.uvy = p.uvy + (size_t)env * p.H;
  ev.uvz = p.uvz[env];
  ev.depth_max = p.depth_max;
  ev.W = p.W; ev.has_box = p.has_box;
  ev.inv_w = 0xffffffffu / (uint32_t)p.W + 1u;
#pragma unroll
  for (int k = 0; k < 6; ++k) ev.box[k] = p.box[k];
  const int NC = p.n_classes;

  // A masked-out pixel with finite depth becomes d = +-0 (pcl_utils.py / task :956-959) and maps to the
  // camera centre whatever its (u, v): w = ext[3,:], o = w @ e2g_inv^T.  Signs of zero do not change
  // the inclusive box test, so unless that one point is inside the box (it never is for the Factory
  // camera, x = 0.731 > 0.7) every masked-out pixel can be skipped without evaluating it; a
  // masked-out miss (-inf * 0 = NaN) is dropped by `d > -depth_max` either way.
  bool zero_kept = true;
  if (g_seg && p.has_box) {
    const float* A = ev.A;
    const float* B = ev.B;
    float o[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
      o[j] = fmaf(A[15], B[4 * j + 3], fmaf(A[14], B[4 * j + 2], fmaf(A[13], B[4 * j + 1], A[12] * B[4 * j])));
    zero_kept = (o[0] >= p.box[0]) && (o[0] <= p.box[1]) && (o[1] >= p.box[2]) && (o[1] <= p.box[3]) &&
                (o[2] >= p.box[4]) && (o[2] <= p.box[5]);
  }
  const bool every_pixel = !g_seg || zero_kept;

  // Thread t owns the consecutive pixels [t*ppt, (t+1)*ppt): thread order = row-major pixel order, which both
  // ordered compactions below keep (pcl_utils.py:71-72 boolean indexing).
  const int ppt = (npix + kCompactBlock - 1) / kCompactBlock;
  const int i_lo = min(tid * ppt, npix), i_hi = min(i_lo + ppt, npix);
  constexpr int R = 2;   // list entries per thread and pass-2 chunk
  for (int c = 0; c < NC; ++c) {
    const int sid_c = p.seg_ids[c];
    // ---- pass 1: the ordered list of the pixels of this class (two compares per image pixel; the plug / socket
    // blobs are ~7 % of the image, so everything after this pass runs on dense lanes)
    int cnt = 0;
    if (every_pixel) cnt = i_hi - i_lo;
    else for (int i = i_lo; i < i_hi; ++i) cnt += (s_seg[i] == sid_c) ? 1 : 0;
    int n_list;
    int off = compact_block_scan(cnt, s_wsum, n_list);
    if (every_pixel) { for (int i = i_lo; i < i_hi; ++i) s_idx[off++] = (uint16_t)i; }
    else { for (int i = i_lo; i < i_hi; ++i) if (s_seg[i] == sid_c) s_idx[off++] = (uint16_t)i; }
    __syncthreads();
    // ---- pass 2: unproject / transform / box-test the listed pixels, R consecutive entries per thread and chunk,
    // and store the kept points at their ordered offsets (one block scan per chunk; plug and socket need one chunk)
    float* out = p.out_pts + ((size_t)env * NC + c) * (size_t)npix * 3;
    int base = 0;
    bool nz = false;   // this thread kept a point with a non-zero coordinate (pcl_utils.py:179 `pts.any()`)
    for (int j0 = 0; j0 < n_list; j0 += kCompactBlock * R) {
      float o[R][3];
      int keep = 0;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int j = j0 + tid * R + r;
        if (j < n_list) {
          const int i = s_idx[j];
          float dd = s_depth[i];
          if (g_seg) dd = __fmul_rn(dd, s_seg[i] == sid_c ? 1.0f : 0.0f);   // -inf*0 = NaN, finite*0 = -0
          if (ev(i, dd, o[r])) {
            keep |= 1 << r;
            nz = nz || (o[r][0] != 0.0f) || (o[r][1] != 0.0f) || (o[r][2] != 0.0f);
          }
        }
      }
      int kept;
      int slot = base + compact_block_scan(__popc(keep), s_wsum, kept);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if ((keep >> r) & 1) {
          float* dst = out + (size_t)slot * 3;
          dst[0] = o[r][0];
          dst[1] = o[r][1];
          dst[2] = o[r][2];
          ++slot;
        }
      }
      base += kept;
    }
    const int bany = __syncthreads_or(nz ? 1 : 0);   // also: every reader of s_idx is done before the next class rewrites it
    if (tid == 0) {
      p.out_count[(size_t)env * NC + c] = base;
      p.out_any[(size_t)env * NC + c] = bany ? 1 : 0;
    }
  }
}

// ---- K5A ---------------------------------------------------------------------
// offsets[e] = m * #{e' < e : any[e']}; one CTA, block scan over envs.
__global__ void __launch_bounds__(1024) pcl_offsets_kernel(const int32_t* any, int n_classes, int cls,
                                                           int n_envs, int m, int32_t* offsets,
                                                           int32_t* consumed) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_envs; base += 1024) {
    const int e = base + tid;
    const int f = (e < n_envs) ? (any[(size_t)e * n_classes + cls] != 0) : 0;
    int incl = f;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, s);
      if (lane >= s) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
    for (int wi = 0; wi < 32; ++wi) {
      const int t = s_w[wi];
      if (wi < warp) woff += t;
      tot += t;
    }
    const int carry = s_carry;
    if (e < n_envs) offsets[e] = (carry + woff + incl - f) * m;
    __syncthreads();
    if (tid == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (tid == 0 && consumed) *consumed = s_carry * m;
}

__global__ void __launch_bounds__(128) pcl_gather_kernel(const float* pts, const int32_t* count,
                                                         const int32_t* any, int n_classes, int cls, int cap,
                                                         const uint32_t* raw, const int32_t* offsets, int m,
                                                         float* out, int64_t out_stride, int32_t* out_idx) {
  const int env = blockIdx.x;
  const size_t t = (size_t)env * n_classes + cls;
  const int n = count[t];
  const bool live = any[t] != 0 && n > 0;
  const float* src = pts + t * (size_t)cap * 3;
  const uint32_t* r = raw + offsets[env];
  float* dst = out + (size_t)env * out_stride;
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    int id = 0;
    float x = 0.f, y = 0.f, z = 0.f;
    if (live) {
      id = (int)(r[i] % (uint32_t)n);  // ATen uniform_int_from_to: random() % range
      x = src[(size_t)id * 3 + 0];
      y = src[(size_t)id * 3 + 1];
      z = src[(size_t)id * 3 + 2];
    }
    dst[(size_t)i * 3 + 0] = x;
    dst[(size_t)i * 3 + 1] = y;
    dst[(size_t)i * 3 + 2] = z;
    if (out_idx) out_idx[(size_t)env * m + i] = id;
  }
}

// ---- K5B FPS -------------------------------------------------------------------
constexpr int kFpsBlock = 128;

__device__ __forceinline__ uint32_t bitrev_n(uint32_t v, int bits) { return __brev(v) >> (32 - bits); }
// a*a + b*b + c*c in the form nvcc's default -fmad=true gives the published kernel's source expression
// (mul(b,b), fma(a,a,.), fma(c,c,.) — oracle/fps.c header); oracle/fps.c evaluates the same three operations.
__device__ __forceinline__ float fps_sumsq(float a, float b, float c) {
  return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}

// Candidate ordering: larger distance wins; among equal distances the upstream
// block reduction keeps the candidate of the thread whose id has the smaller
// bit-reversed value, and inside a thread the smaller k (strict '>' while striding).
// Packed so that a plain unsigned max picks the winner:
//   hi = float_bits(d)+1 (0 = "no candidate" -> index 0), lo = ~tiekey.
struct FpsCand {
  uint32_t hi, lo;
};

__global__ void __launch_bounds__(kFpsBlock) fps_kernel(const float* pts, int64_t task_stride,
                                                        const int32_t* count, const int32_t* any,
                                                        int64_t count_stride, int n_fixed, int n_tasks, int m,
                                                        float* out_pts, int64_t out_stride, int32_t* out_idx,
                                                        int min_n, const int32_t* sched, const int32_t* order,
                                                        int cluster_limit, int cluster_maxn) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ FpsCand s_cand[2][kFpsBlock / 32];
  // a short list of big tasks belongs to fps_cluster_kernel up to its size limit (same rule there)
  if (sched && cluster_limit > 0 && sched[4] <= cluster_limit && min_n <= cluster_maxn) min_n = cluster_maxn + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // persistent CTAs stride over the tasks; tasks below min_n belong to fps_warp_kernel / fps_sorted_kernel.
  // With a schedule (igi_fps_balanced) only the tasks fps_order_kernel listed as big are visited.
  const int n_iter = sched ? sched[4] : n_tasks;
  for (int it = blockIdx.x; it < n_iter; it += gridDim.x) {
    const int task = sched ? order[n_tasks - 1 - it] : it;
    const int n = count ? count[(size_t)task * count_stride] : n_fixed;
    if (n < min_n) continue;
    const bool live = (any ? any[(size_t)task * count_stride] != 0 : true) && n > 0;
    float* dst = out_pts ? out_pts + (size_t)task * out_stride : nullptr;
    int32_t* idst = out_idx ? out_idx + (size_t)task * m : nullptr;
    if (!live) {
      for (int i = tid; i < m; i += kFpsBlock) {
        if (dst) { dst[i * 3 + 0] = 0.f; dst[i * 3 + 1] = 0.f; dst[i * 3 + 2] = 0.f; }
        if (idst) idst[i] = 0;
      }
      continue;
    }
    // smem: x[n] y[n] z[n] temp[n] (cap = n rounded up)
    const int cap = (n + 3) & ~3;
    float* sx = reinterpret_cast<float*>(smem_raw);
    float* sy = sx + cap;
    float* sz = sy + cap;
    float* st = sz + cap;
    __syncthreads();  // previous task's readers are done with the arrays
    const float* src = pts + (size_t)task * task_stride;
    for (int k = tid; k < n; k += kFpsBlock) {
      sx[k] = src[(size_t)k * 3 + 0];
      sy[k] = src[(size_t)k * 3 + 1];
      sz[k] = src[(size_t)k * 3 + 2];
      st[k] = 1e10f;
    }
    // upstream block size: largest power of two <= min(n, 512)
    int lg = 31 - __clz(n);
    if (lg > 9) lg = 9;
    const uint32_t bmask = (1u << lg) - 1u;
    __syncthreads();

    int old = 0;
    if (tid == 0) {
      if (idst) idst[0] = 0;
      if (dst) { dst[0] = sx[0]; dst[1] = sy[0]; dst[2] = sz[0]; }
    }
    int j = 1;
    for (; j < m; ++j) {
      const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
      uint32_t bhi = 0, blo = 0;
      for (int k = tid; k < n; k += kFpsBlock) {
        const float x2 = sx[k], y2 = sy[k], z2 = sz[k];
        const float mag = fps_sumsq(x2, y2, z2);
        if (mag <= 1e-3f) continue;
        const float dx = __fsub_rn(x2, x1), dy = __fsub_rn(y2, y1), dz = __fsub_rn(z2, z1);
        const float d = fps_sumsq(dx, dy, dz);
        const float d2 = fminf(d, st[k]);
        st[k] = d2;
        const uint32_t hi = __float_as_uint(d2) + 1u;
        const uint32_t rev = lg ? bitrev_n((uint32_t)k & bmask, lg) : 0u;
        const uint32_t lo = ~((rev << 16) | (uint32_t)k);
        if (hi > bhi || (hi == bhi && lo > blo)) { bhi = hi; blo = lo; }
      }
      // warp arg-max with two REDUX passes
      const uint32_t whi = __reduce_max_sync(0xffffffffu, bhi);
      const uint32_t wlo = __reduce_max_sync(0xffffffffu, (bhi == whi) ? blo : 0u);
      if (lane == 0) { s_cand[j & 1][warp].hi = whi; s_cand[j & 1][warp].lo = wlo; }
      __syncthreads();
      uint32_t fhi = 0, flo = 0;
#pragma unroll
      for (int wi = 0; wi < kFpsBlock / 32; ++wi) {
        const FpsCand c = s_cand[j & 1][wi];
        if (c.hi > fhi || (c.hi == fhi && c.lo > flo)) { fhi = c.hi; flo = c.lo; }
      }
      old = (fhi == 0) ? 0 : (int)((~flo) & 0xffffu);
      if (tid == 0) {
        if (idst) idst[j] = old;
        if (dst) { dst[j * 3 + 0] = sx[old]; dst[j * 3 + 1] = sy[old]; dst[j * 3 + 2] = sz[old]; }
      }
      // Every remaining distance is 0 (or there is no candidate): the running minima can no longer
      // change, so every later pick is this same index.
      if (fhi <= 1u) { ++j; break; }
    }
    for (int i = j + tid; i < m; i += kFpsBlock) {
      if (idst) idst[i] = old;
      if (dst) { dst[i * 3 + 0] = sx[old]; dst[i * 3 + 1] = sy[old]; dst[i * 3 + 2] = sz[old]; }
    }
  }
}

// ---- resident FPS: tasks of up to FW_COOP_MAXN points --------------------------------------------
// One CTA = FW_WARPS warps owns FW_WARPS consecutive tasks.  A task of up to FW_MAXN points is run by
// ONE warp (points in the warp's shared-memory planes, running minima and tie keys in registers, the
// arg-max is two REDUX instructions, no block barrier); a task of FW_MAXN+1 .. FW_COOP_MAXN points is run
// by ALL warps of the CTA together first (one barrier per pick), so that the few long tasks get four
// warps' worth of issue slots instead of becoming the tail of the launch.  Same selection rule as
// fps_kernel and oracle/fps.py.
//
// Inner loop, branch free: a point that can never be picked (|p|^2 <= 1e-3, or beyond n) keeps a running
// minimum of -1, distances are >= 0, so a float max over min(d, temp) ignores it; the winner among equal
// distances is then the largest tie key among the points whose minimum equals the max.
// FW_MAXN: largest task one warp runs alone.  Above it the four warps of the CTA share the task: the total work per
// pick is about the same (4 warps x n/128 points vs 1 warp x n/32), but the chain of dependent picks - which is what
// the few longest tasks of a step cost - gets 3-4x shorter.
constexpr int FW_WARPS = 4, FW_MAXN = 256, FW_COOP_MAXN = 1024, FW_COOP_PPL = FW_COOP_MAXN / (FW_WARPS * 32);

__device__ __forceinline__ uint32_t fps_tie_key(int k, int lg, uint32_t bmask) {
  const uint32_t rev = lg ? bitrev_n((uint32_t)k & bmask, lg) : 0u;
  return ~((rev << 16) | (uint32_t)k);
}

// PPL points per lane, point index of slot i = first + stride * i
template <int PPL>
__device__ __forceinline__ void fps_init_state(const float* sx, const float* sy, const float* sz, int n, int first,
                                               int stride, float* temp, uint32_t* lokey) {
  int lg = 31 - __clz(n);
  if (lg > 9) lg = 9;
  const uint32_t bmask = (1u << lg) - 1u;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const int k = first + stride * i;
    temp[i] = -1.0f;
    lokey[i] = 0u;
    if (k < n) {
      const float x2 = sx[k], y2 = sy[k], z2 = sz[k];
      const float mag = fps_sumsq(x2, y2, z2);
      if (mag > 1e-3f) {
        temp[i] = 1e10f;
        lokey[i] = fps_tie_key(k, lg, bmask);
      }
    }
  }
}

// one pick, per-thread part: update the minima against point `old`, return the thread's best minimum
template <int PPL>
__device__ __forceinline__ float fps_update(const float* sx, const float* sy, const float* sz, int old, int first,
                                            int stride, float* temp) {
  const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
  float best = -1.0f;
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const int k = first + stride * i;   // planes are padded: reads beyond n are harmless (min with -1 stays -1)
    const float dx = __fsub_rn(sx[k], x1), dy = __fsub_rn(sy[k], y1), dz = __fsub_rn(sz[k], z1);
    const float d = fps_sumsq(dx, dy, dz);
    const float d2 = fminf(d, temp[i]);
    temp[i] = d2;
    best = fmaxf(best, d2);
  }
  return best;
}

// One pick's distance update for PPL register-resident points against (x1,y1,z1); returns the thread's best minimum.
// Two points per instruction with Blackwell's packed f32x2 add / mul / fma (FADD2 / FMUL2 / FFMA2 with the picked
// point's negated coordinate as a broadcast scalar operand): each half rounds to nearest like the scalar operation and
// x - x1 == x + (-x1) exactly, so the result is bit for bit that of fps_sumsq on the differences.  Measured at 4096 envs:
// plug + socket FPS 0.257 -> 0.220 ms.
template <int PPL>
__device__ __forceinline__ float fps_update_reg(const float* px, const float* py, const float* pz, float x1, float y1,
                                                float z1, float* temp) {
  float best = -1.0f;
  const float2 nx = make_float2(-x1, -x1), ny = make_float2(-y1, -y1), nz = make_float2(-z1, -z1);
#pragma unroll
  for (int i = 0; i + 1 < PPL; i += 2) {
    const float2 dx = __fadd2_rn(make_float2(px[i], px[i + 1]), nx);
    const float2 dy = __fadd2_rn(make_float2(py[i], py[i + 1]), ny);
    const float2 dz = __fadd2_rn(make_float2(pz[i], pz[i + 1]), nz);
    const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
    const float a = fminf(d.x, temp[i]), b = fminf(d.y, temp[i + 1]);
    temp[i] = a; temp[i + 1] = b;
    best = fmaxf(best, fmaxf(a, b));
  }
  if (PPL & 1) {
    constexpr int i = PPL - 1;
    const float dx = __fsub_rn(px[i], x1), dy = __fsub_rn(py[i], y1), dz = __fsub_rn(pz[i], z1);
    const float d2 = fminf(fps_sumsq(dx, dy, dz), temp[i]);
    temp[i] = d2;
    best = fmaxf(best, d2);
  }
  return best;
}

template <int PPL>
__device__ __forceinline__ uint32_t fps_tie(const float* temp, const uint32_t* lokey, int wbest) {
  uint32_t blo = 0u;
#pragma unroll
  for (int i = 0; i < PPL; ++i)
    if (__float_as_int(temp[i]) == wbest) blo = max(blo, lokey[i]);
  return blo;
}

// The m-1 dependent picks of one warp-resident task.  Fills sel[0..m).  Once the winning distance is 0
// (all points already picked, or only duplicates left) or nothing is a candidate, the minima cannot
// change any more and every later pick repeats the same index, so the loop stops there.
// REG: the lane's points live in registers for the whole task (3 * PPL registers), so a pick reads shared
// memory only for the coordinates of the point picked last; used while PPL is small enough for the 64-register cap.
template <int PPL, bool REG>
__device__ __forceinline__ void fps_warp_picks(const float* sx, const float* sy, const float* sz, int n, int m,
                                               int lane, unsigned short* sel) {
  float temp[PPL];
  uint32_t lokey[PPL];
  fps_init_state<PPL>(sx, sy, sz, n, lane, 32, temp, lokey);
  float px[REG ? PPL : 1], py[REG ? PPL : 1], pz[REG ? PPL : 1];
  if (REG) {
#pragma unroll
    for (int i = 0; i < PPL; ++i) { px[i] = sx[lane + 32 * i]; py[i] = sy[lane + 32 * i]; pz[i] = sz[lane + 32 * i]; }
  }
  int old = 0;
  if (lane == 0) sel[0] = 0;
  int j = 1;
  for (; j < m; ++j) {
    float best;
    if (REG) {
      best = fps_update_reg<PPL>(px, py, pz, sx[old], sy[old], sz[old], temp);
    } else {
      best = fps_update<PPL>(sx, sy, sz, old, lane, 32, temp);
    }
    const int wb = __reduce_max_sync(0xffffffffu, __float_as_int(best));   // floats >= 0 order as ints; -1 < 0
    const uint32_t wlo = __reduce_max_sync(0xffffffffu, fps_tie<PPL>(temp, lokey, wb));
    old = (wb < 0) ? 0 : (int)((~wlo) & 0xffffu);
    if (lane == 0) sel[j] = (unsigned short)old;
    if (wb <= 0) { ++j; break; }
  }
  for (int i = j + lane; i < m; i += 32) sel[i] = (unsigned short)old;
}

struct FpsTaskArgs {
  const float* pts;
  int64_t task_stride;
  const int32_t* count;
  const int32_t* any;
  int64_t count_stride;
  int n_fixed, n_tasks, m;
  float* out_pts;
  int64_t out_stride;
  int32_t* out_idx;
};
constexpr int FW_PLANE = FW_WARPS * FW_MAXN > FW_COOP_MAXN ? FW_WARPS * FW_MAXN : FW_COOP_MAXN;

// A task of FW_MAXN+1 .. FW_COOP_MAXN points, run by all warps of the CTA (one barrier per pick); PPL points per
// thread, held in registers.
template <int PPL>
__device__ __forceinline__ void fps_coop_picks(const FpsTaskArgs& a, int task, int n, const float* sx, const float* sy,
                                               const float* sz, int2 (*s_cand)[FW_WARPS]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m = a.m;
  float temp[PPL];
  uint32_t lokey[PPL];
  fps_init_state<PPL>(sx, sy, sz, n, tid, FW_WARPS * 32, temp, lokey);
  float px[PPL], py[PPL], pz[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const int k = tid + FW_WARPS * 32 * i;
    px[i] = sx[k]; py[i] = sy[k]; pz[i] = sz[k];
  }
  float* dst = a.out_pts ? a.out_pts + (size_t)task * a.out_stride : nullptr;
  int32_t* idst = a.out_idx ? a.out_idx + (size_t)task * m : nullptr;
  int old = 0;
  if (tid == 0) {
    if (idst) idst[0] = 0;
    if (dst) { dst[0] = sx[0]; dst[1] = sy[0]; dst[2] = sz[0]; }
  }
  int j = 1;
  for (; j < m; ++j) {
    const float best = fps_update_reg<PPL>(px, py, pz, sx[old], sy[old], sz[old], temp);
    const int wb = __reduce_max_sync(0xffffffffu, __float_as_int(best));
    const uint32_t wlo = __reduce_max_sync(0xffffffffu, fps_tie<PPL>(temp, lokey, wb));
    if (lane == 0) s_cand[j & 1][warp] = make_int2(wb, (int)wlo);
    __syncthreads();
    int fb = -2;
    uint32_t flo = 0u;
#pragma unroll
    for (int w = 0; w < FW_WARPS; ++w) {
      const int2 c = s_cand[j & 1][w];
      if (c.x > fb || (c.x == fb && (uint32_t)c.y > flo)) { fb = c.x; flo = (uint32_t)c.y; }
    }
    old = (fb < 0) ? 0 : (int)((~flo) & 0xffffu);
    if (tid == 0) {
      if (idst) idst[j] = old;
      if (dst) { dst[j * 3 + 0] = sx[old]; dst[j * 3 + 1] = sy[old]; dst[j * 3 + 2] = sz[old]; }
    }
    if (fb <= 0) { ++j; break; }
  }
  for (int i = j + tid; i < m; i += FW_WARPS * 32) {
    if (idst) idst[i] = old;
    if (dst) { dst[i * 3 + 0] = sx[old]; dst[i * 3 + 1] = sy[old]; dst[i * 3 + 2] = sz[old]; }
  }
}

__device__ __forceinline__ void fps_coop_task(const FpsTaskArgs& a, int task, int n, float* s_p, int2 (*s_cand)[FW_WARPS]) {
  const int tid = threadIdx.x;
  float* sx = s_p;
  float* sy = sx + FW_PLANE;
  float* sz = sy + FW_PLANE;
  __syncthreads();
  const float* src = a.pts + (size_t)task * a.task_stride;
  for (int i = tid; i < 3 * n; i += FW_WARPS * 32) {
    const int k = i / 3, c = i - 3 * k;
    s_p[c * FW_PLANE + k] = src[i];
  }
  int ppl = (n + FW_WARPS * 32 - 1) / (FW_WARPS * 32);   // 3 .. 8 for n in (256, 1024]
  ppl = ppl < 3 ? 3 : (ppl > 6 ? FW_COOP_PPL : ppl);       // the instantiated sizes
  for (int k = n + tid; k < ppl * FW_WARPS * 32; k += FW_WARPS * 32) { sx[k] = 0.f; sy[k] = 0.f; sz[k] = 0.f; }
  __syncthreads();
  switch (ppl) {
    case 3: fps_coop_picks<3>(a, task, n, sx, sy, sz, s_cand); break;
    case 4: fps_coop_picks<4>(a, task, n, sx, sy, sz, s_cand); break;
    case 5: fps_coop_picks<5>(a, task, n, sx, sy, sz, s_cand); break;
    case 6: fps_coop_picks<6>(a, task, n, sx, sy, sz, s_cand); break;
    default: fps_coop_picks<FW_COOP_PPL>(a, task, n, sx, sy, sz, s_cand); break;
  }
}

// A task of up to FW_MAXN points (or a dead one: zeros) run by the calling warp in its own plane slice.
__device__ __forceinline__ void fps_warp_task(const FpsTaskArgs& a, int task, int n, bool live, float* s_p,
                                              unsigned short* s_sel_all) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = a.m;
  float* dst = a.out_pts ? a.out_pts + (size_t)task * a.out_stride : nullptr;
  int32_t* idst = a.out_idx ? a.out_idx + (size_t)task * m : nullptr;
  if (!live) {
    for (int i = lane; i < m; i += 32) {
      if (dst) { dst[i * 3 + 0] = 0.f; dst[i * 3 + 1] = 0.f; dst[i * 3 + 2] = 0.f; }
      if (idst) idst[i] = 0;
    }
    return;
  }
  if (n > FW_MAXN) return;  // a cooperative task, or fps_kernel's
  float* sx = s_p + (size_t)warp * FW_MAXN;
  float* sy = sx + FW_PLANE;
  float* sz = sy + FW_PLANE;
  unsigned short* sel = s_sel_all + warp * m;
  const float* src = a.pts + (size_t)task * a.task_stride;
  __syncwarp();
  for (int i = lane; i < 3 * n; i += 32) {
    const int k = i / 3, c = i - 3 * k;
    sx[c * FW_PLANE + k] = src[i];
  }
  // points per lane in steps of one up to 8 (the per-pick work is proportional to it), then 10 and 12
  const int ppl = n <= 128 ? 4 : (n + 31) >> 5;   // 4 .. FW_MAXN / 32
  for (int k = n + lane; k < 32 * ppl; k += 32) { sx[k] = 0.f; sy[k] = 0.f; sz[k] = 0.f; }
  __syncwarp();
  switch (ppl) {
    case 4: fps_warp_picks<4, true>(sx, sy, sz, n, m, lane, sel); break;
    case 5: fps_warp_picks<5, true>(sx, sy, sz, n, m, lane, sel); break;
    case 6: fps_warp_picks<6, true>(sx, sy, sz, n, m, lane, sel); break;
    case 7: fps_warp_picks<7, true>(sx, sy, sz, n, m, lane, sel); break;
    default: fps_warp_picks<FW_MAXN / 32, true>(sx, sy, sz, n, m, lane, sel); break;
  }
  static_assert(FW_MAXN / 32 == 8, "the switch above covers 4 .. 8 points per lane");
  __syncwarp();
  for (int i = lane; i < m; i += 32) {
    const int k = sel[i];
    if (idst) idst[i] = k;
    if (dst) { dst[i * 3 + 0] = sx[k]; dst[i * 3 + 1] = sy[k]; dst[i * 3 + 2] = sz[k]; }
  }
}

// Static assignment: CTA b owns tasks 4b .. 4b+3 (igi_fps: no scratch for a schedule).
__global__ void __launch_bounds__(FW_WARPS * 32, 8) fps_warp_kernel(FpsTaskArgs a) {
  // dynamic smem: [3][FW_PLANE] f32 point planes (a warp task uses its own FW_MAXN slice of every plane),
  // then FW_WARPS x m selected indices (u16)
  extern __shared__ __align__(16) unsigned char fw_smem[];
  float* s_p = reinterpret_cast<float*>(fw_smem);
  unsigned short* s_sel_all = reinterpret_cast<unsigned short*>(fw_smem + sizeof(float) * 3 * FW_PLANE);
  __shared__ int2 s_cand[2][FW_WARPS];
  const int warp = threadIdx.x >> 5;
  // long tasks of this CTA's group: all warps together, one after the other
  for (int t = 0; t < FW_WARPS; ++t) {
    const int task = blockIdx.x * FW_WARPS + t;
    if (task >= a.n_tasks) break;
    const int n = a.count ? a.count[(size_t)task * a.count_stride] : a.n_fixed;
    if (n <= FW_MAXN || n > FW_COOP_MAXN) continue;
    if (a.any && a.any[(size_t)task * a.count_stride] == 0) continue;   // written as zeros by its warp below
    fps_coop_task(a, task, n, s_p, s_cand);
  }
  __syncthreads();
  const int task = blockIdx.x * FW_WARPS + warp;
  if (task >= a.n_tasks) return;
  const int n = a.count ? a.count[(size_t)task * a.count_stride] : a.n_fixed;
  const bool live = (a.any ? a.any[(size_t)task * a.count_stride] != 0 : true) && n > 0;
  fps_warp_task(a, task, n, live, s_p, s_sel_all);
}

// ---- cluster FPS: tasks of FW_COOP_MAXN+1 .. FC_MAXN points ------------------------------------------
// One thread-block cluster (FC_CL CTAs x FC_BLOCK threads = 1024 threads on FC_CL SMs) per task.  Every
// thread keeps its PPL points, their running minima and tie keys in registers for the whole task (no
// shared-memory point arrays at all); per pick each warp reduces its lanes with two REDUX instructions and
// writes its candidate - distance, tie key and the candidate's coordinates - into the candidate table of
// EVERY CTA of the cluster through distributed shared memory, one cluster barrier makes the 32 candidates
// visible, and every warp reduces them again with REDUX, so all 1024 threads leave the barrier knowing the
// picked point's index and coordinates without another memory round trip.  Same arithmetic and selection rule
// as the other FPS kernels / oracle/fps.py.
constexpr int FC_CL = 4, FC_BLOCK = 256, FC_THREADS = FC_CL * FC_BLOCK, FC_MAXPPL = 8, FC_MAXN = FC_MAXPPL * FC_THREADS;
// Clusters shorten the chain of dependent picks of ONE task (5x at 8-64 tasks of 5184 points) but finish fewer tasks
// per second than 128-thread CTAs once the machine is full of tasks (measured cross-over between 512 and 4096 tasks).
constexpr int FC_TASK_LIMIT = 512;
constexpr int FC_CANDS = FC_THREADS / 32;   // 32 warps in the cluster = one candidate per lane
static_assert(FC_CANDS == 32, "the second reduction maps one candidate to each lane");
struct __align__(8) FcCand { int hi; uint32_t lo; float x, y, z; uint32_t pad; };

template <int PPL>
__device__ __forceinline__ void fps_cluster_picks(cg::cluster_group& cluster, const float* __restrict__ src, int n, int m,
                                                  float* dst, int32_t* idst, FcCand (*s_cand)[FC_CANDS]) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank();
  const int gtid = rank * FC_BLOCK + tid;
  int lg = 31 - __clz(n);
  if (lg > 9) lg = 9;
  const uint32_t bmask = (1u << lg) - 1u;
  float px[PPL], py[PPL], pz[PPL], temp[PPL];
  uint32_t lokey[PPL];
#pragma unroll
  for (int i = 0; i < PPL; ++i) {
    const int k = gtid + FC_THREADS * i;
    px[i] = 0.f; py[i] = 0.f; pz[i] = 0.f; temp[i] = -1.0f; lokey[i] = 0u;
    if (k < n) {
      px[i] = src[(size_t)k * 3 + 0]; py[i] = src[(size_t)k * 3 + 1]; pz[i] = src[(size_t)k * 3 + 2];
      const float mag = fps_sumsq(px[i], py[i], pz[i]);
      if (mag > 1e-3f) { temp[i] = 1e10f; lokey[i] = fps_tie_key(k, lg, bmask); }
    }
  }
  const float x0 = src[0], y0 = src[1], z0 = src[2];
  float x1 = x0, y1 = y0, z1 = z0;   // coordinates of the point picked last
  int old = 0;
  if (gtid == 0) {
    if (idst) idst[0] = 0;
    if (dst) { dst[0] = x0; dst[1] = y0; dst[2] = z0; }
  }
  int j = 1;
  for (; j < m; ++j) {
    const float best = fps_update_reg<PPL>(px, py, pz, x1, y1, z1, temp);
    const int wb = __reduce_max_sync(0xffffffffu, __float_as_int(best));   // floats >= 0 order as ints; -1 < 0
    const uint32_t wlo = __reduce_max_sync(0xffffffffu, fps_tie<PPL>(temp, lokey, wb));
    // coordinates of the warp's candidate: the (unique) slot whose tie key is wlo
    float cx = 0.f, cy = 0.f, cz = 0.f;
    bool mine = false;
#pragma unroll
    for (int i = 0; i < PPL; ++i)
      if (wlo != 0u && lokey[i] == wlo) { cx = px[i]; cy = py[i]; cz = pz[i]; mine = true; }
    const unsigned owners = __ballot_sync(0xffffffffu, mine);
    const int owner = owners ? __ffs(owners) - 1 : 0;
    cx = __shfl_sync(0xffffffffu, cx, owner); cy = __shfl_sync(0xffffffffu, cy, owner); cz = __shfl_sync(0xffffffffu, cz, owner);
    if (lane < FC_CL) {   // lane r stores the warp's candidate into CTA r's table
      FcCand* table = cluster.map_shared_rank(&s_cand[j & 1][0], lane);
      FcCand c;
      c.hi = wb; c.lo = wlo; c.x = cx; c.y = cy; c.z = cz; c.pad = 0u;
      table[rank * (FC_BLOCK / 32) + warp] = c;
    }
    cluster.sync();
    const FcCand c = s_cand[j & 1][lane];
    const int fb = __reduce_max_sync(0xffffffffu, c.hi);
    const uint32_t flo = __reduce_max_sync(0xffffffffu, c.hi == fb ? c.lo : 0u);
    const unsigned winners = __ballot_sync(0xffffffffu, c.hi == fb && c.lo == flo);
    const int win = __ffs(winners) - 1;
    if (fb < 0) { old = 0; x1 = x0; y1 = y0; z1 = z0; }
    else {
      old = (int)((~flo) & 0xffffu);
      x1 = __shfl_sync(0xffffffffu, c.x, win); y1 = __shfl_sync(0xffffffffu, c.y, win); z1 = __shfl_sync(0xffffffffu, c.z, win);
    }
    if (gtid == 0) {
      if (idst) idst[j] = old;
      if (dst) { dst[j * 3 + 0] = x1; dst[j * 3 + 1] = y1; dst[j * 3 + 2] = z1; }
    }
    if (fb <= 0) { ++j; break; }   // cluster-uniform: every thread reduced the same table
  }
  for (int i = j + gtid; i < m; i += FC_THREADS) {
    if (idst) idst[i] = old;
    if (dst) { dst[i * 3 + 0] = x1; dst[i * 3 + 1] = y1; dst[i * 3 + 2] = z1; }
  }
}

__global__ void __cluster_dims__(FC_CL, 1, 1) __launch_bounds__(FC_BLOCK)
    fps_cluster_kernel(FpsTaskArgs a, int min_n, const int32_t* sched, const int32_t* order) {
  __shared__ FcCand s_cand[2][FC_CANDS];
  cg::cluster_group cluster = cg::this_cluster();
  const int cid = blockIdx.x / FC_CL, n_clusters = gridDim.x / FC_CL;
  const int gtid = (int)cluster.block_rank() * FC_BLOCK + threadIdx.x;
  const int m = a.m;
  // With a schedule (igi_fps_balanced) only the tasks fps_order_kernel listed as big are visited.
  const int n_iter = sched ? sched[4] : a.n_tasks;
  if (sched && n_iter > FC_TASK_LIMIT) return;   // a long list is better served by the one-CTA kernel (uniform exit)
  for (int it = cid; it < n_iter; it += n_clusters) {
    const int task = sched ? order[a.n_tasks - 1 - it] : it;
    const int n = a.count ? a.count[(size_t)task * a.count_stride] : a.n_fixed;
    if (n < min_n || n > FC_MAXN) continue;   // the resident kernels' / fps_kernel's
    const bool live = (a.any ? a.any[(size_t)task * a.count_stride] != 0 : true) && n > 0;
    float* dst = a.out_pts ? a.out_pts + (size_t)task * a.out_stride : nullptr;
    int32_t* idst = a.out_idx ? a.out_idx + (size_t)task * m : nullptr;
    if (!live) {
      for (int i = gtid; i < m; i += FC_THREADS) {
        if (dst) { dst[i * 3 + 0] = 0.f; dst[i * 3 + 1] = 0.f; dst[i * 3 + 2] = 0.f; }
        if (idst) idst[i] = 0;
      }
      continue;
    }
    cluster.sync();   // the previous task's last table has been read by every CTA
    const float* src = a.pts + (size_t)task * a.task_stride;
    const int ppl = (n + FC_THREADS - 1) / FC_THREADS;
    if (ppl <= 2) fps_cluster_picks<2>(cluster, src, n, m, dst, idst, s_cand);
    else if (ppl <= 4) fps_cluster_picks<4>(cluster, src, n, m, dst, idst, s_cand);
    else if (ppl <= 6) fps_cluster_picks<6>(cluster, src, n, m, dst, idst, s_cand);
    else fps_cluster_picks<FC_MAXPPL>(cluster, src, n, m, dst, idst, s_cand);
  }
  cluster.sync();   // no CTA leaves while a peer may still write into its table
}

// ---- size-ordered schedule (igi_fps_balanced) ---------------------------------------------------------
// The cost of a task grows with the square of its point count (n picks before the early exit x n/32
// points per lane), and the counts of neighbouring envs differ by up to 6x, so a static assignment
// leaves most SMs waiting for the few that drew long tasks.  fps_order_kernel counting-sorts the task
// ids by descending point count (64 buckets of 16 points; order inside a bucket is whatever the atomics
// give - it only affects the schedule, never a result); fps_sorted_kernel's persistent CTAs then pull
// from that list, longest first: whole CTAs for the cooperative sizes, then single warps.
// sched: [0] n_coop  [1] n_listed  [2] cooperative cursor  [3] warp cursor  [4] n_big (> FW_COOP_MAXN)
constexpr int FO_BUCKETS = FW_COOP_MAXN / 16;        // 64
constexpr int FO_COOP_FIRST = FW_MAXN / 16;          // 24: buckets >= this are cooperative sizes
static_assert(FW_MAXN % 16 == 0 && FW_COOP_MAXN % 16 == 0, "bucket edges must match the size classes");

__global__ void __launch_bounds__(1024) fps_order_kernel(const int32_t* count, const int32_t* any, int64_t count_stride,
                                                         int n_tasks, int32_t* sched, int32_t* order) {
  __shared__ int s_hist[FO_BUCKETS + 1], s_base[FO_BUCKETS + 1];
  const int tid = threadIdx.x;
  if (tid <= FO_BUCKETS) s_hist[tid] = 0;
  __syncthreads();
  auto bucket_of = [&](int task) {
    const int n = count[(size_t)task * count_stride];
    const bool live = (any ? any[(size_t)task * count_stride] != 0 : true) && n > 0;
    if (!live) return 0;
    if (n > FW_COOP_MAXN) return FO_BUCKETS;
    return (n - 1) >> 4;
  };
  for (int t = tid; t < n_tasks; t += 1024) atomicAdd(&s_hist[bucket_of(t)], 1);
  __syncthreads();
  if (tid == 0) {
    int run = 0, n_coop = 0;
    for (int b = FO_BUCKETS - 1; b >= 0; --b) {
      s_base[b] = run;
      run += s_hist[b];
      if (b == FO_COOP_FIRST) n_coop = run;
    }
    sched[0] = n_coop; sched[1] = run; sched[2] = 0; sched[3] = 0; sched[4] = s_hist[FO_BUCKETS];
    s_base[FO_BUCKETS] = 0;
  }
  __syncthreads();
  if (tid <= FO_BUCKETS) s_hist[tid] = 0;
  __syncthreads();
  for (int t = tid; t < n_tasks; t += 1024) {
    const int b = bucket_of(t);
    const int pos = s_base[b] + atomicAdd(&s_hist[b], 1);
    order[b == FO_BUCKETS ? n_tasks - 1 - pos : pos] = t;   // the big ones are listed from the back
  }
}

__global__ void __launch_bounds__(FW_WARPS * 32, 8) fps_sorted_kernel(FpsTaskArgs a, int32_t* sched, const int32_t* order) {
  extern __shared__ __align__(16) unsigned char fw_smem[];
  float* s_p = reinterpret_cast<float*>(fw_smem);
  unsigned short* s_sel_all = reinterpret_cast<unsigned short*>(fw_smem + sizeof(float) * 3 * FW_PLANE);
  __shared__ int2 s_cand[2][FW_WARPS];
  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_coop = sched[0], n_listed = sched[1];
  for (;;) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(&sched[2], 1);
    __syncthreads();
    const int i = s_item;
    if (i >= n_coop) break;
    const int task = order[i];
    fps_coop_task(a, task, a.count[(size_t)task * a.count_stride], s_p, s_cand);
  }
  for (;;) {
    int i = 0;
    if (lane == 0) i = n_coop + atomicAdd(&sched[3], 1);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_listed) break;
    const int task = order[i];
    const int n = a.count[(size_t)task * a.count_stride];
    const bool live = (a.any ? a.any[(size_t)task * a.count_stride] != 0 : true) && n > 0;
    fps_warp_task(a, task, n, live, s_p, s_sel_all);
  }
}

}  // namespace

extern "C" int igi_pcl_compact(const float* depth, const int32_t* seg, const int32_t* seg_ids, int n_classes,
                               const float* uvx, const float* uvy, const float* uvz, const float* ext,
                               const float* e2g_inv, int n_envs, int H, int W, float depth_max,
                               const float* box, float* out_pts, int32_t* out_count, int32_t* out_any,
                               void* stream) {
  IGI_REQUIRE(depth && uvx && uvy && uvz && ext && e2g_inv && out_pts && out_count && out_any,
              "igi_pcl_compact: null pointer");
  IGI_REQUIRE(n_envs >= 0 && H > 0 && W > 0, "igi_pcl_compact: bad dims");
  IGI_REQUIRE(n_classes >= 1 && n_classes <= kMaxClasses, "igi_pcl_compact: n_classes must be 1..4");
  IGI_REQUIRE(seg == nullptr || seg_ids != nullptr, "igi_pcl_compact: seg given without seg_ids");
  if (n_envs == 0) return IGI_OK;
  CompactParams p{};
  p.depth = depth; p.seg = seg; p.uvx = uvx; p.uvy = uvy; p.uvz = uvz; p.ext = ext; p.e2g_inv = e2g_inv;
  p.out_pts = out_pts; p.out_count = out_count; p.out_any = out_any;
  p.n_envs = n_envs; p.H = H; p.W = W; p.n_classes = n_classes;
  for (int c = 0; c < n_classes; ++c) p.seg_ids[c] = seg_ids ? seg_ids[c] : 0;
  p.depth_max = depth_max;
  p.has_box = box != nullptr;
  if (box) for (int i = 0; i < 6; ++i) p.box[i] = box[i];
  const size_t npix = (size_t)H * W;
  const size_t smem = 2 * ((npix * 4 + 15) & ~(size_t)15) + npix * 2 + 16;   // depth, seg, u16 pixel list
  IGI_REQUIRE(smem <= 200 * 1024 && npix < 65536, "igi_pcl_compact: image too large for one CTA (%d x %d)", H, W);
  p.use_bulk = ((npix * 4) % 16 == 0) && ((uintptr_t)depth % 16 == 0) && (!seg || (uintptr_t)seg % 16 == 0);
  {   // the opt-in shared-memory size is a per-DEVICE function attribute
    static bool attr_set[64] = {false};
    int dev = 0;
    IGI_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      IGI_CUDA(cudaFuncSetAttribute(pcl_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  pcl_compact_kernel<<<n_envs, kCompactBlock, smem, (cudaStream_t)stream>>>(p);
  IGI_CHECK_LAUNCH("pcl_compact_kernel");
  return IGI_OK;
}

extern "C" int igi_pcl_sample_gather(const float* pts, const int32_t* count, const int32_t* any,
                                     int n_classes, int cls, int cap, const uint32_t* raw, int n_envs, int m,
                                     float* out, int64_t out_stride, int32_t* out_idx, int32_t* out_consumed,
                                     int32_t* scratch_offsets, void* stream) {
  IGI_REQUIRE(pts && count && any && raw && out && scratch_offsets, "igi_pcl_sample_gather: null pointer");
  IGI_REQUIRE(n_classes >= 1 && cls >= 0 && cls < n_classes && cap > 0 && m > 0 && n_envs >= 0,
              "igi_pcl_sample_gather: bad dims");
  IGI_REQUIRE(out_stride >= (int64_t)m * 3, "igi_pcl_sample_gather: out_stride < 3*m");
  if (n_envs == 0) return IGI_OK;
  int32_t* d_offsets = scratch_offsets;
  cudaStream_t s = (cudaStream_t)stream;
  pcl_offsets_kernel<<<1, 1024, 0, s>>>(any, n_classes, cls, n_envs, m, d_offsets, out_consumed);
  IGI_CHECK_LAUNCH("pcl_offsets_kernel");
  pcl_gather_kernel<<<n_envs, 128, 0, s>>>(pts, count, any, n_classes, cls, cap, raw, d_offsets, m, out,
                                            out_stride, out_idx);
  IGI_CHECK_LAUNCH("pcl_gather_kernel");
  return IGI_OK;
}

// use_cluster = 0 (flag IGI_FPS_NO_CLUSTER of the call): big tasks go to the one-CTA fps_kernel instead (A/B, tests)
static int fps_block_launch(const FpsTaskArgs& a, int64_t nmax, int min_n, const int32_t* sched, const int32_t* order,
                            cudaStream_t st, int use_cluster) {
  int dev0 = 0, sms0 = 148;
  cudaGetDevice(&dev0);
  cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
  // Tasks the resident kernels leave (more than FW_COOP_MAXN points) go to thread-block clusters up to FC_MAXN
  // points; only larger ones (or every task, when m is too large for the resident kernels: min_n == 0) to fps_kernel.
  const bool all_big = a.count == nullptr && a.n_fixed > FW_COOP_MAXN;   // fixed-size call above the resident sizes
  // With a schedule the number of big tasks is only known on the device: both kernels are launched and read
  // sched[4]; without one (igi_fps) every task may be big and the host decides on n_tasks.
  const bool cluster_ok = use_cluster && (min_n > 0 || all_big) && (sched != nullptr || a.n_tasks <= FC_TASK_LIMIT);
  if (cluster_ok) {
    int clusters = sms0 / FC_CL * 3;
    if (clusters > a.n_tasks) clusters = a.n_tasks;
    if (clusters < 1) clusters = 1;
    fps_cluster_kernel<<<clusters * FC_CL, FC_BLOCK, 0, st>>>(a, min_n, sched, order);
    IGI_CHECK_LAUNCH("fps_cluster_kernel");
    if (!sched) {
      if (nmax <= FC_MAXN) return IGI_OK;
      min_n = FC_MAXN + 1;
    }
  }
  const int cluster_limit = cluster_ok && sched ? FC_TASK_LIMIT : 0;
  // worst-case dynamic smem: nmax points (count is on the device)
  const size_t smem = (size_t)((nmax + 3) & ~3) * 16;
  IGI_REQUIRE(smem <= 220 * 1024, "igi_fps: %lld points per task exceed shared memory", (long long)nmax);
  int dev = 0, sms = 148;
  IGI_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static size_t attr_smem[64] = {0};   // per device
  if (dev < 0 || dev >= 64 || smem > attr_smem[dev]) {
    IGI_CUDA(cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (dev >= 0 && dev < 64) attr_smem[dev] = smem;
  }
  const int per_sm = smem > 100 * 1024 ? 1 : (smem > 64 * 1024 ? 2 : 3);
  const int grid = a.n_tasks < sms * per_sm ? a.n_tasks : sms * per_sm;
  fps_kernel<<<grid, kFpsBlock, smem, st>>>(a.pts, a.task_stride, a.count, a.any, a.count_stride, a.n_fixed, a.n_tasks,
                                            a.m, a.out_pts, a.out_stride, a.out_idx, min_n, sched, order, cluster_limit, FC_MAXN);
  IGI_CHECK_LAUNCH("fps_kernel");
  return IGI_OK;
}

static int fps_warp_attr() {
  static bool done[64] = {false};   // per device
  int dev = 0;
  IGI_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    IGI_CUDA(cudaFuncSetAttribute(fps_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
    IGI_CUDA(cudaFuncSetAttribute(fps_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 56 * 1024));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return IGI_OK;
}

extern "C" int igi_fps(const float* pts, int64_t task_stride, const int32_t* count, const int32_t* any,
                       int64_t count_stride, int n_fixed, int n_tasks, int m, float* out_pts,
                       int64_t out_stride, int32_t* out_idx, int flags, void* stream) {
  IGI_REQUIRE(pts && (out_pts || out_idx), "igi_fps: null pointer");
  IGI_REQUIRE((flags & ~IGI_FPS_NO_CLUSTER) == 0, "igi_fps: unknown flag bits");
  IGI_REQUIRE(n_tasks >= 0 && m > 0, "igi_fps: bad dims");
  IGI_REQUIRE(count || n_fixed > 0, "igi_fps: need count or n_fixed");
  IGI_REQUIRE(!out_pts || out_stride >= (int64_t)m * 3, "igi_fps: out_stride < 3*m");
  if (n_tasks == 0) return IGI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nmax = count ? task_stride / 3 : n_fixed;
  IGI_REQUIRE(nmax <= 0xffff, "igi_fps: at most 65535 points per task");
  FpsTaskArgs a{pts, task_stride, count, any, count_stride, n_fixed, n_tasks, m, out_pts, out_stride, out_idx};
  const size_t warp_smem = sizeof(float) * 3 * FW_PLANE + (size_t)FW_WARPS * m * 2;
  const bool warp_ok = warp_smem <= 28 * 1024;  // keeps 8 CTAs (32 task warps) per SM
  const bool need_block = !warp_ok || nmax > FW_COOP_MAXN;
  const bool need_warp = warp_ok && (count != nullptr || n_fixed <= FW_COOP_MAXN);
  if (need_warp) {
    if (int rc = fps_warp_attr()) return rc;
    fps_warp_kernel<<<(n_tasks + FW_WARPS - 1) / FW_WARPS, FW_WARPS * 32, warp_smem, st>>>(a);
    IGI_CHECK_LAUNCH("fps_warp_kernel");
  }
  if (need_block) return fps_block_launch(a, nmax, need_warp ? FW_COOP_MAXN + 1 : 0, nullptr, nullptr, st, !(flags & IGI_FPS_NO_CLUSTER));
  return IGI_OK;
}

extern "C" int igi_fps_balanced(const float* pts, int64_t task_stride, const int32_t* count, const int32_t* any,
                                int64_t count_stride, int n_tasks, int m, float* out_pts, int64_t out_stride,
                                int32_t* out_idx, int32_t* scratch, int flags, void* stream) {
  IGI_REQUIRE(pts && count && scratch && (out_pts || out_idx), "igi_fps_balanced: null pointer");
  IGI_REQUIRE((flags & ~IGI_FPS_NO_CLUSTER) == 0, "igi_fps_balanced: unknown flag bits");
  IGI_REQUIRE(n_tasks >= 0 && m > 0, "igi_fps_balanced: bad dims");
  IGI_REQUIRE(!out_pts || out_stride >= (int64_t)m * 3, "igi_fps_balanced: out_stride < 3*m");
  if (n_tasks == 0) return IGI_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nmax = task_stride / 3;
  IGI_REQUIRE(nmax <= 0xffff, "igi_fps_balanced: at most 65535 points per task");
  const size_t warp_smem = sizeof(float) * 3 * FW_PLANE + (size_t)FW_WARPS * m * 2;
  if (warp_smem > 28 * 1024)   // m too large for the resident kernel: the static path handles it
    return igi_fps(pts, task_stride, count, any, count_stride, 0, n_tasks, m, out_pts, out_stride, out_idx, flags, stream);
  FpsTaskArgs a{pts, task_stride, count, any, count_stride, 0, n_tasks, m, out_pts, out_stride, out_idx};
  int32_t* sched = scratch;
  int32_t* order = scratch + 8;
  fps_order_kernel<<<1, 1024, 0, st>>>(count, any, count_stride, n_tasks, sched, order);
  IGI_CHECK_LAUNCH("fps_order_kernel");
  if (int rc = fps_warp_attr()) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int want = (n_tasks + FW_WARPS - 1) / FW_WARPS;
  fps_sorted_kernel<<<want < sms * 8 ? want : sms * 8, FW_WARPS * 32, warp_smem, st>>>(a, sched, order);
  IGI_CHECK_LAUNCH("fps_sorted_kernel");
  if (nmax > FW_COOP_MAXN) return fps_block_launch(a, nmax, FW_COOP_MAXN + 1, sched, order, st, !(flags & IGI_FPS_NO_CLUSTER));
  return IGI_OK;
}
