// SURVEY 8f "next" rows on either side of the hot path: the image observations of the external
// camera, the RNG-defined noise stages (counter-based Philox so results do not depend on how envs
// are sharded over GPUs), and the first consumer of the observation buffer, the student's
// running mean/std normaliser.
//
//   S1 cam_image_obs_kernel   depth noise + clip + normalise -> image_buf ; seg copy + object-pixel
//                             flips -> seg_buf                     (factory_task_insertion.py:925-943,
//                                                                   factory_utils.py:23-37,55-72)
//   S2 pcl_noise_kernel       PointCloudAugmentations.random_noise (factory_utils.py:93-100) under
//                             the per-env pcl_noise mask            (factory_task_insertion.py:966-969)
//   S3 rms_moments_kernel / rms_normalize_kernel   RunningMeanStd.forward (algo/models/
//                             running_mean_std.py:60-93) as called by process_obs (ext_adapt.py:405)
//   S4 seg_valid_mask_kernel  process_obs seg / img masking        (ext_adapt.py:391-396)
//   S5 queue_push_kernel      history queues: q[:,1:] = q[:,:-1]; q[:,0] = x  (:1046-1056, :512-513)
#include "igi_common.cuh"
#include "../../include/igi_b200.h"

namespace {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (idx_lo, idx_hi, step, stream), key = seed ----------
struct U4 { uint32_t x, y, z, w; };
__device__ __forceinline__ U4 philox(uint64_t idx, uint32_t step, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = step, c3 = stream;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}
// 24-bit uniforms: [0,1) for comparisons and offsets, (0,1] for the logarithm of Box-Muller
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float u01_open(uint32_t r) { return (float)((r >> 8) + 1u) * (1.0f / 16777216.0f); }
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float r = sqrtf(-2.0f * logf(u01_open(a)));
  float s, c;
  sincospif(2.0f * u01(b), &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// ---- S1 ------------------------------------------------------------------------------------------
struct CamObsArgs {
  const float* depth;        // (n, npix)
  const int32_t* seg;        // (n, npix) or null
  const uint8_t* update;     // (n) image_buf rows to refresh (update_freq & update_delay), null = all
  const uint8_t* update_seg; // (n) seg_buf rows to refresh, null = all
  const uint8_t* seg_noise;  // (n) rows whose seg gets flips (ANDed with update_seg here), null = none
  float* image_buf;          // (n, npix) or null
  int32_t* seg_buf;          // (n, npix) or null
  int n_envs, npix;
  long long env0;            // global id of env 0 of this shard
  float noise_scale;         // dis_noise * 2
  float far_clip, near_clip, inv_range_den;  // inv_range_den = far - near (the divisor; torch CPU divides, it does not multiply by 1/x)
  float flip_prob;
  uint64_t seed;
  uint32_t step;
};

// One thread = 4 consecutive pixels of one env (npix % 4 == 0): one Philox block per stream.
__global__ void __launch_bounds__(256) cam_image_obs_kernel(CamObsArgs a) {
  const int q = a.npix >> 2;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.n_envs * q) return;
  const int e = (int)(t / q), j = (int)(t - (long long)e * q);
  const uint64_t gidx = (uint64_t)(a.env0 + e) * (uint64_t)q + (uint64_t)j;
  const size_t off = (size_t)e * a.npix + 4 * (size_t)j;
  if (a.image_buf && (!a.update || a.update[e])) {
    const float4 d4 = __ldcs(reinterpret_cast<const float4*>(a.depth + off));
    const U4 r = philox(gidx, a.step, 0u, a.seed);
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // noise = dis_noise*2 * (rand - 0.5); depth += noise; clip(-far, -near); *-1; (d - near)/(far - near)
      const float nz = __fmul_rn(a.noise_scale, __fsub_rn(u01(rr[k]), 0.5f));
      float v = __fadd_rn(d[k], nz);
      v = fminf(fmaxf(v, -a.far_clip), -a.near_clip);   // NaN propagates like torch.clip: keep it
      if (d[k] != d[k]) v = d[k];
      v = -v;
      o[k] = __fdiv_rn(__fsub_rn(v, a.near_clip), a.inv_range_den);
    }
    __stcs(reinterpret_cast<float4*>(a.image_buf + off), make_float4(o[0], o[1], o[2], o[3]));
  }
  if (a.seg_buf && a.seg && (!a.update_seg || a.update_seg[e])) {
    int4 s4 = __ldcs(reinterpret_cast<const int4*>(a.seg + off));
    if (a.seg_noise && a.seg_noise[e]) {
      const U4 r = philox(gidx, a.step, 1u, a.seed);
      // seg[(seg > 0) & (rand < flip_prob)] = 0
      if (s4.x > 0 && u01(r.x) < a.flip_prob) s4.x = 0;
      if (s4.y > 0 && u01(r.y) < a.flip_prob) s4.y = 0;
      if (s4.z > 0 && u01(r.z) < a.flip_prob) s4.z = 0;
      if (s4.w > 0 && u01(r.w) < a.flip_prob) s4.w = 0;
    }
    __stcs(reinterpret_cast<int4*>(a.seg_buf + off), s4);
  }
}

// ---- S2 ------------------------------------------------------------------------------------------
// One thread per point: 3 clamped normals (sigma), one Bernoulli(noise_prob) gate, + the env's constant offset.
__global__ void __launch_bounds__(256) pcl_noise_kernel(float* pts, int64_t env_stride, int n_envs, int n_pts,
                                                         const uint8_t* mask, const float* pcl_noise /*(n,3)*/,
                                                         long long env0, float sigma, float clip, float const_noise,
                                                         float noise_prob, uint64_t seed, uint32_t step) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)n_envs * n_pts) return;
  const int e = (int)(t / n_pts), i = (int)(t - (long long)e * n_pts);
  if (mask && !mask[e]) return;
  const uint64_t gidx = (uint64_t)(env0 + e) * (uint64_t)n_pts + (uint64_t)i;
  const U4 r0 = philox(gidx, step, 2u, seed), r1 = philox(gidx, step, 3u, seed);
  float z[4];
  box_muller(r0.x, r0.y, z[0], z[1]);
  box_muller(r0.z, r0.w, z[2], z[3]);
  const float gate = u01(r1.x) < noise_prob ? 1.0f : 0.0f;
  float* p = pts + (size_t)e * env_stride + (size_t)i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float pw = fminf(fmaxf(__fmul_rn(z[c], sigma), -clip), clip);
    const float cn = fminf(fmaxf(__fmul_rn(pcl_noise[(size_t)e * 3 + c], const_noise), -clip), clip);
    p[c] = __fadd_rn(__fadd_rn(p[c], __fmul_rn(pw, gate)), cn);
  }
}

// ---- S3 ------------------------------------------------------------------------------------------
constexpr int RMS_MAX_C = 256;   // <= RMS_BLOCK: one thread per channel folds the CTA partials
constexpr int RMS_GRID = 592;   // 4 CTAs per SM on 148 SMs
constexpr int RMS_BLOCK = 256;

// Per-channel sums of (x - pivot) and (x - pivot)^2 in f64, pivot = the first row.  The grid is a multiple
// of C CTAs, so the grid stride is a multiple of C and a thread only ever sees ONE channel while the warp
// reads consecutive floats.  Fixed grid, fixed per-thread element order and fixed-order folds (per CTA, then
// by the last CTA to finish), so the statistics are reproducible run to run.  The last CTA applies the
// reference's parallel-variance update to the f64 running statistics; the batch mean / variance are
// rounded to f32 first, as torch's input.mean() / input.var() return f32.
__global__ void __launch_bounds__(RMS_BLOCK) rms_moments_kernel(const float* __restrict__ x, long long rows, int C,
                                                                 double* running_mean, double* running_var,
                                                                 double* count, double* partial /*[grid][2C]*/,
                                                                 unsigned int* done) {
  __shared__ double s_s1[RMS_BLOCK], s_s2[RMS_BLOCK];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const long long total = rows * C;
  const long long g = (long long)blockIdx.x * RMS_BLOCK + tid;
  const long long stride = (long long)gridDim.x * RMS_BLOCK;   // % C == 0
  const int ch = (int)(g % C);
  const double pivot = (double)__ldg(x + ch);
  double s1 = 0.0, s2 = 0.0;
  long long i = g;
  for (; i + 3 * stride < total; i += 4 * stride) {   // four independent loads in flight
    const float v0 = __ldg(x + i), v1 = __ldg(x + i + stride), v2 = __ldg(x + i + 2 * stride), v3 = __ldg(x + i + 3 * stride);
    const double d0 = (double)v0 - pivot, d1 = (double)v1 - pivot, d2 = (double)v2 - pivot, d3 = (double)v3 - pivot;
    s1 += d0; s2 += d0 * d0;
    s1 += d1; s2 += d1 * d1;
    s1 += d2; s2 += d2 * d2;
    s1 += d3; s2 += d3 * d3;
  }
  for (; i < total; i += stride) {
    const double d = (double)__ldg(x + i) - pivot;
    s1 += d; s2 += d * d;
  }
  s_s1[tid] = s1;
  s_s2[tid] = s2;
  __syncthreads();
  if (tid < C) {
    // threads of this CTA whose channel is `tid`: t = first, first + C, ...
    const int base = (int)(((long long)blockIdx.x * RMS_BLOCK) % C);
    const int first = (tid - base + C) % C;
    double a1 = 0.0, a2 = 0.0;
    for (int t = first; t < RMS_BLOCK; t += C) { a1 += s_s1[t]; a2 += s_s2[t]; }
    partial[(size_t)blockIdx.x * 2 * C + 2 * tid] = a1;
    partial[(size_t)blockIdx.x * 2 * C + 2 * tid + 1] = a2;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid < C) {
    const int c = tid;
    double a1 = 0.0, a2 = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
      a1 += partial[(size_t)b * 2 * C + 2 * c];
      a2 += partial[(size_t)b * 2 * C + 2 * c + 1];
    }
    const double n = (double)rows;
    const double piv = (double)x[c];
    const double bmean = (double)(float)(piv + a1 / n);
    // unbiased variance (torch.var default); rows == 1 gives NaN like torch
    const double bvar = (double)(float)((a2 - a1 * a1 / n) / (n - 1.0));
    // _update_mean_var_count_from_moments (running_mean_std.py:47-57)
    const double mean = running_mean[c], var = running_var[c], cnt = *count;
    const double delta = bmean - mean;
    const double tot = cnt + n;
    const double new_mean = mean + delta * n / tot;
    const double M2 = var * cnt + bvar * n + delta * delta * cnt * n / tot;
    running_mean[c] = new_mean;
    running_var[c] = M2 / tot;
  }
  __syncthreads();
  if (tid == 0) {
    *count = *count + (double)rows;
    *done = 0u;   // re-armed for the next call
  }
}

// y = clamp((x - mean.float()) / sqrt(var.float() + eps), -5, 5)   (running_mean_std.py:82-92); the other two
// modes of the module: norm_only (no centring, no clamp) and unnorm (clamp first, then scale back).
// y may alias x (element-wise, each element is read before it is written by the same thread).
__global__ void __launch_bounds__(256) rms_normalize_kernel(const float* x, long long total, int C,
                                                             const double* __restrict__ running_mean,
                                                             const double* __restrict__ running_var, float eps,
                                                             int mode /*0 norm, 1 norm_only, 2 unnorm*/, float* y) {
  __shared__ float s_mean[RMS_MAX_C], s_den[RMS_MAX_C];
  if (threadIdx.x < C) {
    s_mean[threadIdx.x] = (float)running_mean[threadIdx.x];
    s_den[threadIdx.x] = __fsqrt_rn(__fadd_rn((float)running_var[threadIdx.x], eps));
  }
  __syncthreads();
  auto one = [&](float v, int c) {
    if (mode == 0) return fminf(fmaxf(__fdiv_rn(__fsub_rn(v, s_mean[c]), s_den[c]), -5.0f), 5.0f);
    if (mode == 1) return __fdiv_rn(v, s_den[c]);
    return __fadd_rn(__fmul_rn(s_den[c], fminf(fmaxf(v, -5.0f), 5.0f)), s_mean[c]);
  };
  // float4 chunks; the channel of element i is i % C
  const long long nvec = total >> 2;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const float4 in = reinterpret_cast<const float4*>(x)[v];
    int c = (int)((v * 4) % C);
    float4 o;
    o.x = one(in.x, c); c = c + 1 == C ? 0 : c + 1;
    o.y = one(in.y, c); c = c + 1 == C ? 0 : c + 1;
    o.z = one(in.z, c); c = c + 1 == C ? 0 : c + 1;
    o.w = one(in.w, c);
    __stcs(reinterpret_cast<float4*>(y) + v, o);
  }
  if (blockIdx.x == 0)
    for (long long i = (nvec << 2) + threadIdx.x; i < total; i += blockDim.x) y[i] = one(x[i], (int)(i % C));
}

// ---- S4 ------------------------------------------------------------------------------------------
// valid = (seg == obj) | (seg == socket); seg = distinct ? seg * valid : valid; img = img * valid
__global__ void __launch_bounds__(256) seg_valid_mask_kernel(const float* __restrict__ seg, const float* __restrict__ img,
                                                              long long n4, float obj_id, float socket_id, int distinct,
                                                              float* __restrict__ seg_out, float* __restrict__ img_out) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n4; v += (long long)gridDim.x * blockDim.x) {
    const float4 s = __ldcs(reinterpret_cast<const float4*>(seg) + v);
    const float m0 = (s.x == obj_id || s.x == socket_id) ? 1.0f : 0.0f, m1 = (s.y == obj_id || s.y == socket_id) ? 1.0f : 0.0f,
                m2 = (s.z == obj_id || s.z == socket_id) ? 1.0f : 0.0f, m3 = (s.w == obj_id || s.w == socket_id) ? 1.0f : 0.0f;
    const float4 so = distinct ? make_float4(__fmul_rn(s.x, m0), __fmul_rn(s.y, m1), __fmul_rn(s.z, m2), __fmul_rn(s.w, m3))
                               : make_float4(m0, m1, m2, m3);
    __stcs(reinterpret_cast<float4*>(seg_out) + v, so);
    if (img) {
      const float4 g = __ldcs(reinterpret_cast<const float4*>(img) + v);
      __stcs(reinterpret_cast<float4*>(img_out) + v,
             make_float4(__fmul_rn(g.x, m0), __fmul_rn(g.y, m1), __fmul_rn(g.z, m2), __fmul_rn(g.w, m3)));
    }
  }
}

// ---- S5 ------------------------------------------------------------------------------------------
// q (n, T, L): slots shift by one towards the back, slot 0 = x (f32, or i32 converted like the
// reference's float queue of the int32 seg_buf).  One CTA row-chunk walks its slots back to front, so
// every element is read before it is overwritten without the reference's clone().
template <typename TIn>
__global__ void __launch_bounds__(256) queue_push_kernel(float* q, const TIn* __restrict__ x, int64_t x_stride, int n,
                                                         int T, long long L4) {
  const long long per_env = L4;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)n * per_env;
       v += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(v / per_env);
    const long long j = v - (long long)e * per_env;
    float4* row = reinterpret_cast<float4*>(q) + (size_t)e * T * per_env + j;
    for (int t = T - 1; t > 0; --t) row[(size_t)t * per_env] = row[(size_t)(t - 1) * per_env];
    const TIn* src = x + (size_t)e * x_stride + 4 * j;
    row[0] = make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]);
  }
}

// dst[r, :] = src[r, :] for the rows whose flag equals `want` (rows of row16 uint4 each); rows that do not match cost
// one flag read.  grid = (chunks, rows).
__global__ void __launch_bounds__(256) copy_rows_where_kernel(uint4* __restrict__ dst, long long dst_stride16,
                                                              const uint4* __restrict__ src, long long src_stride16,
                                                              const uint8_t* __restrict__ flag, int want, long long rows,
                                                              long long row16) {
  for (long long r = blockIdx.y; r < rows; r += gridDim.y) {
    if ((flag[r] != 0) != (want != 0)) continue;
    const uint4* s = src + (size_t)r * src_stride16;
    uint4* d = dst + (size_t)r * dst_stride16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row16; i += (long long)gridDim.x * blockDim.x)
      __stcs(d + i, __ldcs(s + i));
  }
}

// Masks of one update_external_cam call (factory_task_insertion.py:896-989), one thread per env:
//   upd_seg = update_freq & seg_update_delay                      (:934-940)
//   restarted = socket_pending & (got_socket == 0)                (:981)
//   upd_pcl = (update_freq & update_delay) | restarted            (:988-989); got_socket[restarted] = 1
__global__ void __launch_bounds__(256) cam_masks_kernel(const uint8_t* __restrict__ freq, const uint8_t* __restrict__ delay,
                                                        const uint8_t* __restrict__ seg_delay, int32_t* __restrict__ got_socket,
                                                        int socket_pending, uint8_t* __restrict__ upd_seg,
                                                        uint8_t* __restrict__ upd_pcl, uint8_t* __restrict__ restarted, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const bool f = freq[e] != 0;
  const bool r = socket_pending == 2 || (socket_pending && got_socket[e] == 0);   // 2: every env counts as restarted
  if (upd_seg) upd_seg[e] = (f && seg_delay[e] != 0) ? 1 : 0;
  upd_pcl[e] = ((f && delay[e] != 0) || r) ? 1 : 0;
  if (restarted) restarted[e] = r ? 1 : 0;
  if (r) got_socket[e] = 1;
}

// pcl[e, :] = src[e, :] where update[e]  (factory_task_insertion.py:1027), then the history queue push
// q[:, 1:] = q[:, :-1]; q[:, 0] = pcl  (:1046-1048), in one pass over the rows.
__global__ void __launch_bounds__(256) pcl_assemble_kernel(const float4* __restrict__ src, int64_t src_stride4,
                                                           float4* __restrict__ pcl, int64_t pcl_stride4,
                                                           const uint8_t* __restrict__ update, float4* __restrict__ q, int n,
                                                           int T, long long L4) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)n * L4;
       v += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(v / L4);
    const long long j = v - (long long)e * L4;
    float4* prow = pcl + (size_t)e * pcl_stride4 + j;
    float4 val;
    if (update == nullptr || update[e] != 0) {
      val = __ldg(src + (size_t)e * src_stride4 + j);
      *prow = val;
    } else {
      val = *prow;
    }
    if (q) {
      float4* row = q + (size_t)e * T * L4 + j;
      for (int t = T - 1; t > 0; --t) row[(size_t)t * L4] = row[(size_t)(t - 1) * L4];
      row[0] = val;
    }
  }
}

int grid_for(long long work_items, int block, int cap = 148 * 16) {
  long long g = (work_items + block - 1) / block;
  if (g < 1) g = 1;
  return (int)(g < cap ? g : cap);
}

}  // namespace

extern "C" int igi_cam_image_obs(const float* depth, const int32_t* seg, const uint8_t* update, const uint8_t* update_seg,
                                 const uint8_t* seg_noise, int n_envs, int npix, long long env0, double dis_noise,
                                 double far_clip, double near_clip, float flip_prob, uint64_t seed, uint32_t step,
                                 float* image_buf, int32_t* seg_buf, void* stream) {
  IGI_REQUIRE(depth != nullptr, "igi_cam_image_obs: null depth");
  IGI_REQUIRE(image_buf != nullptr || seg_buf != nullptr, "igi_cam_image_obs: no output");
  IGI_REQUIRE(seg_buf == nullptr || seg != nullptr, "igi_cam_image_obs: seg_buf given without seg");
  IGI_REQUIRE(n_envs >= 0 && npix > 0 && npix % 4 == 0, "igi_cam_image_obs: npix must be a positive multiple of 4");
  IGI_REQUIRE(far_clip > near_clip, "igi_cam_image_obs: far_clip must exceed near_clip");
  if (n_envs == 0) return IGI_OK;
  CamObsArgs a{};
  a.depth = depth; a.seg = seg; a.update = update; a.update_seg = update_seg; a.seg_noise = seg_noise;
  a.image_buf = image_buf; a.seg_buf = seg_buf; a.n_envs = n_envs; a.npix = npix; a.env0 = env0;
  // python scalars of the reference: dis_noise * 2 and far - near are f64 products rounded to f32 by the tensor op
  a.noise_scale = (float)(dis_noise * 2.0);
  a.far_clip = (float)far_clip; a.near_clip = (float)near_clip;
  a.inv_range_den = (float)(far_clip - near_clip);
  a.flip_prob = flip_prob; a.seed = seed; a.step = step;
  const long long items = (long long)n_envs * (npix / 4);
  cam_image_obs_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
  IGI_CHECK_LAUNCH("cam_image_obs_kernel");
  return IGI_OK;
}

extern "C" int igi_pcl_noise(float* pts, int64_t env_stride, int n_envs, int n_pts, const uint8_t* mask,
                             const float* pcl_noise, long long env0, float sigma, float noise_clip, float const_noise,
                             float noise_prob, uint64_t seed, uint32_t step, void* stream) {
  IGI_REQUIRE(pts && pcl_noise, "igi_pcl_noise: null pointer");
  IGI_REQUIRE(n_envs >= 0 && n_pts > 0 && env_stride >= (int64_t)n_pts * 3, "igi_pcl_noise: bad dims");
  if (n_envs == 0) return IGI_OK;
  const long long items = (long long)n_envs * n_pts;
  pcl_noise_kernel<<<(unsigned)((items + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      pts, env_stride, n_envs, n_pts, mask, pcl_noise, env0, sigma, noise_clip, const_noise, noise_prob, seed, step);
  IGI_CHECK_LAUNCH("pcl_noise_kernel");
  return IGI_OK;
}

extern "C" long long igi_rms_scratch_bytes(int channels) {
  if (channels < 1 || channels > RMS_MAX_C) return -1;
  return (long long)RMS_GRID * 2 * channels * (long long)sizeof(double) + 16;
}

extern "C" int igi_rms_forward(const float* x, long long rows, int channels, double* running_mean, double* running_var,
                               double* count, float epsilon, int training, int mode, float* y, void* scratch,
                               void* stream) {
  IGI_REQUIRE(x && running_mean && running_var && count && y, "igi_rms_forward: null pointer");
  IGI_REQUIRE(channels >= 1 && channels <= RMS_MAX_C, "igi_rms_forward: channels must be 1..256");
  IGI_REQUIRE(rows >= 0 && mode >= 0 && mode <= 2, "igi_rms_forward: bad rows / mode");
  IGI_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0, "igi_rms_forward: x and y must be 16-byte aligned");
  if (rows == 0) return IGI_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (training) {
    IGI_REQUIRE(scratch != nullptr, "igi_rms_forward: training needs scratch (igi_rms_scratch_bytes, zero-initialised once)");
    // scratch: [0,16) the done counter (must start as 0; the kernel re-arms it), then the per-CTA partial sums
    unsigned int* done = reinterpret_cast<unsigned int*>(scratch);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 16);
    rms_moments_kernel<<<(RMS_GRID / channels) * channels, RMS_BLOCK, 0, s>>>(x, rows, channels, running_mean, running_var, count, partial, done);
    IGI_CHECK_LAUNCH("rms_moments_kernel");
  }
  const long long total = rows * channels;
  rms_normalize_kernel<<<grid_for(total / 4, 256), 256, 0, s>>>(x, total, channels, running_mean, running_var, epsilon,
                                                                 mode, y);
  IGI_CHECK_LAUNCH("rms_normalize_kernel");
  return IGI_OK;
}

extern "C" int igi_seg_valid_mask(const float* seg, const float* img, long long n, float obj_id, float socket_id,
                                  int distinct, float* seg_out, float* img_out, void* stream) {
  IGI_REQUIRE(seg && seg_out, "igi_seg_valid_mask: null pointer");
  IGI_REQUIRE(img == nullptr || img_out != nullptr, "igi_seg_valid_mask: img given without img_out");
  IGI_REQUIRE(n >= 0 && n % 4 == 0, "igi_seg_valid_mask: element count must be a multiple of 4");
  if (n == 0) return IGI_OK;
  seg_valid_mask_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(seg, img, n / 4, obj_id, socket_id,
                                                                                 distinct, seg_out, img_out);
  IGI_CHECK_LAUNCH("seg_valid_mask_kernel");
  return IGI_OK;
}

extern "C" int igi_queue_push(float* queue, const void* x, int x_is_int32, int64_t x_stride, int n_envs, int hist_len,
                              long long row_len, void* stream) {
  IGI_REQUIRE(queue && x, "igi_queue_push: null pointer");
  IGI_REQUIRE(n_envs >= 0 && hist_len >= 1 && row_len > 0 && row_len % 4 == 0 && x_stride >= row_len,
              "igi_queue_push: row_len must be a positive multiple of 4 and <= x_stride");
  if (n_envs == 0) return IGI_OK;
  const long long items = (long long)n_envs * (row_len / 4);
  if (x_is_int32)
    queue_push_kernel<int32_t><<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
        queue, reinterpret_cast<const int32_t*>(x), x_stride, n_envs, hist_len, row_len / 4);
  else
    queue_push_kernel<float><<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
        queue, reinterpret_cast<const float*>(x), x_stride, n_envs, hist_len, row_len / 4);
  IGI_CHECK_LAUNCH("queue_push_kernel");
  return IGI_OK;
}

extern "C" int igi_copy_rows_where(void* dst, long long dst_stride_bytes, const void* src, long long src_stride_bytes,
                                   const uint8_t* flag, int want, long long rows, long long row_bytes, void* stream) {
  IGI_REQUIRE(dst && src && flag, "igi_copy_rows_where: null pointer");
  IGI_REQUIRE(rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0, "igi_copy_rows_where: row_bytes must be a positive multiple of 16");
  IGI_REQUIRE(dst_stride_bytes >= row_bytes && src_stride_bytes >= row_bytes && dst_stride_bytes % 16 == 0 &&
                  src_stride_bytes % 16 == 0, "igi_copy_rows_where: strides must be multiples of 16 and >= row_bytes");
  IGI_REQUIRE(((uintptr_t)dst % 16) == 0 && ((uintptr_t)src % 16) == 0, "igi_copy_rows_where: buffers must be 16-byte aligned");
  if (rows == 0) return IGI_OK;
  const long long row16 = row_bytes / 16;
  long long gx = (row16 + 255) / 256;
  if (gx > 32) gx = 32;
  const long long gy = rows < 65535 ? rows : 65535;
  copy_rows_where_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(
      (uint4*)dst, dst_stride_bytes / 16, (const uint4*)src, src_stride_bytes / 16, flag, want, rows, row16);
  IGI_CHECK_LAUNCH("copy_rows_where_kernel");
  return IGI_OK;
}

extern "C" int igi_cam_masks(const uint8_t* update_freq, const uint8_t* update_delay, const uint8_t* seg_update_delay,
                             int32_t* got_socket, int socket_pending, uint8_t* upd_seg, uint8_t* upd_pcl,
                             uint8_t* restarted, int n_envs, void* stream) {
  IGI_REQUIRE(update_freq && update_delay && upd_pcl && n_envs >= 0, "igi_cam_masks: null pointer");
  IGI_REQUIRE(!upd_seg || seg_update_delay, "igi_cam_masks: upd_seg needs seg_update_delay");
  IGI_REQUIRE(!socket_pending || got_socket, "igi_cam_masks: socket_pending needs got_socket");
  if (n_envs == 0) return IGI_OK;
  cam_masks_kernel<<<(n_envs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(update_freq, update_delay, seg_update_delay,
                                                                           got_socket, socket_pending, upd_seg, upd_pcl,
                                                                           restarted, n_envs);
  IGI_CHECK_LAUNCH("cam_masks_kernel");
  return IGI_OK;
}

extern "C" int igi_pcl_assemble(const float* src, int64_t src_stride, float* pcl, int64_t pcl_stride, const uint8_t* update,
                                float* queue, int n_envs, int hist_len, long long row_len, void* stream) {
  IGI_REQUIRE(src && pcl, "igi_pcl_assemble: null pointer");
  IGI_REQUIRE(n_envs >= 0 && hist_len >= 1 && row_len > 0 && row_len % 4 == 0 && src_stride >= row_len &&
                  pcl_stride >= row_len && src_stride % 4 == 0 && pcl_stride % 4 == 0,
              "igi_pcl_assemble: row_len and strides must be multiples of 4, strides >= row_len");
  IGI_REQUIRE(((uintptr_t)src % 16) == 0 && ((uintptr_t)pcl % 16) == 0 && ((uintptr_t)queue % 16) == 0,
              "igi_pcl_assemble: buffers must be 16-byte aligned");
  if (n_envs == 0) return IGI_OK;
  const long long items = (long long)n_envs * (row_len / 4);
  pcl_assemble_kernel<<<grid_for(items, 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(src), src_stride / 4, reinterpret_cast<float4*>(pcl), pcl_stride / 4, update,
      reinterpret_cast<float4*>(queue), n_envs, hist_len, row_len / 4);
  IGI_CHECK_LAUNCH("pcl_assemble_kernel");
  return IGI_OK;
}
