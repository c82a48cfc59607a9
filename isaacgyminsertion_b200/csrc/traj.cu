// (L) trajectory logger buffers: the device side of DataLoggerSim
// (algo/ppo/experience.py:352-490).  The per-step observation rows go from the task's device buffers
// straight into per-env episode buffers at each env's own step counter; finished trajectories are
// gathered into one contiguous staging block for a single device->host copy.
//
//   L1 traj_append_kernel   log[e, counter[e], :] = f32(x[e, :])        (:426-434)
//   L2 traj_step_kernel     done_log[e, counter[e]] = done[e]; ++counter[e]; compacts the ids of the
//                           envs that finished, in env order                (:436-444)
//   L3 traj_gather_kernel   staging[j] = log[ids[j]]  (and, with zero_after, log[ids[j]] = 0)   (:448-455, :417-420)
#include "igi_common.cuh"
#include "../../include/igi_b200.h"

namespace {

// One thread = 4 consecutive floats of one env's row (row_len % 4 == 0).
template <typename TIn>
__global__ void __launch_bounds__(256) traj_append_kernel(float* __restrict__ log, const TIn* __restrict__ x,
                                                          int64_t x_stride, const long long* __restrict__ counter,
                                                          int n_envs, int T, long long L4, int* __restrict__ overflow) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)n_envs * L4;
       v += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(v / L4);
    const long long j = v - (long long)e * L4;
    const long long t = counter[e];
    if (t < 0 || t >= T) {   // the reference raises an index error here
      if (j == 0) atomicExch(overflow, 1);
      continue;
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);   // x == NULL: the reference logs zeros for a None value (:430-431)
    if (x) {
      const TIn* src = x + (size_t)e * x_stride + 4 * j;
      o = make_float4((float)src[0], (float)src[1], (float)src[2], (float)src[3]);
    }
    __stcs(reinterpret_cast<float4*>(log) + ((size_t)e * T + (size_t)t) * L4 + j, o);
  }
}

// Rows whose length or stride is not a multiple of 4 floats (arm_joints 7, action 6, ...): one thread per element.
template <typename TIn>
__global__ void __launch_bounds__(256) traj_append_scalar_kernel(float* __restrict__ log, const TIn* __restrict__ x,
                                                                 int64_t x_stride, const long long* __restrict__ counter,
                                                                 int n_envs, int T, long long L, int* __restrict__ overflow) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < (long long)n_envs * L;
       v += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(v / L);
    const long long j = v - (long long)e * L;
    const long long t = counter[e];
    if (t < 0 || t >= T) {
      if (j == 0) atomicExch(overflow, 1);
      continue;
    }
    log[((size_t)e * T + (size_t)t) * L + j] = x ? (float)x[(size_t)e * x_stride + j] : 0.0f;
  }
}

// One CTA, block scan over envs: done flags into the done log, counters advanced, finished env ids compacted in
// env order (torch.nonzero order, :443-445).
__global__ void __launch_bounds__(1024) traj_step_kernel(uint8_t* __restrict__ done_log, const uint8_t* __restrict__ done,
                                                         long long* __restrict__ counter, int n_envs, int T, int pitch,
                                                         int32_t* __restrict__ ids, int32_t* __restrict__ n_done,
                                                         int* __restrict__ overflow) {
  __shared__ int s_w[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_envs; base += 1024) {
    const int e = base + tid;
    int d = 0;
    if (e < n_envs) {
      d = done && done[e] ? 1 : 0;
      const long long t = counter[e];
      if (t >= 0 && t < T) done_log[(size_t)e * pitch + t] = (uint8_t)d;
      else atomicExch(overflow, 1);
      counter[e] = t + 1;
    }
    const unsigned bm = __ballot_sync(0xffffffffu, d);
    if (lane == 0) s_w[warp] = __popc(bm);
    __syncthreads();
    int woff = 0, total = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = s_w[w];
      if (w < warp) woff += c;
      total += c;
    }
    if (d) ids[s_carry + woff + __popc(bm & ((1u << lane) - 1u))] = e;
    __syncthreads();
    if (tid == 0) s_carry += total;
    __syncthreads();
  }
  if (tid == 0) *n_done = s_carry;
}

// staging[j, :] = buf[ids[j], :] for j < n_ids (rows of `row_bytes`, a multiple of 16); zero_after clears the
// source row behind the copy (the reference's _reset_buffers).  One CTA column per listed env.
template <typename W>
__global__ void __launch_bounds__(256) traj_gather_kernel(W* __restrict__ buf, const int32_t* __restrict__ ids,
                                                          const int32_t* __restrict__ n_ids, int max_ids,
                                                          long long row_w, W* __restrict__ staging, int zero_after) {
  const int n = min(*n_ids, max_ids);
  W zero;
  memset(&zero, 0, sizeof(W));
  for (int j = blockIdx.y; j < n; j += gridDim.y) {
    W* src = buf + (size_t)ids[j] * row_w;
    W* dst = staging ? staging + (size_t)j * row_w : nullptr;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < row_w; i += (long long)gridDim.x * blockDim.x) {
      if (dst) dst[i] = src[i];
      if (zero_after) src[i] = zero;
    }
  }
}

__global__ void traj_reset_counters_kernel(long long* counter, const int32_t* ids, const int32_t* n_ids, int max_ids) {
  const int n = min(*n_ids, max_ids);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) counter[ids[j]] = 0;
}

}  // namespace

extern "C" int igi_traj_append(float* log, const void* x, int x_is_int32, int64_t x_stride, const long long* counter,
                               int n_envs, int episode_len, long long row_len, int32_t* overflow, void* stream) {
  IGI_REQUIRE(log && counter && overflow, "igi_traj_append: null pointer");
  IGI_REQUIRE(n_envs >= 0 && episode_len > 0 && row_len > 0, "igi_traj_append: bad dims");
  IGI_REQUIRE(!x || x_stride >= row_len, "igi_traj_append: x_stride < row_len");
  if (n_envs == 0) return IGI_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec = row_len % 4 == 0 && ((uintptr_t)log % 16) == 0;   // x is read element-wise either way
  if (!vec) {
    long long g = ((long long)n_envs * row_len + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (x_is_int32)
      traj_append_scalar_kernel<int32_t><<<(unsigned)g, 256, 0, s>>>(log, (const int32_t*)x, x_stride, counter, n_envs, episode_len, row_len, overflow);
    else
      traj_append_scalar_kernel<float><<<(unsigned)g, 256, 0, s>>>(log, (const float*)x, x_stride, counter, n_envs, episode_len, row_len, overflow);
    IGI_CHECK_LAUNCH("traj_append_scalar_kernel");
    return IGI_OK;
  }
  const long long L4 = row_len / 4, items = (long long)n_envs * L4;
  long long g = (items + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  if (x_is_int32)
    traj_append_kernel<int32_t><<<(unsigned)g, 256, 0, s>>>(log, (const int32_t*)x, x_stride, counter, n_envs, episode_len, L4, overflow);
  else
    traj_append_kernel<float><<<(unsigned)g, 256, 0, s>>>(log, (const float*)x, x_stride, counter, n_envs, episode_len, L4, overflow);
  IGI_CHECK_LAUNCH("traj_append_kernel");
  return IGI_OK;
}

extern "C" int igi_traj_step(uint8_t* done_log, int done_pitch, const uint8_t* done, long long* counter, int n_envs,
                             int episode_len, int32_t* done_ids, int32_t* n_done, int32_t* overflow, void* stream) {
  IGI_REQUIRE(done_log && counter && done_ids && n_done && overflow, "igi_traj_step: null pointer");
  IGI_REQUIRE(n_envs >= 0 && episode_len > 0 && done_pitch >= episode_len, "igi_traj_step: bad dims");
  traj_step_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(done_log, done, counter, n_envs, episode_len, done_pitch, done_ids, n_done, overflow);
  IGI_CHECK_LAUNCH("traj_step_kernel");
  return IGI_OK;
}

extern "C" int igi_traj_gather(void* buf, const int32_t* ids, const int32_t* n_ids, int max_ids, long long row_bytes,
                               void* staging, int zero_after, void* stream) {
  IGI_REQUIRE(buf && ids && n_ids, "igi_traj_gather: null pointer");
  IGI_REQUIRE(max_ids >= 0 && row_bytes > 0 && row_bytes % 4 == 0, "igi_traj_gather: row_bytes must be a positive multiple of 4");
  IGI_REQUIRE(staging || zero_after, "igi_traj_gather: nothing to do");
  if (max_ids == 0) return IGI_OK;
  const bool vec = row_bytes % 16 == 0 && ((uintptr_t)buf % 16) == 0 && ((uintptr_t)staging % 16) == 0;
  const long long row_w = row_bytes / (vec ? 16 : 4);
  long long gx = (row_w + 255) / 256;
  if (gx > 64) gx = 64;
  const int gy = max_ids < 1024 ? max_ids : 1024;
  const dim3 grid((unsigned)gx, (unsigned)gy);
  if (vec)
    traj_gather_kernel<uint4><<<grid, 256, 0, (cudaStream_t)stream>>>((uint4*)buf, ids, n_ids, max_ids, row_w, (uint4*)staging, zero_after);
  else
    traj_gather_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((uint32_t*)buf, ids, n_ids, max_ids, row_w, (uint32_t*)staging, zero_after);
  IGI_CHECK_LAUNCH("traj_gather_kernel");
  return IGI_OK;
}

extern "C" int igi_traj_reset_counters(long long* counter, const int32_t* ids, const int32_t* n_ids, int max_ids, void* stream) {
  IGI_REQUIRE(counter && ids && n_ids && max_ids >= 0, "igi_traj_reset_counters: bad args");
  if (max_ids == 0) return IGI_OK;
  traj_reset_counters_kernel<<<(max_ids + 255) / 256, 256, 0, (cudaStream_t)stream>>>(counter, ids, n_ids, max_ids);
  IGI_CHECK_LAUNCH("traj_reset_counters_kernel");
  return IGI_OK;
}
