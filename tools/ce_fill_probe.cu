// Round-2 probe for DESIGN.md section 8 item 0: can the DMA engines write the per-step no-contact fill
// (gel_depth zeros: one 2.47 GB memset; colour: 12288 copies of 150 528 B from 8 background images) WHILE a
// kernel that owns every SM's register file runs?   Build + run on the box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ce_fill_probe tools/ce_fill_probe.cu && /tmp/ce_fill_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// 2 CTAs x 512 threads x 64 registers per SM = the whole register file, pure issue-bound work for ~`iters` FMAs
__global__ void __launch_bounds__(512, 2) busy(float* out, int iters) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 0.5f, d = 0.25f;
  for (int i = 0; i < iters; ++i) { a = fmaf(a, b, c); c = fmaf(c, b, d); d = fmaf(d, b, a); b = fmaf(b, 0.99999f, 1e-7f); }
  if (a + b + c + d == 12345.f) out[blockIdx.x] = a;
}

static float run(cudaStream_t sk, cudaStream_t sc, float* out, int sms, int iters, int mode, void* gel, size_t gel_bytes,
                 std::vector<void*>& dsts, std::vector<void*>& srcs, std::vector<size_t>& sizes) {
  cudaEvent_t e0, e1, ec;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&ec));
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0, sk));
  CK(cudaStreamWaitEvent(sc, e0, 0));
  if (mode & 1) busy<<<sms * 2, 512, 0, sk>>>(out, iters);
  if (mode & 2) CK(cudaMemsetAsync(gel, 0, gel_bytes, sc));
  if (mode & 4) {
    cudaMemcpyAttributes at{};
    at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    at.flags = cudaMemcpyFlagPreferOverlapWithCompute;
    size_t idx = 0, fail = 0;
    CK(cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &at, &idx, 1, &fail, sc));
  }
  CK(cudaEventRecord(ec, sc));
  CK(cudaStreamWaitEvent(sk, ec, 0));
  CK(cudaEventRecord(e1, sk));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms;
}

int main() {
  int sms = 148;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t F = 12288, img = 224 * 224 * 3, gel_bytes = F * 224 * 224 * 4;
  void *gel, *color, *bg;
  float* out;
  CK(cudaMalloc(&gel, gel_bytes)); CK(cudaMalloc(&color, F * img)); CK(cudaMalloc(&bg, 8 * img)); CK(cudaMalloc(&out, 4096));
  std::vector<void*> dsts(F), srcs(F);
  std::vector<size_t> sizes(F, img);
  for (size_t f = 0; f < F; ++f) { dsts[f] = (char*)color + f * img; srcs[f] = (char*)bg + (f * 2654435761u % 8) * img; }
  cudaStream_t sk, sc;
  CK(cudaStreamCreate(&sk)); CK(cudaStreamCreate(&sc));
  int iters = 1 << 16;
  run(sk, sc, out, sms, iters, 1, gel, gel_bytes, dsts, srcs, sizes);
  float t = run(sk, sc, out, sms, iters, 1, gel, gel_bytes, dsts, srcs, sizes);
  iters = (int)(iters * 3.0f / t);   // ~3 ms of compute, the length of one tactile step
  const char* names[] = {"", "kernel alone", "memset alone (2.47 GB)", "kernel + memset", "batch copy alone (12288 x 150 KB)",
                         "kernel + batch copy", "memset + batch copy", "kernel + memset + batch copy"};
  for (int mode = 1; mode <= 7; ++mode) {
    run(sk, sc, out, sms, iters, mode, gel, gel_bytes, dsts, srcs, sizes);
    printf("%-40s %8.3f ms\n", names[mode], run(sk, sc, out, sms, iters, mode, gel, gel_bytes, dsts, srcs, sizes));
  }
  printf("overlap is free when 'kernel + X' ~= max(kernel alone, X alone)\n");
  return 0;
}
