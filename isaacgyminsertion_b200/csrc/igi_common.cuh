// Shared helpers for the sm_100a kernels behind the C-ABI in include/igi_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define IGI_OK 0
#define IGI_ERR_BAD_ARG (-1)
#define IGI_ERR_CUDA (-2)
#define IGI_ERR_UNSUPPORTED (-3)

void igi_set_error(const char* fmt, ...);
void igi_count_launch();  // every kernel launch of the library goes through IGI_CHECK_LAUNCH

#define IGI_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      igi_set_error(__VA_ARGS__);              \
      return IGI_ERR_BAD_ARG;                  \
    }                                          \
  } while (0)

// Launch errors are surfaced without synchronising (callee never syncs).
#define IGI_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    igi_count_launch();                                                     \
    cudaError_t e__ = cudaPeekAtLastError();                                \
    if (e__ != cudaSuccess) {                                               \
      igi_set_error("%s: %s", name, cudaGetErrorString(e__));               \
      (void)cudaGetLastError();                                             \
      return IGI_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

#define IGI_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      igi_set_error("%s: %s", #call, cudaGetErrorString(e__));              \
      return IGI_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----------------
__device__ __forceinline__ uint32_t igi_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void igi_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(igi_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void igi_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(igi_smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void igi_mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(igi_smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void igi_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          igi_smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(igi_smem_u32(bar))
      : "memory");
}
// shared -> global bulk copy (asynchronous, bulk-group completion); bytes % 16 == 0, 16-byte aligned.
__device__ __forceinline__ void igi_bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(igi_smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void igi_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until the bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void igi_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// waits until the bulk groups of this thread have COMPLETED (their global writes are performed)
__device__ __forceinline__ void igi_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// ... until at most the most recent bulk group of this thread is still pending
__device__ __forceinline__ void igi_bulk_wait1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void igi_fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void igi_fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// fire-and-forget prefetch of the 128-byte line at p into L2
__device__ __forceinline__ void igi_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ int igi_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ int igi_warp() { return threadIdx.x >> 5; }
