// (G) observation gather onto the learner rank without SM work: peer buffers over CUDA IPC, copy-engine
// peer copies over NVLink, and stream memory operations (write / wait on a 32-bit word) as the only
// synchronisation, so neither the data movement nor the hand-shake needs a kernel while tac_contact's
// persistent CTAs own every SM (north star "NCCL over NVLink ... to all-gather observations
// onto the learner rank" - the NCCL transport stays available in dist.ObsGather).
//
// Host side: isaacgyminsertion_b200/dist.py (ObsGather, transport "p2p").
#include <cuda.h>
#include <string.h>

#include "igi_common.cuh"
#include "../../include/igi_b200.h"

namespace {
typedef CUresult (*StreamWriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWriteValue32Fn g_write32 = nullptr;
StreamWaitValue32Fn g_wait32 = nullptr;

// Driver entry points through the runtime (no link-time dependency on libcuda: the build container has no driver).
int load_memops() {
  if (g_write32 && g_wait32) return IGI_OK;
  void* fw = nullptr;
  void* fa = nullptr;
  cudaDriverEntryPointQueryResult qr;
  IGI_CUDA(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fw, cudaEnableDefault, &qr));
  IGI_REQUIRE(qr == cudaDriverEntryPointSuccess && fw, "cuStreamWriteValue32 not available in this driver");
  IGI_CUDA(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fa, cudaEnableDefault, &qr));
  IGI_REQUIRE(qr == cudaDriverEntryPointSuccess && fa, "cuStreamWaitValue32 not available in this driver");
  g_write32 = (StreamWriteValue32Fn)fw;
  g_wait32 = (StreamWaitValue32Fn)fa;
  return IGI_OK;
}
}  // namespace

extern "C" int igi_peer_alloc(unsigned long long bytes, void** ptr_out, unsigned char* handle64_out) {
  IGI_REQUIRE(bytes > 0 && ptr_out && handle64_out, "igi_peer_alloc: bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void* p = nullptr;
  IGI_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    igi_set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    return IGI_ERR_CUDA;
  }
  IGI_CUDA(cudaMemset(p, 0, (size_t)bytes));
  memcpy(handle64_out, &h, 64);
  *ptr_out = p;
  return IGI_OK;
}

extern "C" int igi_peer_free(void* ptr) {
  if (ptr) IGI_CUDA(cudaFree(ptr));
  return IGI_OK;
}

extern "C" int igi_peer_open(const unsigned char* handle64, void** ptr_out) {
  IGI_REQUIRE(handle64 && ptr_out, "igi_peer_open: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  IGI_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return IGI_OK;
}

extern "C" int igi_peer_close(void* ptr) {
  if (ptr) IGI_CUDA(cudaIpcCloseMemHandle(ptr));
  return IGI_OK;
}

extern "C" int igi_peer_copy_async(void* dst, const void* src, unsigned long long bytes, void* stream) {
  IGI_REQUIRE(dst && src, "igi_peer_copy_async: null pointer");
  if (bytes == 0) return IGI_OK;
  // unified addressing resolves the devices; a device-to-device copy between peers runs on a copy engine
  IGI_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return IGI_OK;
}

extern "C" int igi_stream_write_value32(void* stream, void* addr, unsigned int value) {
  IGI_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "igi_stream_write_value32: address must be 4-byte aligned");
  int rc = load_memops();
  if (rc != IGI_OK) return rc;
  CUresult r = g_write32((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
  if (r != CUDA_SUCCESS) {
    igi_set_error("cuStreamWriteValue32 failed (CUresult %d)", (int)r);
    return IGI_ERR_CUDA;
  }
  return IGI_OK;
}

extern "C" int igi_stream_wait_value32_geq(void* stream, void* addr, unsigned int value) {
  IGI_REQUIRE(addr && ((uintptr_t)addr & 3) == 0, "igi_stream_wait_value32_geq: address must be 4-byte aligned");
  int rc = load_memops();
  if (rc != IGI_OK) return rc;
  // (int32)(*addr - value) >= 0: cyclic comparison, safe across the 2^32 wrap of a step counter
  CUresult r = g_wait32((CUstream)stream, (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
  if (r != CUDA_SUCCESS) {
    igi_set_error("cuStreamWaitValue32 failed (CUresult %d)", (int)r);
    return IGI_ERR_CUDA;
  }
  return IGI_OK;
}
