// Peer-visible device buffers for the tensor-parallel step (include/cgq.h, cgq_ipc_*): plain cudaMalloc memory
// exported / imported with the legacy CUDA IPC handles (64 bytes, exchanged by the host over torch.distributed),
// so that the step kernel of one rank can store its partial sums / logits straight into its peers' buffers over
// NVLink.  No reference counterpart (the reference is single-GPU).
#include <string.h>

#include "common.cuh"

using namespace cgq;

extern "C" int cgq_ipc_alloc(size_t bytes, void** ptr, void* handle64) {
  if (bytes == 0 || ptr == nullptr || handle64 == nullptr) {
    set_error("cgq_ipc_alloc: bad arguments");
    return CGQ_ERR_BAD_SHAPE;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void* p = nullptr;
  CGQ_CUDA_TRY(cudaMalloc(&p, bytes));
  CGQ_CUDA_TRY(cudaMemset(p, 0, bytes));
  CGQ_CUDA_TRY(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "cudaIpcGetMemHandle");
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return CGQ_OK;
}

extern "C" int cgq_ipc_open(const void* handle64, void** ptr) {
  if (handle64 == nullptr || ptr == nullptr) {
    set_error("cgq_ipc_open: bad arguments");
    return CGQ_ERR_BAD_SHAPE;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  CGQ_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr = p;
  return CGQ_OK;
}

extern "C" int cgq_ipc_close(void* ptr) {
  if (ptr != nullptr) CGQ_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return CGQ_OK;
}

extern "C" int cgq_ipc_free(void* ptr) {
  if (ptr != nullptr) CGQ_CUDA_TRY(cudaFree(ptr));
  return CGQ_OK;
}
