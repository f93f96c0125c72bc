// Peer-visible device buffers for the tensor-parallel step (include/cgq.h, cgq_ipc_*): plain cudaMalloc memory
// exported / imported with the legacy CUDA IPC handles (64 bytes, exchanged by the host over torch.distributed),
// so that the step kernel of one rank can store its partial sums / logits straight into its peers' buffers over
// NVLink.  No reference counterpart (the reference is single-GPU).
#include <string.h>

#include "common.cuh"

using namespace cgq;

namespace {
// cross-GPU barrier of one token step (after the broadcast lm_head): kernel boundaries order this rank's peer stores
// before the flag, the flag is released at system scope, and every rank waits for all ranks' flags.
__device__ __forceinline__ void tp_barrier_body(uint32_t* const* flags, int world, int rank, const int* step, uint32_t* err) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint32_t epoch = static_cast<uint32_t>(*step);
  __threadfence_system();
  for (int r = 0; r < world; ++r)
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags[r] + rank), "r"(epoch) : "memory");
  for (int r = 0; r < world; ++r) {
    uint32_t v, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags[rank] + r) : "memory");
      if (static_cast<int32_t>(v - epoch) >= 0) break;
      if (++spins > (1u << 26)) {
        if (err != nullptr) *err = 0x80000000u | epoch;
        break;
      }
    }
  }
}
struct FlagPtrs {
  uint32_t* p[8];
};
__global__ void tp_barrier_kernel8(FlagPtrs f, int world, int rank, const int* step, uint32_t* err) {
  tp_barrier_body(f.p, world, rank, step, err);
}
}  // namespace

extern "C" int cgq_tp_next(const cgq_tp_ctx* ctx, uint32_t idx) {
  if (ctx == nullptr || ctx->world < 1 || ctx->world > 8 || ctx->rank < 0 || ctx->rank >= ctx->world || idx >= 127 ||
      (ctx->recv[0] != nullptr && (ctx->max_n <= 0 || ctx->step == nullptr))) {
    set_error("cgq_tp_next: bad context (world 1..8, rank < world, idx < 127, max_n > 0, step != NULL)");
    return CGQ_ERR_BAD_SHAPE;
  }
  set_w4_tp(*ctx, idx);
  return CGQ_OK;
}

extern "C" int cgq_tp_barrier(uint32_t* const* flags, int world, int rank, const int* step, uint32_t* err, void* stream) {
  if (flags == nullptr || world < 1 || world > 8 || rank < 0 || rank >= world || step == nullptr) {
    set_error("cgq_tp_barrier: bad arguments");
    return CGQ_ERR_BAD_SHAPE;
  }
  FlagPtrs f = {};
  for (int r = 0; r < world; ++r) f.p[r] = flags[r];
  tp_barrier_kernel8<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, world, rank, step, err);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

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
