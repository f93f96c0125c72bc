// Shared host/device helpers: dtype traits, error plumbing, launch-parameter structs.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cgq.h"

namespace cgq {

// ------------------------------------------------------------------ errors (host)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CGQ_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::cgq::cuda_fail(_e, #expr); \
  } while (0)

// ------------------------------------------------------------------ dtype traits (device)
template <typename T>
struct DT;

template <>
struct DT<__half> {
  using T2 = __half2;
  static constexpr int code = CGQ_DTYPE_F16;
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ __half from_i(int v) { return __int2half_rn(v); }
};
template <>
struct DT<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  static constexpr int code = CGQ_DTYPE_BF16;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_i(int v) { return __int2bfloat16_rn(v); }
};

// Reference dequant of one int4 element: round_T( (q - 8) * s ), one rounding
// (chatglm_q/int4/qlinear.py:31-32: int8 tensor times scale tensor in the scale's dtype).
// (q-8) is exact in fp16/bf16 and the product of two such values is exact in fp32.
template <typename T>
__device__ __forceinline__ T dequant4(int nib, T s) {
  return DT<T>::from_f(static_cast<float>(nib - 8) * DT<T>::to_f(s));
}
// Reference dequant of one int8 element: round_T( q * s ) (int8/qlinear.py:38, `B * b_scale`).
template <typename T>
__device__ __forceinline__ T dequant8(int q, T s) {
  return DT<T>::from_f(static_cast<float>(q) * DT<T>::to_f(s));
}
// Epilogue shared by every GEMM kernel: round the fp32 accumulator to T, then (optionally) add the
// bias as a second rounded T operation — `out += self.bias` (int4/qlinear.py:92-93).
template <typename T>
__device__ __forceinline__ T epilogue(float acc, const T* bias, int n) {
  T c = DT<T>::from_f(acc);
  if (bias != nullptr) c = DT<T>::from_f(DT<T>::to_f(c) + DT<T>::to_f(bias[n]));
  return c;
}

// ------------------------------------------------------------------ workspace layout
// [0, kCounterBytes)            : int32 tile counters (self-cleaning), one per 128-byte line
// [kCounterBytes, total)        : fp32 stream-K partial tiles, 2 slots per CTA
constexpr int kMaxCtas = 148 * 4;
constexpr int kMaxTiles = 8192;
constexpr int kCounterStride = 32;  // ints: one 128-byte line per tile counter (atomics on one line serialise)
constexpr size_t kCounterBytes = sizeof(int) * kCounterStride * kMaxTiles;
constexpr size_t kSlotFloats = 8 * 128;  // M_MAX x BN
constexpr size_t kStreamKBytes = sizeof(float) * kSlotFloats * 2 * kMaxCtas;
// split-K partial tiles of the tcgen05 prefill kernel at small M (gemm_tc.cu): [items][128][MB] fp32
constexpr size_t kPartialOffset = (kCounterBytes + kStreamKBytes + 1023) / 1024 * 1024;
constexpr size_t kPartialBytes = 64u << 20;
constexpr size_t kWorkspaceBytes = kPartialOffset + kPartialBytes;

// ------------------------------------------------------------------ kernel launchers (one per .cu)
struct GemmArgs {
  const void* A;
  int64_t lda;
  const void* Wq;
  const void* scale;
  const void* bias;
  void* C;
  int64_t ldc;
  int M, N, K;
  int dtype;
  void* workspace;
  cudaStream_t stream;
};

// Fused decode-step extras of the M == 1 int4 kernel (cgq_w4a16_gemv_fused).
struct GemvFused {
  int prologue;        // CGQ_PRO_*
  const void* norm_w;  // [K] RMSNorm weight (CGQ_PRO_RMSNORM)
  float eps;
  const void* resid;   // [N] residual added after the product is rounded, or null
};

int launch_w4_simple(const GemmArgs& a);
int launch_w4_gemv_fused(const GemmArgs& a, const GemvFused& fu);
void set_next_w4_hint(const void* w, const void* s, int N, int K);
void set_w4_handover(const unsigned* wait_ctr, unsigned wait_count, unsigned* signal_ctr);
void set_w4_tp(const cgq_tp_ctx& ctx, unsigned idx);
int w4_gemv_tiles(int N);
int launch_w8_simple(const GemmArgs& a);
// arithmetic of the int4 decode kernel (gemv_w4.cu).  DEFAULT = CGQ_GEMV_ARITH (0 = IMMA unless set): integer MMA on
// base-256 digits of the activation at M == 1, the subnormal-operand f16 MMA at M > 1; EXACT = (q - 8) converted
// exactly to T, f16 / bf16 MMA; SUBNORMAL = fp16 nibbles as subnormal operands at every M (round-1/2 kernel).
enum { W4_ARITH_DEFAULT = -1, W4_ARITH_IMMA = 0, W4_ARITH_EXACT = 1, W4_ARITH_SUBNORMAL = 2 };
int launch_w4_gemv(const GemmArgs& a, int arith);
int default_w4_arith(int set);
int launch_w4_gemv_umma(const GemmArgs& a, bool* taken);
int launch_w8_gemv(const GemmArgs& a);
int launch_w8_gemv_fused(const GemmArgs& a, const GemvFused& fu);
int launch_w4_tc(const GemmArgs& a);
bool w4_tc_supported(const GemmArgs& a);
int launch_w8_tc(const GemmArgs& a);
bool w8_tc_supported(const GemmArgs& a);
bool w4_gemv_supported(const GemmArgs& a);
bool w8_gemv_supported(const GemmArgs& a);

int sm_count();
// One-shot timeline buffer for the next decode-kernel launch (cgq_debug_trace); nullptr if none.
void* take_trace_buffer();

}  // namespace cgq
