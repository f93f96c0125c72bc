// Backward of the dequant-matmul with respect to the activation:
//     grad_A[M, K] = grad_out[M, N] · dequant(W)^T
// int4g32: `grad_out.matmul(unpack_int4(B, b_scale).t())` (chatglm_q/int4/qlinear.py:53-64; Triton twin
// int4/triton_ops.py:142-264); int8: `grad_out.matmul((B * b_scale).t())` (int8/qlinear.py:41-52,
// int8/triton_ops.py:130-245).  Only needed for P-tuning / LoRA-style training on a frozen quantised model
// (`generate` runs under no_grad), so this is a CUDA-core kernel, correct for every shape (the reference's Triton
// version asserts power-of-two sizes and cannot take out_features = 13696 / 27392 / 65024): each weight element is
// dequantised with the reference's single rounding, products are accumulated in fp32, the result is rounded once.
// A [32 k x 32 m] output tile per CTA, 32-wide chunks of N staged through shared memory.
#include "common.cuh"

namespace cgq {
namespace {

constexpr int TK = 32, TM = 32, TNB = 32;

template <typename T, bool kW8>
__global__ void __launch_bounds__(256)
    grad_a_kernel(const T* __restrict__ G, int64_t ldg, const void* __restrict__ Wq, const T* __restrict__ S,
                  T* __restrict__ out, int64_t ldo, int M, int N, int K) {
  __shared__ float Ws[TNB][TK + 1];   // [n][k] dequantised weight values (as T, held in fp32)
  __shared__ float Gs[TM][TNB + 1];   // [m][n]
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int k0 = blockIdx.x * TK, m0 = blockIdx.y * TM;
  float acc[TM / 8];
#pragma unroll
  for (int i = 0; i < TM / 8; ++i) acc[i] = 0.f;
  for (int n0 = 0; n0 < N; n0 += TNB) {
#pragma unroll
    for (int i = 0; i < TK / 8; ++i) {
      const int kk = ty + 8 * i, k = k0 + kk, n = n0 + tx;
      float w = 0.f;
      if (k < K && n < N) {
        if (kW8) {
          const int8_t q = static_cast<const int8_t*>(Wq)[static_cast<int64_t>(n) * K + k];
          w = DT<T>::to_f(dequant8<T>(q, S[n]));
        } else {
          const uint8_t b = static_cast<const uint8_t*>(Wq)[static_cast<int64_t>(k >> 1) * N + n];
          w = DT<T>::to_f(dequant4<T>((b >> ((k & 1) * 4)) & 0xF, S[static_cast<int64_t>(k >> 5) * N + n]));
        }
      }
      Ws[tx][kk] = w;
    }
#pragma unroll
    for (int i = 0; i < TM / 8; ++i) {
      const int mm = ty + 8 * i, m = m0 + mm, n = n0 + tx;
      Gs[mm][tx] = (m < M && n < N) ? DT<T>::to_f(G[static_cast<int64_t>(m) * ldg + n]) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int nn = 0; nn < TNB; ++nn) {
      const float w = Ws[nn][tx];
#pragma unroll
      for (int i = 0; i < TM / 8; ++i) acc[i] = fmaf(Gs[ty + 8 * i][nn], w, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM / 8; ++i) {
    const int m = m0 + ty + 8 * i, k = k0 + tx;
    if (m < M && k < K) out[static_cast<int64_t>(m) * ldo + k] = DT<T>::from_f(acc[i]);
  }
}

template <bool kW8>
int launch(const void* G, int64_t ldg, const void* Wq, const void* scale, void* out, int64_t ldo, int M, int N, int K,
           int dtype, cudaStream_t st) {
  const dim3 grid((K + TK - 1) / TK, (M + TM - 1) / TM), block(256);
  if (dtype == CGQ_DTYPE_F16)
    grad_a_kernel<__half, kW8><<<grid, block, 0, st>>>(static_cast<const __half*>(G), ldg, Wq,
                                                       static_cast<const __half*>(scale), static_cast<__half*>(out), ldo,
                                                       M, N, K);
  else
    grad_a_kernel<__nv_bfloat16, kW8><<<grid, block, 0, st>>>(static_cast<const __nv_bfloat16*>(G), ldg, Wq,
                                                              static_cast<const __nv_bfloat16*>(scale),
                                                              static_cast<__nv_bfloat16*>(out), ldo, M, N, K);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

int check(const char* fn, const void* G, int64_t ldg, const void* Wq, const void* scale, const void* out, int64_t ldo,
          int M, int N, int K, int dtype) {
  if (M < 0 || N <= 0 || K <= 0 || ldg < N || ldo < K) {
    set_error("%s: bad shape M=%d N=%d K=%d ldg=%lld ldo=%lld", fn, M, N, K, (long long)ldg, (long long)ldo);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("%s: unsupported dtype code %d (0=f16, 1=bf16)", fn, dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  if (M > 0 && (G == nullptr || out == nullptr || Wq == nullptr || scale == nullptr)) {
    set_error("%s: null pointer", fn);
    return CGQ_ERR_MISALIGNED;
  }
  return CGQ_OK;
}

}  // namespace
}  // namespace cgq

using namespace cgq;

extern "C" int cgq_w4a16_grad_a(const void* grad_out, int64_t ldg, const uint8_t* Wq, const void* scale, void* grad_a,
                                int64_t ldo, int M, int N, int K, int group, int dtype, void* stream) {
  int rc = check("cgq_w4a16_grad_a", grad_out, ldg, Wq, scale, grad_a, ldo, M, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  if (group != 32 || K % 32 != 0) {
    set_error("cgq_w4a16_grad_a: group must be 32 and divide K (group=%d, K=%d)", group, K);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (M == 0) return CGQ_OK;
  return launch<false>(grad_out, ldg, Wq, scale, grad_a, ldo, M, N, K, dtype, static_cast<cudaStream_t>(stream));
}

extern "C" int cgq_w8a16_grad_a(const void* grad_out, int64_t ldg, const int8_t* Wq, const void* scale, void* grad_a,
                                int64_t ldo, int M, int N, int K, int dtype, void* stream) {
  int rc = check("cgq_w8a16_grad_a", grad_out, ldg, Wq, scale, grad_a, ldo, M, N, K, dtype);
  if (rc != CGQ_OK) return rc;
  if (M == 0) return CGQ_OK;
  return launch<true>(grad_out, ldg, Wq, scale, grad_a, ldo, M, N, K, dtype, static_cast<cudaStream_t>(stream));
}
