// Element-wise pieces of the path (unpack / dequant / embedding gather) and the shape-general
// CUDA-core GEMM kernels.  The "simple" GEMMs are bit-faithful to the reference dequant
// (weight rounded once to the activation dtype, fp32 accumulation) and take ANY shape and
// alignment; they serve shapes the TMA kernels cannot (N % 16 != 0, odd strides) and are the
// on-device cross-check of the fast kernels.  They are CUDA kernels, not a CPU fallback.
#include "common.cuh"

namespace cgq {

// ------------------------------------------------------------------ int4 unpack -> int8
// chatglm_q/int4/qlinear.py:29-31
__global__ void w4_unpack_i8_kernel(const uint8_t* __restrict__ Wq, int8_t* __restrict__ out,
                                    int64_t rows, int N) {
  int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int64_t total = rows * N;
  if (idx >= total) return;
  int64_t r = idx / N;
  int n = static_cast<int>(idx - r * N);
  uint8_t b = Wq[idx];
  out[(2 * r) * N + n] = static_cast<int8_t>(static_cast<int>(b & 0xF) - 8);
  out[(2 * r + 1) * N + n] = static_cast<int8_t>(static_cast<int>(b >> 4) - 8);
}

// ------------------------------------------------------------------ int4 dequant -> T
// chatglm_q/int4/qlinear.py:20-33
template <typename T>
__global__ void w4_dequant_kernel(const uint8_t* __restrict__ Wq, const T* __restrict__ scale,
                                  T* __restrict__ out, int64_t rows, int N, int group) {
  int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int64_t total = rows * N;
  if (idx >= total) return;
  int64_t r = idx / N;
  int n = static_cast<int>(idx - r * N);
  uint8_t b = Wq[idx];
  T s = scale[((2 * r) / group) * N + n];  // group is even: k=2r and k=2r+1 share a group
  out[(2 * r) * N + n] = dequant4<T>(b & 0xF, s);
  out[(2 * r + 1) * N + n] = dequant4<T>(b >> 4, s);
}

// ------------------------------------------------------------------ embedding gathers
// int4: chatglm_q/int4/qlinear.py:122-130 (packed along the vocab axis)
template <typename T>
__global__ void w4_embedding_kernel(const int64_t* __restrict__ ids, int n_ids,
                                    const uint8_t* __restrict__ Wq, const T* __restrict__ scale,
                                    T* __restrict__ out, int D, int group) {
  int i = blockIdx.x;
  if (i >= n_ids) return;
  int64_t t = ids[i];
  const uint8_t* wrow = Wq + (t >> 1) * D;
  const T* srow = scale + (t / group) * D;
  int shift = static_cast<int>(t & 1) * 4;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[static_cast<int64_t>(i) * D + d] = dequant4<T>((wrow[d] >> shift) & 0xF, srow[d]);
}
// int8: chatglm_q/int8/qlinear.py:118-120
template <typename T>
__global__ void w8_embedding_kernel(const int64_t* __restrict__ ids, int n_ids,
                                    const int8_t* __restrict__ Wq, const T* __restrict__ scale,
                                    T* __restrict__ out, int D) {
  int i = blockIdx.x;
  if (i >= n_ids) return;
  const int8_t* wrow = Wq + ids[i] * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    out[static_cast<int64_t>(i) * D + d] = dequant8<T>(wrow[d], scale[d]);
}

// ------------------------------------------------------------------ simple int4 GEMM
// One thread per output column, kRows token rows per CTA; the weight is dequantised exactly as
// unpack_int4 does and the dot accumulates in fp32 (int4/triton_ops.py:66-79).
constexpr int kSimpleRows = 8;
constexpr int kSimpleCols = 128;

template <typename T>
__global__ void __launch_bounds__(kSimpleCols)
    w4_simple_gemm_kernel(const T* __restrict__ A, int64_t lda, const uint8_t* __restrict__ Wq,
                          const T* __restrict__ scale, const T* __restrict__ bias,
                          T* __restrict__ C, int64_t ldc, int M, int N, int K) {
  __shared__ float a_s[kSimpleRows][32];
  const int n = blockIdx.x * kSimpleCols + threadIdx.x;
  const int m0 = blockIdx.y * kSimpleRows;
  const bool active = n < N;
  float acc[kSimpleRows];
#pragma unroll
  for (int m = 0; m < kSimpleRows; ++m) acc[m] = 0.f;

  const int G = K / 32;
  for (int g = 0; g < G; ++g) {
    __syncthreads();
    for (int i = threadIdx.x; i < kSimpleRows * 32; i += kSimpleCols) {
      int m = i / 32, kk = i % 32;
      a_s[m][kk] = (m0 + m < M) ? DT<T>::to_f(A[(m0 + m) * lda + g * 32 + kk]) : 0.f;
    }
    __syncthreads();
    if (active) {
      const T s = scale[static_cast<int64_t>(g) * N + n];
#pragma unroll 4
      for (int r = 0; r < 16; ++r) {
        uint8_t b = Wq[(static_cast<int64_t>(g) * 16 + r) * N + n];
        float w0 = DT<T>::to_f(dequant4<T>(b & 0xF, s));
        float w1 = DT<T>::to_f(dequant4<T>(b >> 4, s));
#pragma unroll
        for (int m = 0; m < kSimpleRows; ++m) {
          acc[m] = fmaf(a_s[m][2 * r], w0, acc[m]);
          acc[m] = fmaf(a_s[m][2 * r + 1], w1, acc[m]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int m = 0; m < kSimpleRows; ++m)
      if (m0 + m < M) C[(m0 + m) * ldc + n] = epilogue<T>(acc[m], bias, n);
  }
}

// ------------------------------------------------------------------ simple int8 GEMM
// Weight [N, K] K-contiguous; one warp per output column, lanes stride K, shuffle reduce.
// Dequant `q * scale[n]` is rounded to T per element before the dot (int8/triton_ops.py:70-71).
template <typename T>
__global__ void __launch_bounds__(256)
    w8_simple_gemm_kernel(const T* __restrict__ A, int64_t lda, const int8_t* __restrict__ Wq,
                          const T* __restrict__ scale, const T* __restrict__ bias,
                          T* __restrict__ C, int64_t ldc, int M, int N, int K) {
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int n = blockIdx.x * 8 + warp;
  const int m0 = blockIdx.y * kSimpleRows;
  if (n >= N) return;
  const T s = scale[n];
  const int8_t* wrow = Wq + static_cast<int64_t>(n) * K;
  float acc[kSimpleRows];
#pragma unroll
  for (int m = 0; m < kSimpleRows; ++m) acc[m] = 0.f;
  for (int k = lane; k < K; k += 32) {
    float w = DT<T>::to_f(dequant8<T>(wrow[k], s));
#pragma unroll
    for (int m = 0; m < kSimpleRows; ++m)
      if (m0 + m < M) acc[m] = fmaf(DT<T>::to_f(A[(m0 + m) * lda + k]), w, acc[m]);
  }
#pragma unroll
  for (int m = 0; m < kSimpleRows; ++m) {
    float v = acc[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && m0 + m < M) C[(m0 + m) * ldc + n] = epilogue<T>(v, bias, n);
  }
}

// ------------------------------------------------------------------ launchers
template <typename T>
static int launch_w4_simple_t(const GemmArgs& a) {
  dim3 grid((a.N + kSimpleCols - 1) / kSimpleCols, (a.M + kSimpleRows - 1) / kSimpleRows);
  w4_simple_gemm_kernel<T><<<grid, kSimpleCols, 0, a.stream>>>(
      static_cast<const T*>(a.A), a.lda, static_cast<const uint8_t*>(a.Wq),
      static_cast<const T*>(a.scale), static_cast<const T*>(a.bias), static_cast<T*>(a.C), a.ldc,
      a.M, a.N, a.K);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}
int launch_w4_simple(const GemmArgs& a) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_w4_simple_t<__half>(a)
                                  : launch_w4_simple_t<__nv_bfloat16>(a);
}

template <typename T>
static int launch_w8_simple_t(const GemmArgs& a) {
  dim3 grid((a.N + 7) / 8, (a.M + kSimpleRows - 1) / kSimpleRows);
  w8_simple_gemm_kernel<T><<<grid, 256, 0, a.stream>>>(
      static_cast<const T*>(a.A), a.lda, static_cast<const int8_t*>(a.Wq),
      static_cast<const T*>(a.scale), static_cast<const T*>(a.bias), static_cast<T*>(a.C), a.ldc,
      a.M, a.N, a.K);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}
int launch_w8_simple(const GemmArgs& a) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_w8_simple_t<__half>(a)
                                  : launch_w8_simple_t<__nv_bfloat16>(a);
}

}  // namespace cgq

// ------------------------------------------------------------------ C-ABI: element-wise entry points
using namespace cgq;

extern "C" int cgq_w4_unpack_i8(const uint8_t* Wq, int8_t* out, int K, int N, void* stream) {
  if (K <= 0 || N <= 0 || (K & 1)) {
    set_error("cgq_w4_unpack_i8: bad shape K=%d N=%d", K, N);
    return CGQ_ERR_BAD_SHAPE;
  }
  int64_t rows = K / 2, total = rows * N;
  int block = 256;
  w4_unpack_i8_kernel<<<static_cast<unsigned>((total + block - 1) / block), block, 0,
                        static_cast<cudaStream_t>(stream)>>>(Wq, out, rows, N);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

extern "C" int cgq_w4_dequant(const uint8_t* Wq, const void* scale, void* out, int K, int N,
                              int group, int dtype, void* stream) {
  if (K <= 0 || N <= 0 || group <= 0 || (group & 1) || K % group != 0) {
    set_error("cgq_w4_dequant: bad shape K=%d N=%d group=%d", K, N, group);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("cgq_w4_dequant: bad dtype %d", dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  int64_t rows = K / 2, total = rows * N;
  int block = 256;
  unsigned grid = static_cast<unsigned>((total + block - 1) / block);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CGQ_DTYPE_F16)
    w4_dequant_kernel<__half><<<grid, block, 0, st>>>(Wq, static_cast<const __half*>(scale),
                                                      static_cast<__half*>(out), rows, N, group);
  else
    w4_dequant_kernel<__nv_bfloat16><<<grid, block, 0, st>>>(
        Wq, static_cast<const __nv_bfloat16*>(scale), static_cast<__nv_bfloat16*>(out), rows, N,
        group);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

extern "C" int cgq_w4_embedding(const int64_t* ids, int n_ids, const uint8_t* Wq,
                                const void* scale, void* out, int V, int D, int group, int dtype,
                                void* stream) {
  if (n_ids < 0 || V <= 0 || D <= 0 || group <= 0 || (group & 1) || V % group != 0) {
    set_error("cgq_w4_embedding: bad shape n=%d V=%d D=%d group=%d", n_ids, V, D, group);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("cgq_w4_embedding: bad dtype %d", dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  if (n_ids == 0) return CGQ_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CGQ_DTYPE_F16)
    w4_embedding_kernel<__half><<<n_ids, 256, 0, st>>>(
        ids, n_ids, Wq, static_cast<const __half*>(scale), static_cast<__half*>(out), D, group);
  else
    w4_embedding_kernel<__nv_bfloat16><<<n_ids, 256, 0, st>>>(
        ids, n_ids, Wq, static_cast<const __nv_bfloat16*>(scale),
        static_cast<__nv_bfloat16*>(out), D, group);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

extern "C" int cgq_w8_embedding(const int64_t* ids, int n_ids, const int8_t* Wq,
                                const void* scale, void* out, int V, int D, int dtype,
                                void* stream) {
  if (n_ids < 0 || V <= 0 || D <= 0) {
    set_error("cgq_w8_embedding: bad shape n=%d V=%d D=%d", n_ids, V, D);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("cgq_w8_embedding: bad dtype %d", dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  if (n_ids == 0) return CGQ_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CGQ_DTYPE_F16)
    w8_embedding_kernel<__half><<<n_ids, 256, 0, st>>>(
        ids, n_ids, Wq, static_cast<const __half*>(scale), static_cast<__half*>(out), D);
  else
    w8_embedding_kernel<__nv_bfloat16><<<n_ids, 256, 0, st>>>(
        ids, n_ids, Wq, static_cast<const __nv_bfloat16*>(scale),
        static_cast<__nv_bfloat16*>(out), D);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}
