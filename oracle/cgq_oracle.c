/*
 * ORACLE — test infrastructure, NOT product code.
 *
 * Plain-C restatement of the reference's torch CPU path for the dequant-matmul, used (a) as a
 * second, independent checker of the numpy oracle and (b) as the timed CPU baseline of bench.py
 * ("port": same algorithm as the reference runs on CPU tensors — materialise the dequantised
 * [K, N] matrix, then a dense matmul — multi-threaded with OpenMP like ATen's kernels).
 *
 *   int4: chatglm_q/int4/qlinear.py:20-33 (unpack_int4) and :50 (A.matmul(unpack_int4(B, scale)))
 *   int8: chatglm_q/int8/qlinear.py:38    (A.matmul(B * b_scale))
 *   bias: chatglm_q/int4/qlinear.py:92-93 (out += bias, a second rounded op)
 *
 * All tensors cross this interface as float32 holding values exactly representable in the
 * activation dtype (0 = float16, 1 = bfloat16, 2 = float32); results are rounded to that dtype.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline float round_bf16(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return x; /* NaN */
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
  memcpy(&x, &u, 4);
  return x;
}
static inline float round_dt(float x, int dtype) {
  if (dtype == 0) return (float)(_Float16)x;
  if (dtype == 1) return round_bf16(x);
  return x;
}

/* out[K, N] int8 = nibble - 8 (low nibble = even k) */
void oracle_w4_unpack_i8(const uint8_t* wq, int8_t* out, int K, int N) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < K / 2; ++r)
    for (int n = 0; n < N; ++n) {
      uint8_t b = wq[(size_t)r * N + n];
      out[(size_t)(2 * r) * N + n] = (int8_t)((int)(b & 0xF) - 8);
      out[(size_t)(2 * r + 1) * N + n] = (int8_t)((int)(b >> 4) - 8);
    }
}

/* out[K, N] = round_dt((nibble - 8) * scale[k / group, n]) */
void oracle_w4_dequant(const uint8_t* wq, const float* scale, float* out, int K, int N, int group,
                       int dtype) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < K / 2; ++r) {
    const float* s = scale + (size_t)((2 * r) / group) * N;
    for (int n = 0; n < N; ++n) {
      uint8_t b = wq[(size_t)r * N + n];
      out[(size_t)(2 * r) * N + n] = round_dt((float)((int)(b & 0xF) - 8) * s[n], dtype);
      out[(size_t)(2 * r + 1) * N + n] = round_dt((float)((int)(b >> 4) - 8) * s[n], dtype);
    }
  }
}

/* C[M, N] = round_dt(A[M, K] . W[K, N]) (+ bias, second rounding); fp32 accumulation */
static void dense_matmul(const float* A, const float* W, const float* bias, float* C, int M, int N,
                         int K, int dtype) {
  enum { NB = 512 };
#pragma omp parallel for schedule(static) collapse(2)
  for (int m = 0; m < M; ++m)
    for (int n0 = 0; n0 < N; n0 += NB) {
      float acc[NB];
      const int nb = (N - n0 < NB) ? (N - n0) : NB;
      for (int j = 0; j < nb; ++j) acc[j] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float a = A[(size_t)m * K + k];
        const float* w = W + (size_t)k * N + n0;
        for (int j = 0; j < nb; ++j) acc[j] += a * w[j];
      }
      for (int j = 0; j < nb; ++j) {
        float c = round_dt(acc[j], dtype);
        if (bias) c = round_dt(c + bias[n0 + j], dtype);
        C[(size_t)m * N + n0 + j] = c;
      }
    }
}

/* int4g32 linear forward exactly as the reference's CPU path does it: unpack, then matmul.
 * `scratch` must hold K*N floats (the materialised dequantised weight). */
void oracle_w4a16_gemm(const float* A, const uint8_t* wq, const float* scale, const float* bias,
                       float* C, int M, int N, int K, int group, int dtype, float* scratch) {
  oracle_w4_dequant(wq, scale, scratch, K, N, group, dtype);
  dense_matmul(A, scratch, bias, C, M, N, K, dtype);
}

/* int8 linear forward: W[k, n] = round_dt(wq[n, k] * scale[n]) (the `weight.t() * scale` tensor),
 * then matmul.  wq is the module buffer [N, K]. */
void oracle_w8a16_gemm(const float* A, const int8_t* wq, const float* scale, const float* bias,
                       float* C, int M, int N, int K, int dtype, float* scratch) {
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n)
      scratch[(size_t)k * N + n] = round_dt((float)wq[(size_t)n * K + k] * scale[n], dtype);
  dense_matmul(A, scratch, bias, C, M, N, K, dtype);
}

int oracle_version(void) { return 1; }
