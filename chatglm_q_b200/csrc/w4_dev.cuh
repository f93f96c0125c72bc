// Device pieces shared by the int4g32 decode kernels (gemv_w4.cu: one launch per linear;
// decode_program.cu: one persistent launch per token step): tile geometry, nibble -> MMA-operand
// conversions, the fused activation prologues.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace cgq {
namespace w4 {

constexpr int BN = 128;            // columns per tile (TMA inner box, bytes)
constexpr int CW = 4;              // consumer warps == quantisation groups per stage
constexpr int ROWS = 16 * CW;      // packed byte rows per stage
constexpr int KSTAGE = 32 * CW;    // k values per stage
constexpr int MMAX = 8;            // token rows (MMA n)
constexpr int W_BYTES = ROWS * BN;
constexpr int S_BYTES = CW * BN * 2;
constexpr int kThreads = (CW + 1) * 32;

__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t bf2_sub(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("sub.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// v = [byte(r), x, byte(r'), x]  ->  packed pair of the low / high nibbles as T values.
template <typename T, bool kTrick>
struct Nib;
template <>
struct Nib<__half, false> {  // exact q-8 through the 1024+q magic number
  __device__ static __forceinline__ uint32_t lo(uint32_t v) {
    return h2_sub(ptx::and_or(v, 0x000F000Fu, 0x64006400u), 0x64086408u);  // (1024+q) - 1032
  }
  __device__ static __forceinline__ uint32_t hi(uint32_t v) {
    // (1024 + 16q) / 16 - 72
    return h2_fma(ptx::and_or(v, 0x00F000F0u, 0x64006400u), 0x2C002C00u, 0xD480D480u);
  }
};
template <>
struct Nib<__half, true> {  // fp16 subnormals: q * 2^-24 and q * 2^-20
  __device__ static __forceinline__ uint32_t lo(uint32_t v) { return v & 0x000F000Fu; }
  __device__ static __forceinline__ uint32_t hi(uint32_t v) { return v & 0x00F000F0u; }
};
template <>
struct Nib<__nv_bfloat16, false> {  // 128+q magic number, mantissa has 7 bits: shift the high nibble down
  __device__ static __forceinline__ uint32_t lo(uint32_t v) {
    return bf2_sub(ptx::and_or(v, 0x000F000Fu, 0x43004300u), 0x43084308u);  // (128+q) - 136
  }
  __device__ static __forceinline__ uint32_t hi(uint32_t v) {
    return bf2_sub(ptx::and_or(v >> 4, 0x000F000Fu, 0x43004300u), 0x43084308u);
  }
};

enum { PRO_NONE = 0, PRO_RMSNORM = 1, PRO_SILU_GATE = 2 };

__device__ __forceinline__ uint4 ldcg128(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldnc128(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
template <typename T>
__device__ __forceinline__ float sumsq8(const uint4& v) {
  const T* h = reinterpret_cast<const T*>(&v);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float f = DT<T>::to_f(h[j]);
    s = fmaf(f, f, s);
  }
  return s;
}
// RMSNorm of 8 elements, the reference's roundings: round_T(x * rstd) then round_T(. * w)
// (chatglm_q/model.py:68-73: `_norm(x.float()).type_as(x)`, then `output * self.weight`).
template <typename T>
__device__ __forceinline__ uint4 rmsnorm8(const uint4& x, const uint4& w, float rstd) {
  uint4 o;
  const T* xh = reinterpret_cast<const T*>(&x);
  const T* wh = reinterpret_cast<const T*>(&w);
  T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const T n = DT<T>::from_f(DT<T>::to_f(xh[j]) * rstd);
    oh[j] = DT<T>::from_f(DT<T>::to_f(n) * DT<T>::to_f(wh[j]));
  }
  return o;
}
// SwiGLU of 8 elements: round_T(round_T(silu(h)) * gate)  (model.py:200-201, F.silu computes in fp32)
template <typename T>
__device__ __forceinline__ uint4 silu_gate8(const uint4& h, const uint4& g) {
  uint4 o;
  const T* hh = reinterpret_cast<const T*>(&h);
  const T* gh = reinterpret_cast<const T*>(&g);
  T* oh = reinterpret_cast<T*>(&o);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = DT<T>::to_f(hh[j]);
    // fast exp / divide (a few fp32 ulps from torch's expf + IEEE divide; the result is rounded to T anyway:
    // every CTA of a k-band recomputes its slice, so the accurate versions cost ~1.5 us per launch)
    const T act = DT<T>::from_f(__fdividef(x, 1.f + __expf(-x)));
    oh[j] = DT<T>::from_f(DT<T>::to_f(act) * DT<T>::to_f(gh[j]));
  }
  return o;
}
template <typename T>
__device__ __forceinline__ T add_resid(T c, const T* resid, int n) {
  return resid == nullptr ? c : DT<T>::from_f(DT<T>::to_f(resid[n]) + DT<T>::to_f(c));
}

}  // namespace w4
}  // namespace cgq
