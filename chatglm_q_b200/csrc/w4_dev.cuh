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
// one 16-bit element written by the previous kernel of the chain: read past L1
template <typename T>
__device__ __forceinline__ T ldcg_t(const T* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return *reinterpret_cast<T*>(&v);
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
__device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b, __half) { return h2_mul(a, b); }
__device__ __forceinline__ uint32_t mul2(uint32_t a, uint32_t b, __nv_bfloat16) {
  uint32_t r;
  asm("mul.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
template <typename T>
__device__ __forceinline__ uint4 rmsnorm8(const uint4& x, const uint4& w, float rstd) {
  uint4 n;
  const T* xh = reinterpret_cast<const T*>(&x);
  T* nh = reinterpret_cast<T*>(&n);
#pragma unroll
  for (int j = 0; j < 8; ++j) nh[j] = DT<T>::from_f(DT<T>::to_f(xh[j]) * rstd);
  // round_T(float(n) * float(w)): the product of two 16-bit floats is exact in fp32, so the packed 16-bit multiply
  // (one correctly rounded operation) gives the same bits
  return make_uint4(mul2(n.x, w.x, T()), mul2(n.y, w.y, T()), mul2(n.z, w.z, T()), mul2(n.w, w.w, T()));
}
// SwiGLU of 8 elements: round_T(round_T(silu(h)) * gate)  (model.py:200-201, F.silu computes in fp32)
template <typename T>
__device__ __forceinline__ uint4 silu_gate8(const uint4& h, const uint4& g) {
  uint4 a;
  const T* hh = reinterpret_cast<const T*>(&h);
  T* ah = reinterpret_cast<T*>(&a);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float x = DT<T>::to_f(hh[j]);
    // fast exp / divide (a few fp32 ulps from torch's expf + IEEE divide; the result is rounded to T anyway:
    // every CTA of a k-band recomputes its slice, so the accurate versions cost ~1.5 us per launch)
    ah[j] = DT<T>::from_f(__fdividef(x, 1.f + __expf(-x)));
  }
  // round_T(float(act) * float(gate)) == the packed 16-bit multiply (exact product, one rounding)
  return make_uint4(mul2(a.x, g.x, T()), mul2(a.y, g.y, T()), mul2(a.z, g.z, T()), mul2(a.w, g.w, T()));
}
// ---- integer-MMA arithmetic of the one-token kernel (gemv_w4.cu, kImma): the activation as base-256 digits --------
// A 32-k quantisation group of the activation row is scaled by a power of two so that its largest magnitude lands in
// [2^29, 2^30) and every element becomes the integer X = s3 2^24 + s2 2^16 + s1 2^8 + s0 with SIGNED bytes s_i (the B
// operand of IMMA.16832.U8.S8): X + 0x00808080 has the bytes s_i + 128 (s3: as is), one add and one xor after the
// float -> int conversion.  fp16 / bf16 elements down to 2^-19 of the group's maximum are represented EXACTLY, smaller
// ones to 2^-30 of it -- far below the rounding of the fp32 accumulation.  The odd k of a byte pair enter divided by
// 16, because the high nibble is used in place as the unsigned byte 16 q.
constexpr int DIG_STAGE = CW * 4 * 32;        // bytes of digits per k-stage: [group][digit][tig] x (b0, b1)
constexpr int DIG_INFO = CW * 4;              // + one fp32 per group: 2^-shift
constexpr float kMagicF = 12582912.f;         // 1.5 * 2^23: fp32 whose mantissa field is a two's-complement integer
constexpr int kMagicI = 0x4B400000;

// One 16-byte chunk (8 consecutive k, lanes 4i .. 4i+3 of a warp hold one group) -> digits in MMA B-fragment order.
// Must be called by all 32 lanes (shuffles); `store` predicates the shared-memory writes.
template <typename T>
__device__ __forceinline__ void put_digits(uint32_t dig_base, uint32_t info_base, int chunk, bool store, const uint4& v) {
  constexpr bool kHalf = (DT<T>::code == CGQ_DTYPE_F16);
  // group maximum on the 16-bit patterns (sign cleared: the integer order is the magnitude order, NaN / inf on top)
  uint32_t mx = ptx::max_u16x2(ptx::max_u16x2(v.x & 0x7FFF7FFFu, v.y & 0x7FFF7FFFu),
                               ptx::max_u16x2(v.z & 0x7FFF7FFFu, v.w & 0x7FFF7FFFu));
  mx = max(mx & 0xFFFFu, mx >> 16);
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  // biased fp32 exponent of the maximum's binade
  uint32_t mb = kHalf ? max(mx >> 10, 1u) + 112u : max(mx >> 7, 30u);
  const bool bad = kHalf ? (mx >> 10) == 31u : (mx >> 7) == 255u;     // inf / NaN in the group: poison its sums
  if (bad) mb = 127u;
  const float sc = __uint_as_float((283u - mb) << 23);                 // max * sc in [2^29, 2^30)
  const float sc16 = sc * 0.0625f;
  const T* h = reinterpret_cast<const T*>(&v);
  uint32_t z[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int x;
    asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(x) : "f"(DT<T>::to_f(h[i]) * ((i & 1) ? sc16 : sc)));   // exact scaling
    z[i] = (static_cast<uint32_t>(x) + 0x00808080u) ^ 0x00808080u;     // bytes = signed digits, least significant first
  }
  if (store) {
    const int st = chunk >> 4, grp = (chunk >> 2) & 3, tg = chunk & 3;
    const uint32_t dst = dig_base + st * DIG_STAGE + grp * 128 + tg * 8;
    // 4 x 4 byte transposes: even k -> b0 words, odd k -> b1 words, one word per digit
    const uint32_t e01a = __byte_perm(z[0], z[2], 0x5140), e23a = __byte_perm(z[4], z[6], 0x5140);
    const uint32_t e01b = __byte_perm(z[0], z[2], 0x7362), e23b = __byte_perm(z[4], z[6], 0x7362);
    const uint32_t o01a = __byte_perm(z[1], z[3], 0x5140), o23a = __byte_perm(z[5], z[7], 0x5140);
    const uint32_t o01b = __byte_perm(z[1], z[3], 0x7362), o23b = __byte_perm(z[5], z[7], 0x7362);
    ptx::sts64(dst + 0, __byte_perm(e01a, e23a, 0x5410), __byte_perm(o01a, o23a, 0x5410));
    ptx::sts64(dst + 32, __byte_perm(e01a, e23a, 0x7632), __byte_perm(o01a, o23a, 0x7632));
    ptx::sts64(dst + 64, __byte_perm(e01b, e23b, 0x5410), __byte_perm(o01b, o23b, 0x5410));
    ptx::sts64(dst + 96, __byte_perm(e01b, e23b, 0x7632), __byte_perm(o01b, o23b, 0x7632));
    if (tg == 0) ptx::sts32(info_base + st * DIG_INFO + grp * 4, bad ? 0x7FC00000u : (mb - 29u) << 23);
  }
}

template <typename T>
__device__ __forceinline__ T add_resid(T c, const T* resid, int n) {
  return resid == nullptr ? c : DT<T>::from_f(DT<T>::to_f(resid[n]) + DT<T>::to_f(c));
}

}  // namespace w4
}  // namespace cgq
