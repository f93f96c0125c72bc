// int4g32 batch-1 decode kernel (M == 1): C[1,N] = a[K] · ((nib(Wq) - 8) * scale).
// The single-token specialisation of gemv_w4.cu (same (tile x k-band) cluster decomposition, TMA
// ring, DSMEM band reduction); what changes is the arithmetic core, because at M = 1 the f16 MMA
// path is bound by the ALU pipe (one PRMT/LOP3 per two weights) long before HBM:
//
//   * FP8 tensor-core MMA with EXACT operands.  A nibble q in the low 4 bits of a byte IS the
//     e4m3 number q·2^-9 (0000qqqq: subnormals 0..7·2^-9, then exponent 1: (8+m)·2^-9), so
//     `word & 0x0F0F0F0F` and `(word >> 4) & 0x0F0F0F0F` turn four packed bytes into eight MMA-ready
//     weights: 3 ALU instructions per 8 weights instead of 6, and m16n8k32 e4m3 issues twice as
//     fast as m16n8k16 f16 on sm_100a (measured 4.3 vs 8.1 cycles per SMSP).
//   * the bytes of four consecutive packed rows of one column — what one 32-bit A-fragment
//     register needs — come straight from `ldmatrix.m16n16.trans.b8` on the swizzled TMA tile.
//   * the fp16/bf16 activation is split EXACTLY into three e4m3 terms per element
//     (x·2^e = t0 + t1/16 + t2/256, each 4 significant bits, e chosen per 32-k group so that
//     max|x·2^e| is in [128, 256)); the three terms occupy three of the eight otherwise idle token
//     columns of the MMA.  Products and sums are exact in the fp32 accumulator, so the result equals
//     Σ_g s_g Σ_k a_k (q_k - 8) to fp32 rounding — the same value the f16 kernel computes.
//     The split is done once per CTA for its whole k-band, before the main loop.
//   * the -8 offset is -8·Σ_{k∈g} a_k (fp32, from the same pre-pass), the group scale multiplies
//     the group's partial sum, epilogue = reference's two roundings (int4/qlinear.py:91-93).
#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace cgq {
namespace {

constexpr int BN = 128;            // columns per tile
constexpr int CW = 4;              // consumer warps == groups per stage
constexpr int ROWS = 16 * CW;      // packed rows per stage
constexpr int KSTAGE = 32 * CW;    // k per stage
constexpr int W_BYTES = ROWS * BN;
constexpr int S_BYTES = CW * BN * 2;
constexpr int STAGE_BYTES = W_BYTES + S_BYTES;
constexpr int kThreads = (CW + 1) * 32;
constexpr int BQ_STAGE = 3 * CW * 32;   // e4m3 B fragments of one stage: [term][group][tig] x 8 B
constexpr int GI_STAGE = CW * 8;        // per group: (512 / 2^e, 8 * Σ x) as two floats
constexpr int RED_BYTES = CW * BN * 4;
constexpr int XRED_BYTES = 8 * BN * 4;

struct Params {
  const void* A;
  const void* bias;
  void* C;
  int N, K;
  int SPT, Z, S;
  int max_units;   // k-stages per CTA (upper bound, sizes the activation pre-pass buffers)
  int xred_bytes;  // 0 when Z == 1
  unsigned long long* trace;
};

__device__ __forceinline__ void stamp(const Params& p, int slot) {
  if (p.trace != nullptr && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + slot] = t;
  }
}

__device__ __forceinline__ void qmma_16832(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%10,%10,%10,%10};"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void ldsm_x2_trans_b8(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// two f16 (packed) -> two e4m3 (low half -> low byte), round-to-nearest, saturating
__device__ __forceinline__ uint32_t f16x2_to_e4m3x2(uint32_t h2) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
  return r;
}
__device__ __forceinline__ uint32_t e4m3x2_to_f16x2(uint32_t e2) {
  uint32_t r;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(r) : "h"(static_cast<uint16_t>(e2)));
  return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
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
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
    w4_gemv_m1_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmS,
                      const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int S = p.S;
  const uint32_t Wsm = base;
  const uint32_t Ssm = Wsm + S * W_BYTES;
  const uint32_t Bq = Ssm + S * S_BYTES;                    // [max_units][3][CW][4] x 8 B
  const uint32_t Gi = Bq + p.max_units * BQ_STAGE;          // [max_units][CW] x (float, float)
  const uint32_t off_red = S * STAGE_BYTES + p.max_units * (BQ_STAGE + GI_STAGE);
  float* red = reinterpret_cast<float*>(gen + off_red);
  float* xred = reinterpret_cast<float*>(gen + off_red + RED_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + RED_BYTES + p.xred_bytes);
  uint64_t* empty = full + S;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Z = p.Z;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;
  const int n_units = u1 - u0;

  if (threadIdx.x == 0) stamp(p, 0);
  if (threadIdx.x == CW * 32) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmS);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  ptx::pdl_launch_dependents();   // the next kernel may start prefetching its weights

  const uint64_t pol = ptx::policy_evict_first();
  auto issue_w = [&](int i, int slot) {
    const int ks = u0 + i;
    ptx::mbar_expect_tx(&full[slot], W_BYTES + S_BYTES);
    ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, tile * BN, ks * ROWS, &full[slot], pol);
    ptx::tma_load_2d(gen + S * W_BYTES + slot * S_BYTES, &tmS, tile * BN, ks * CW, &full[slot], pol);
  };
  const int prefill = min(n_units, S);
  // weights do not depend on the previous kernel: stream them before the PDL wait
  if (threadIdx.x == CW * 32)
    for (int i = 0; i < prefill; ++i) issue_w(i, i);
  ptx::pdl_wait_prior_grid();
  if (threadIdx.x == 0) stamp(p, 2);

  // ---------------- activation pre-pass: exact 3-term e4m3 split of this CTA's k-band
  // one warp-pass = 128 k (one stage): 8 lanes per 32-k group, 4 consecutive k per lane
  {
    const T* A = static_cast<const T*>(p.A);
    for (int st = warp; st < n_units; st += CW + 1) {
      const int k = (u0 + st) * KSTAGE + lane * 4;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (k < p.K) {   // K % 32 == 0: a lane's 4 values are all valid or all past the end
        const uint2 raw = *reinterpret_cast<const uint2*>(A + k);
        union {
          uint2 u;
          T h[4];
        } cv;
        cv.u = raw;
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = DT<T>::to_f(cv.h[i]);
      }
      float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
      m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
      const uint32_t mb = __float_as_uint(m) >> 23;                     // biased exponent of the group max
      const bool ok = (mb >= 8u) && (mb < 255u);                        // zero / vanishing group (or inf/nan): scale 1
      const float sc = ok ? __uint_as_float((261u - mb) << 23) : 1.f;   // 2^(7 - e): max*sc in [128, 256)
      const float inv = ok ? __uint_as_float((mb - 7u) << 23) : 1.f;    // 1 / sc
      float xs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = x[i] * sc;                    // exact
      float gs = (xs[0] + xs[1]) + (xs[2] + xs[3]);
      gs += __shfl_xor_sync(0xffffffffu, gs, 1);
      gs += __shfl_xor_sync(0xffffffffu, gs, 2);
      gs += __shfl_xor_sync(0xffffffffu, gs, 4);
      const uint32_t v01 = pack_f16x2(xs[0], xs[1]), v23 = pack_f16x2(xs[2], xs[3]);  // exact (|xs| < 256)
      const uint32_t k16 = 0x4C004C00u;  // (16, 16)
      // term 0
      const uint32_t a01 = f16x2_to_e4m3x2(v01), a23 = f16x2_to_e4m3x2(v23);
      const uint32_t r01 = h2_mul(h2_sub(v01, e4m3x2_to_f16x2(a01)), k16);   // exact residual * 16
      const uint32_t r23 = h2_mul(h2_sub(v23, e4m3x2_to_f16x2(a23)), k16);
      // term 1
      const uint32_t b01 = f16x2_to_e4m3x2(r01), b23 = f16x2_to_e4m3x2(r23);
      const uint32_t q01 = h2_mul(h2_sub(r01, e4m3x2_to_f16x2(b01)), k16);
      const uint32_t q23 = h2_mul(h2_sub(r23, e4m3x2_to_f16x2(b23)), k16);
      // term 2
      const uint32_t c01 = f16x2_to_e4m3x2(q01), c23 = f16x2_to_e4m3x2(q23);
      // lane L of the group (0..7): tig = L / 2, byte pair (L & 1) of b0 (even k) and b1 (odd k)
      const int grp = lane >> 3, L = lane & 7;
      const uint32_t dst = Bq + st * BQ_STAGE + grp * 32 + (L >> 1) * 8 + (L & 1) * 2;
      const uint32_t terms01[3] = {a01, b01, c01}, terms23[3] = {a23, b23, c23};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint32_t w = terms01[t] | (terms23[t] << 16);             // bytes: k0 k1 k2 k3
        sts16(dst + t * (CW * 32), __byte_perm(w, 0, 0x4420));          // b0: (k0, k2)
        sts16(dst + t * (CW * 32) + 4, __byte_perm(w, 0, 0x4431));      // b1: (k1, k3)
      }
      if (L == 0) {
        float2 gi;
        gi.x = inv * 512.f;        // weights enter as q * 2^-9
        gi.y = 8.f * inv * gs;     // 8 * Σ_k a_k of the group
        *reinterpret_cast<float2*>(gen + (Gi - base) + (st * CW + grp) * 8) = gi;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) stamp(p, 1);

  if (warp == CW) {
    // =========================== producer: one lane drives TMA ===========================
    if (lane == 0) {
      int slot = 0, phase = 1;
      for (int i = prefill; i < n_units; ++i) {
        ptx::mbar_wait(&empty[slot], phase ^ 1);
        issue_w(i, slot);
        if (++slot == S) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // =========================== consumers ===========================
    const int g = lane >> 2, tig = lane & 3;
    float tot[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) tot[j][0] = tot[j][1] = 0.f;
    // which term(s) this lane's accumulator columns carry: tig 0 -> terms 0,1; tig 1 -> term 2
    const float cf0 = (tig == 0) ? 1.f : (tig == 1) ? (1.f / 256.f) : 0.f;
    const float cf1 = (tig == 0) ? (1.f / 16.f) : 0.f;
    const float co = (tig == 0) ? 1.f : 0.f;
    // ldmatrix row address of this lane: packed row (16 warp + (lane & 15)), 16-byte chunk index
    // 2 jj + (lane >> 4), 128-byte swizzle (chunk ^= row & 7)
    const int lrow = 16 * warp + (lane & 15);
    const uint32_t ld_off = lrow * BN;
    const uint32_t ld_sw = lrow & 7, ld_hi = lane >> 4;

    int slot = 0, phase = 0;
    for (int it = 0; it < n_units; ++it) {
      // B fragment (e4m3 terms of this group's activations): token column g < 3 <-> term g
      uint32_t b0 = 0, b1 = 0;
      if (g < 3) {
        const uint2 bv = ptx::lds64(Bq + it * BQ_STAGE + g * (CW * 32) + warp * 32 + tig * 8);
        b0 = bv.x;
        b1 = bv.y;
      }
      const uint2 giu = ptx::lds64(Gi + (it * CW + warp) * 8);
      const float I = __uint_as_float(giu.x), O = __uint_as_float(giu.y);
      const float f0 = cf0 * I, f1 = cf1 * I, off = co * O;

      ptx::mbar_wait(&full[slot], phase);
      if (it == 0 && threadIdx.x == 0) stamp(p, 3);
      const uint32_t wbase = Wsm + slot * W_BYTES + ld_off;
      const uint32_t sbase = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 2;
      float d[8][4];
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {
        uint32_t r[4];
        ldsm_x2_trans_b8(wbase + (((2 * jp + ld_hi) ^ ld_sw) << 4), r);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t x0 = r[2 * h], x1 = r[2 * h + 1];   // column 16 jj + g / + g + 8, packed rows 4 tig..+3
          const uint32_t a[4] = {x0 & 0x0F0F0F0Fu, x1 & 0x0F0F0F0Fu, (x0 >> 4) & 0x0F0F0F0Fu,
                                 (x1 >> 4) & 0x0F0F0F0Fu};
          qmma_16832(d[2 * jp + h], a, b0, b1);
        }
      }
      uint32_t sraw[16];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        sraw[2 * jj] = lds16(sbase + (16 * jj) * 2);
        sraw[2 * jj + 1] = lds16(sbase + (16 * jj + 8) * 2);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&empty[slot]);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        union {
          uint16_t u;
          T h;
        } ca, cb;
        ca.u = static_cast<uint16_t>(sraw[2 * jj]);
        cb.u = static_cast<uint16_t>(sraw[2 * jj + 1]);
        const float sa = DT<T>::to_f(ca.h), sb = DT<T>::to_f(cb.h);
        const float ua = fmaf(f1, d[jj][1], fmaf(f0, d[jj][0], -off));
        const float ub = fmaf(f1, d[jj][3], fmaf(f0, d[jj][2], -off));
        tot[jj][0] = fmaf(sa, ua, tot[jj][0]);
        tot[jj][1] = fmaf(sb, ub, tot[jj][1]);
      }
      if (++slot == S) {
        slot = 0;
        phase ^= 1;
      }
    }
    if (threadIdx.x == 0) stamp(p, 4);

    // ---------------- band sum of this CTA: terms live in lanes tig 0/1 -> add across tig, then warps
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = tot[jj][h];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (tig == 0) red[warp * BN + 16 * jj + g + 8 * h] = v;
      }
    }
    ptx::named_bar_sync(1, CW * 32);
    {
      const int t = threadIdx.x;
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < CW; ++w) v += red[w * BN + t];
      if (Z == 1) {
        const int n = tile * BN + t;
        if (n < p.N) static_cast<T*>(p.C)[n] = epilogue<T>(v, static_cast<const T*>(p.bias), n);
      } else {
        const uint32_t local = ptx::smem_u32(xred) + static_cast<uint32_t>(z * BN + t) * 4u;
        ptx::st_cluster_f32(ptx::mapa_rank(local, 0), v);
      }
    }
  }
  if (Z > 1) {
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
    if (z == 0 && threadIdx.x < BN) {
      const int t = threadIdx.x, n = tile * BN + t;
      if (n < p.N) {
        float acc = 0.f;
        for (int zz = 0; zz < Z; ++zz) acc += xred[zz * BN + t];
        static_cast<T*>(p.C)[n] = epilogue<T>(acc, static_cast<const T*>(p.bias), n);
      }
    }
  }
  if (threadIdx.x == 0) stamp(p, 5);
}

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* s = getenv(name);
  if (s == nullptr || *s == 0) return dflt;
  int v = atoi(s);
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

template <typename T>
int launch_t(const GemmArgs& a, bool* taken) {
  const int G = a.K / 32;
  const int SPT = (G + CW - 1) / CW;
  const int tiles = (a.N + BN - 1) / BN;
  static const int stages_env = env_int("CGQ_GEMV_STAGES", 0, 0, 16);
  static const int z_env = env_int("CGQ_GEMV_Z", 0, 0, 8);
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 4, 1, 4);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  const int slots = cps * sm_count();
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2) Z *= 2;
  if (z_env > 0) Z = z_env;
  if (Z > SPT) Z = 1;
  const int grid = tiles * Z;
  int stages = stages_env > 0 ? stages_env : (grid * 4 <= slots * 3 ? 6 : 4);
  const int per_cta = (SPT + Z - 1) / Z;
  if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
  const int xred_bytes = Z > 1 ? XRED_BYTES : 0;
  const size_t smem = 1024 + static_cast<size_t>(stages) * STAGE_BYTES +
                      static_cast<size_t>(per_cta) * (BQ_STAGE + GI_STAGE) + RED_BYTES + xred_bytes +
                      16 * stages + 16;
  if (smem > 100 * 1024) {   // very long k-band: the general decode kernel takes it
    *taken = false;
    return CGQ_OK;
  }
  *taken = true;

  CUtensorMap tmW, tmS;
  TmapKey kw{a.Wq, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K / 2),
             static_cast<uint64_t>(a.N), BN, ROWS, CU_TENSOR_MAP_DATA_TYPE_UINT8,
             CU_TENSOR_MAP_SWIZZLE_128B};
  int rc = get_tmap_2d(kw, &tmW);
  if (rc != CGQ_OK) return rc;
  TmapKey ks{a.scale, static_cast<uint64_t>(a.N), static_cast<uint64_t>(G),
             static_cast<uint64_t>(a.N) * 2, BN, CW,
             a.dtype == CGQ_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                      : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
             CU_TENSOR_MAP_SWIZZLE_NONE};
  rc = get_tmap_2d(ks, &tmS);
  if (rc != CGQ_OK) return rc;

  Params prm;
  prm.A = a.A;
  prm.bias = a.bias;
  prm.C = a.C;
  prm.N = a.N;
  prm.K = a.K;
  prm.SPT = SPT;
  prm.Z = Z;
  prm.S = stages;
  prm.max_units = per_cta;
  prm.xred_bytes = xred_bytes;
  prm.trace = static_cast<unsigned long long*>(take_trace_buffer());

  auto kern = w4_gemv_m1_kernel<T>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && smem > configured[dev]) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    configured[dev] = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (Z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(Z);
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmW, tmS, prm));
  return CGQ_OK;
}

}  // namespace

// M == 1 fast path; *taken = false when the shape is left to the general decode kernel.
int launch_w4_gemv_m1(const GemmArgs& a, bool* taken) {
  static const bool enabled = env_int("CGQ_GEMV_M1", 1, 0, 1) != 0;
  if (!enabled || a.M != 1) {
    *taken = false;
    return CGQ_OK;
  }
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, taken) : launch_t<__nv_bfloat16>(a, taken);
}

}  // namespace cgq
