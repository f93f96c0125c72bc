// int4g32 batch-1 decode kernel on the 5th-generation tensor cores (M == 1):
//     C[1,N] = a[K] · ((nib(Wq) - 8) * scale)
// Replaces _dynamic_quant_matmul_s4_kernel (chatglm_q/int4/triton_ops.py:18-87) for the batch-1
// decode token.  Same work split as gemv_w4.cu — (128-column tile) x (Z lockstep k-bands), the Z
// CTAs of a tile are a cluster and reduce through distributed shared memory — but the arithmetic
// core is INTEGER tcgen05, because with mma.sync the decode kernel is bound by the CUDA-core
// unpack + per-fragment scale work (measured: 4.85 TB/s with the maths, 6.1 TB/s without it):
//
//   * a nibble masked in place, `word & 0x0F0F0F0F` / `(word >> 4) & 0x0F0F0F0F`, IS the int8 weight
//     q (0..15): 3 ALU instructions per 8 weights turn the packed tile [16 rows x 128 B] of one
//     32-k group into the int8 MN-major UMMA A operand [32 k x 128 n] (128-byte swizzle) — the
//     columns stay contiguous exactly as they are in memory, nothing is transposed or converted;
//   * the fp16 / bf16 activation is split into three int8 digits per element,
//     x·2^(5-e) = t0 + t1/128 + t2/128^2 (e = exponent of the group's max, |t| <= 64): 21 bits
//     below the group maximum, i.e. every fp16 element within 9 binades of the maximum is exact
//     and the rest is truncated at 2^-20 of it.  The digits of a group are the K-major B operand
//     (3 of the 16 UMMA-N columns);
//   * one `tcgen05.mma.cta_group::1.kind::i8` (M=128 columns, N=16, K=32) per quantisation group
//     computes Σ_k q_k·t_k exactly in int32 in TMEM; the epilogue thread that owns the column
//     recombines the digits in integer arithmetic, subtracts 8·Σ_k T_k (the zero point, exact),
//     converts once and applies the group scale in fp32:
//         acc_n += s[g,n] · 2^(e-19) · Σ_k (q_k - 8)·T_k
//     which is Σ_g s_g Σ_k a_k (q_k-8) — the value the mma.sync kernel computes — to fp32 rounding.
//
// The int8 A operand never touches shared memory: `ldmatrix.m16n16.trans.b8` hands a thread four
// consecutive packed rows of one weight column, two masks turn that register into the column's
// (even-k, odd-k) int8 words, and `tcgen05.st.16x256b` — whose register layout is the same
// (row = lane/4, column pair = lane%4) fragment — stores them as the column's TMEM row.  The MMA reads
// A from TMEM, B (the digits, 2 KB per stage) from shared memory; the B k-order is permuted to match.
// Shared-memory traffic per 9 KB stage: TMA write + one ldmatrix read + ~3 KB of scales / digits.
//
// STATUS (B200, round 1): parity-tested (tests/test_gpu_parity.py::test_int4_decode_umma), 1 028 us per
// ChatGLM2-6B token against 1 012 us for the mma.sync kernel (gemv_w4.cu), so it stays opt-in
// (CGQ_IMPL_GEMV_UMMA / env CGQ_GEMV_UMMA=1) until the stage hand-off latency is tuned.
//
// Warp roles (384 threads): warps 0-3 epilogue (TMEM lane quarters), 4-7 unpack (32 weight columns ==
// one TMEM lane quarter each, all four groups of a stage), 8 TMA producer, 9 MMA issuer + TMEM
// allocator, 10-11 activation digits.  All hand-offs are per 128-k stage (4 groups): TMA ring ->
// {unpack, digits} -> 4 MMAs + commit -> epilogue, over a ring of NS stage-slots (32 TMEM columns of
// A + 32 of D + 2 KB digits + 1 KB scales each).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace cgq {
namespace {

constexpr int BN = 128;            // columns per tile (UMMA M)
constexpr int GPS = 4;             // groups per TMA stage
constexpr int ROWS = 16 * GPS;     // packed rows per stage
constexpr int KSTAGE = 32 * GPS;   // k per stage
constexpr int W_BYTES = ROWS * BN;          // 8192
constexpr int S_BYTES = GPS * BN * 2;       // 1024
constexpr int STAGE_BYTES = W_BYTES + S_BYTES;
constexpr int NU = 16;                      // UMMA N (3 digit columns used)
constexpr int B_SLOT = NU * 32;             // 512: int8 [16 n x 32 k], K-major, no swizzle
constexpr int ACOLS = 8;                    // TMEM columns of one group's A: 32 int8 per lane
constexpr int DCOLS = 4;                    // TMEM column stride of one group's D (3 digit sums used;
                                            // the 16-wide MMA outputs overlap, later groups only clobber unused columns)
constexpr int SLOT_COLS = 32;               // per stage-slot: 32 columns of D, 32 of A
constexpr int B_STAGE = GPS * B_SLOT;           // 2048
constexpr int GI_STAGE = GPS * 8;               // (float I, int 8*ΣT) per group
constexpr int SC_STAGE = GPS * BN * 2;          // 1024: the stage's group scales, copied out of the TMA ring
constexpr int kThreads = 12 * 32;
constexpr int XRED_BYTES = 8 * BN * 4;

struct Params {
  const void* A;
  const void* bias;
  void* C;
  int N, K;
  int SPT, Z, S, NA;   // stages per tile, bands, TMA ring depth, group-slot ring depth
  int max_units;       // k-stages per CTA (sizes the activation band buffer)
  int xred_bytes;
  uint32_t idesc;
  unsigned long long* trace;
};

__device__ __forceinline__ void stamp(const Params& p, int slot) {
  if (p.trace != nullptr && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + slot] = t;
  }
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_noswz(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_i8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 0, 0;\n\t"   // never accumulate: one group per MMA
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc)
      : "memory");
}
// D[tmem] = A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_i8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, 0, 0;\n\t"   // never accumulate: one group per MMA
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc)
      : "memory");
}
// 16 lanes x 256 bit: thread (g = lane/4, t = lane%4) -> r0,r1 = (lane g, columns 2t, 2t+1), r2,r3 = lane g+8
__device__ __forceinline__ void tmem_st_16x256b(uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1),
               "r"(r2), "r"(r3)
               : "memory");
}
__device__ __forceinline__ void ldsm_x2_trans_b8(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(static_cast<uint16_t>(v)) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
  return v;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
    w4_gemv_umma_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmS,
                        const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int S = p.S, NS = p.NA;
  // layout: TMA stages {packed (1024-aligned, 128 B swizzle) | scales} | B stage-slots | group info |
  //         scales | xred | barriers | tmem ptr
  const uint32_t Wsm = base;
  const uint32_t Bsl = Wsm + S * STAGE_BYTES;
  const uint32_t Gi = Bsl + NS * B_STAGE;                 // [NS][GPS] x (float I, int G8)
  const uint32_t Ssl = Gi + NS * GI_STAGE;                // [NS][GPS][128] scales
  const uint32_t off_x = (Ssl - base) + NS * SC_STAGE;
  float* xred = reinterpret_cast<float*>(gen + off_x);
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + off_x + p.xred_bytes);
  uint64_t* full_tma = bars;               // [S]   TMA landed
  uint64_t* empty_tma = full_tma + S;      // [S]   4 unpack warps have the stage in registers / copied
  uint64_t* ab_full = empty_tma + S;       // [NS]  int8 A in TMEM + scales (4 unpack warps) and digits (1 warp) written
  uint64_t* mma_done = ab_full + NS;       // [NS]  tcgen05.commit: D ready
  uint64_t* d_empty = mma_done + NS;       // [NS]  4 epilogue warps have read D / scales / group info: slot free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + NS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Z = p.Z;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;
  const int n_units = u1 - u0;
  const uint32_t need_cols = static_cast<uint32_t>(NS * 2 * SLOT_COLS);
  const uint32_t tmem_cols = need_cols <= 32 ? 32u : need_cols <= 64 ? 64u : need_cols <= 128 ? 128u
                             : need_cols <= 256 ? 256u : 512u;

  if (threadIdx.x == 0) stamp(p, 0);
  if (threadIdx.x == 8 * 32) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmS);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full_tma[s], 1);
      ptx::mbar_init(&empty_tma[s], GPS);
    }
    for (int a = 0; a < NS; ++a) {
      ptx::mbar_init(&ab_full[a], GPS + 1);
      ptx::mbar_init(&mma_done[a], 1);
      ptx::mbar_init(&d_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 9) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  // unused B rows (digit columns 3..15) stay zero for the whole kernel
  for (int i = threadIdx.x; i < NS * B_STAGE / 16; i += kThreads)
    ptx::sts128(Bsl + i * 16, make_uint4(0, 0, 0, 0));
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();   // the next kernel may start prefetching its weights

  if (warp == 8) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      int s = 0, ph = 0;
      for (int i = 0; i < n_units; ++i) {   // weights do not depend on the previous kernel
        ptx::mbar_wait(&empty_tma[s], ph ^ 1);
        const int ks = u0 + i;
        uint8_t* st = gen + (Wsm - base) + s * STAGE_BYTES;
        ptx::mbar_expect_tx(&full_tma[s], W_BYTES + S_BYTES);
        ptx::tma_load_2d(st, &tmW, tile * BN, ks * ROWS, &full_tma[s], pol);
        ptx::tma_load_2d(st + W_BYTES, &tmS, tile * BN, ks * GPS, &full_tma[s], pol);
        if (++s == S) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp >= 10) {
    // =========================== activation digits: warp 10 even stages, warp 11 odd ===========================
    ptx::pdl_wait_prior_grid();
    if (threadIdx.x == 10 * 32) stamp(p, 2);
    const T* A = static_cast<const T*>(p.A);
    const int dw = warp - 10;
    // lane -> 4 consecutive k of the stage; 8 lanes per 32-k group
    const int grp = lane >> 3, L = lane & 7;
    // the activation slice of a stage is 8 B per lane from L2 (~0.6 us away): keep PF stages in flight
    constexpr int PF = 4;
    uint2 pre[PF];
    auto fetch = [&](int st) -> uint2 {
      const int k = (u0 + st) * KSTAGE + lane * 4;
      return (st < n_units && k < p.K) ? *reinterpret_cast<const uint2*>(A + k) : make_uint2(0u, 0u);
    };
#pragma unroll
    for (int i = 0; i < PF; ++i) pre[i] = fetch(dw + 2 * i);
    for (int st0 = dw; st0 < n_units; st0 += 2 * PF) {
#pragma unroll
      for (int pi = 0; pi < PF; ++pi) {
        const int st = st0 + 2 * pi;
        if (st >= n_units) break;
        float x[4];
        {
          union {
            uint2 u;
            T h[4];
          } cv;
          cv.u = pre[pi];
          pre[pi] = fetch(st + 2 * PF);
#pragma unroll
          for (int i = 0; i < 4; ++i) x[i] = DT<T>::to_f(cv.h[i]);
        }
        float m = fmaxf(fmaxf(fabsf(x[0]), fabsf(x[1])), fmaxf(fabsf(x[2]), fabsf(x[3])));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        const uint32_t mb = __float_as_uint(m) >> 23;                      // biased exponent of the group max
        const bool ok = (mb >= 24u) && (mb < 255u);                        // zero / vanishing group: scale 1
        const float sc = ok ? __uint_as_float((259u - mb) << 23) : 1.f;    // 2^(5-e): |x*sc| < 64
        const float I = ok ? __uint_as_float((mb - 19u) << 23) : 6.103515625e-05f;  // 2^(e-5-14)
        uint32_t dig[3] = {0, 0, 0};
        int tsum = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float X = x[i] * sc;
          const float t0 = rintf(X);
          const float r1 = (X - t0) * 128.f;
          const float t1 = rintf(r1);
          const float r2 = (r1 - t1) * 128.f;
          const float t2 = rintf(r2);
          const int i0 = __float2int_rn(t0), i1 = __float2int_rn(t1), i2 = __float2int_rn(t2);
          dig[0] |= static_cast<uint32_t>(i0 & 0xFF) << (8 * i);
          dig[1] |= static_cast<uint32_t>(i1 & 0xFF) << (8 * i);
          dig[2] |= static_cast<uint32_t>(i2 & 0xFF) << (8 * i);
          tsum += (i0 * 128 + i1) * 128 + i2;
        }
        tsum += __shfl_xor_sync(0xffffffffu, tsum, 1);
        tsum += __shfl_xor_sync(0xffffffffu, tsum, 2);
        tsum += __shfl_xor_sync(0xffffffffu, tsum, 4);
        const int ss = st % NS, sph = (st / NS) & 1;
        ptx::mbar_wait(&d_empty[ss], sph ^ 1);    // previous user of the slot has been multiplied and read
        // B[n = digit][kk]: core matrix (8 n x 16 B), kk chunk (kk / 16) is 128 B further.  The MMA k
        // order follows the TMEM A words: kk = 8 (k/8) + j for the even k = 8 (k/8) + 2j, kk + 4 for the odd
        const int kk0 = 8 * (L >> 1) + 2 * (L & 1);
        const uint32_t dst = Bsl + ss * B_STAGE + grp * B_SLOT + (kk0 >> 4) * 128 + (kk0 & 15);
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          sts16(dst + t * 16, __byte_perm(dig[t], 0, 0x4420));       // (k0, k2)
          sts16(dst + t * 16 + 4, __byte_perm(dig[t], 0, 0x4431));   // (k1, k3)
        }
        if (L == 0) {
          sts32(Gi + ss * GI_STAGE + grp * 8, __float_as_uint(I));
          sts32(Gi + ss * GI_STAGE + grp * 8 + 4, static_cast<uint32_t>(8 * tsum));
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ab_full[ss]);
      }
    }
  } else if (warp == 9) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // the slot is known to be free (D read, A/B consumed) once its producers have filled it again
      const uint64_t bdesc0 = desc_noswz(Bsl, 128, 256);    // B: int8 K-major core matrices, k chunks 128 B, n groups 256 B
      const uint32_t a_tmem0 = tmem_base + static_cast<uint32_t>(NS * SLOT_COLS);
      int ss = 0, ph = 0;
      for (int st = 0; st < n_units; ++st) {
        ptx::mbar_wait(&ab_full[ss], ph);
        ptx::tc_fence_after();
#pragma unroll
        for (int g = 0; g < GPS; ++g)
          umma_i8_ts(tmem_base + static_cast<uint32_t>(ss * SLOT_COLS + g * DCOLS),
                     a_tmem0 + static_cast<uint32_t>(ss * SLOT_COLS + g * ACOLS),
                     bdesc0 + static_cast<uint64_t>((ss * B_STAGE + g * B_SLOT) >> 4), p.idesc);
        ptx::umma_commit(&mma_done[ss]);
        if (++ss == NS) {
          ss = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // =========================== unpack warps: 32 weight columns (one TMEM lane quarter) each ===========================
    const int q = warp - 4;
    const int g8 = lane >> 2;      // fragment row (weight column within a 16-column block)
    (void)g8;
    int s = 0, ph = 0, ss = 0, sph = 0;
    // ldmatrix row address: packed row (16 g + (lane & 15)) of the stage, 16-byte chunk 2q + (lane >> 4),
    // 128-byte swizzle (chunk ^= row & 7; 16 g does not change row & 7)
    const uint32_t ld_row = (lane & 15) * BN + (((2 * q + (lane >> 4)) ^ (lane & 7)) << 4);
    const uint32_t a_lane = static_cast<uint32_t>(32 * q) << 16;
    for (int st = 0; st < n_units; ++st) {
      ptx::mbar_wait(&full_tma[s], ph);
      const uint32_t src = Wsm + s * STAGE_BYTES + ld_row;
      uint32_t r[GPS][4];
#pragma unroll
      for (int g = 0; g < GPS; ++g) ldsm_x2_trans_b8(src + g * 16 * BN, r[g]);
      uint4 sc = make_uint4(0, 0, 0, 0);
      if (lane < 16) sc = ptx::lds128(Wsm + s * STAGE_BYTES + W_BYTES + q * (BN * 2) + lane * 16);
      __syncwarp();
      // release only once the loads from the slot have returned (ptx::mbar_arrive_after_loads)
      if (lane == 0)
        ptx::mbar_arrive_after_loads(&empty_tma[s], r[0][0] | r[GPS - 1][3] | sc.x, static_cast<uint32_t>(p.K) >> 31);
      ptx::mbar_wait(&d_empty[ss], sph ^ 1);             // previous user of the slot has been multiplied and read
      ptx::tc_fence_after();
      if (lane < 16) ptx::sts128(Ssl + ss * SC_STAGE + q * (BN * 2) + lane * 16, sc);
      const uint32_t a_addr = tmem_base + a_lane + static_cast<uint32_t>(NS * SLOT_COLS + ss * SLOT_COLS);
#pragma unroll
      for (int g = 0; g < GPS; ++g) {
        // r[g][0], r[g][1]: columns (16-block 0) g8, g8 + 8; r[g][2], r[g][3]: 16-block 1; 4 packed rows each
        tmem_st_16x256b(a_addr + g * ACOLS, r[g][0] & 0x0F0F0F0Fu, (r[g][0] >> 4) & 0x0F0F0F0Fu,
                        r[g][1] & 0x0F0F0F0Fu, (r[g][1] >> 4) & 0x0F0F0F0Fu);
        tmem_st_16x256b(a_addr + (16u << 16) + g * ACOLS, r[g][2] & 0x0F0F0F0Fu, (r[g][2] >> 4) & 0x0F0F0F0Fu,
                        r[g][3] & 0x0F0F0F0Fu, (r[g][3] >> 4) & 0x0F0F0F0Fu);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ab_full[ss]);
      if (++s == S) {
        s = 0;
        ph ^= 1;
      }
      if (++ss == NS) {
        ss = 0;
        sph ^= 1;
      }
    }
  } else {
    // =========================== epilogue warps: thread <-> column ===========================
    const int col = warp * 32 + lane;   // column of the tile == TMEM lane
    float acc = 0.f;
    int ss = 0, sph = 0;
    for (int st = 0; st < n_units; ++st) {
      ptx::mbar_wait(&mma_done[ss], sph);
      if (st == 0 && threadIdx.x == 0) stamp(p, 3);
      ptx::tc_fence_after();
      uint32_t d[16];   // d[4 g + digit]
      ptx::tmem_ld_32x32b_x16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + static_cast<uint32_t>(ss * SLOT_COLS), d);
      uint32_t sraw[GPS];
#pragma unroll
      for (int g = 0; g < GPS; ++g) sraw[g] = lds16(Ssl + ss * SC_STAGE + g * (BN * 2) + col * 2);
      const uint4 gi0 = ptx::lds128(Gi + ss * GI_STAGE), gi1 = ptx::lds128(Gi + ss * GI_STAGE + 16);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&d_empty[ss]);
      const uint32_t giw[8] = {gi0.x, gi0.y, gi0.z, gi0.w, gi1.x, gi1.y, gi1.z, gi1.w};
#pragma unroll
      for (int g = 0; g < GPS; ++g) {
        const float I = __uint_as_float(giw[2 * g]);
        const int G8 = static_cast<int>(giw[2 * g + 1]);
        const int uu = (static_cast<int>(d[4 * g]) * 128 + static_cast<int>(d[4 * g + 1])) * 128 +
                       static_cast<int>(d[4 * g + 2]) - G8;
        union {
          uint16_t u;
          T h;
        } cs;
        cs.u = static_cast<uint16_t>(sraw[g]);
        acc = fmaf(DT<T>::to_f(cs.h) * I, static_cast<float>(uu), acc);
      }
      if (++ss == NS) {
        ss = 0;
        sph ^= 1;
      }
    }
    if (threadIdx.x == 0) stamp(p, 4);
    if (Z == 1) {
      const int n = tile * BN + col;
      if (n < p.N) static_cast<T*>(p.C)[n] = epilogue<T>(acc, static_cast<const T*>(p.bias), n);
    } else {
      const uint32_t local = ptx::smem_u32(xred) + static_cast<uint32_t>(z * BN + col) * 4u;
      ptx::st_cluster_f32(ptx::mapa_rank(local, 0), acc);
    }
  }
  __syncwarp();
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
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 9) ptx::tmem_dealloc(tmem_base, tmem_cols);
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
  const int SPT = (G + GPS - 1) / GPS;
  const int tiles = (a.N + BN - 1) / BN;
  static const int stages_env = env_int("CGQ_UMMA_STAGES", 5, 2, 16);
  static const int na_env = env_int("CGQ_UMMA_SLOTS", 2, 2, 8);
  static const int z_env = env_int("CGQ_GEMV_Z", 0, 0, 8);
  static const int cps = env_int("CGQ_UMMA_CTAS_PER_SM", 4, 1, 4);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  const int slots = cps * sm_count();
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2) Z *= 2;
  if (z_env > 0) Z = z_env;
  if (Z > SPT) Z = 1;
  const int grid = tiles * Z;
  const int per_cta = (SPT + Z - 1) / Z;
  int stages = stages_env;
  if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
  const int NA = na_env;
  const int xred_bytes = Z > 1 ? XRED_BYTES : 0;
  const size_t smem = 1024 + static_cast<size_t>(NA) * (B_STAGE + GI_STAGE + SC_STAGE) +
                      static_cast<size_t>(stages) * STAGE_BYTES + xred_bytes + 8 * (2 * stages + 4 * NA) + 16;
  if (smem > 113 * 1024 + 512) {
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
             static_cast<uint64_t>(a.N) * 2, BN, GPS,
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
  prm.NA = NA;
  prm.max_units = 0;   // the digit warp reads the activations from global memory (L2-resident)
  prm.xred_bytes = xred_bytes;
  // c = S32 | a = u8 | b = s8 | A from TMEM (K along the columns) | B K-major | N = 16 | M = 128
  prm.idesc = (2u << 4) | (0u << 7) | (1u << 10) | (0u << 15) | (0u << 16) |
              (static_cast<uint32_t>(NU >> 3) << 17) | (static_cast<uint32_t>(BN >> 4) << 24);
  prm.trace = static_cast<unsigned long long*>(take_trace_buffer());

  auto kern = w4_gemv_umma_kernel<T>;
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

// M == 1 tensor-core decode path; *taken = false when the shape is left to the mma.sync kernel.
int launch_w4_gemv_umma(const GemmArgs& a, bool* taken) {
  if (a.M != 1) {
    *taken = false;
    return CGQ_OK;
  }
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, taken) : launch_t<__nv_bfloat16>(a, taken);
}

}  // namespace cgq
