// int4g32 prefill kernel (M > 8): C[M,N] = A[M,K] · ((nib(Wq) - 8) * scale), tcgen05 + TMEM.
// Replaces _dynamic_quant_matmul_s4_kernel (chatglm_q/int4/triton_ops.py:18-87) for prefill, where
// the reference re-reads and re-dequantises the whole weight for every 16-row block of A.
//
// The product is computed TRANSPOSED on the 5th-generation tensor cores:
//     D[n, m] = Σ_k  Wdq^T[n, k] · A^T[k, m]          (UMMA  M = 128 weight columns, N = MB tokens, K = 16)
// so that
//   * the dequantised weight tile — N-contiguous in memory, exactly like the packed tensor — is the
//     MN-major A operand (no transpose anywhere), and
//   * the activation tile [tokens x 64 k] (K-contiguous rows of A) is the K-major B operand, TMA-loaded
//     with the 128-byte swizzle straight into its canonical UMMA layout; any token count > 8 maps to
//     MB in {32, 64, 128, 256} without padding the 128-wide UMMA M.
//
// Warp roles (448 threads, persistent over (n-tile, token-block[, k-split]) work items, weights of one n-tile are
// shared through L2 by the CTAs working on its token blocks at the same time):
//   warp 12     TMA producer: packed weights [32 x 128 B], scales [2 x 128], activations [MB x 64]
//   warps 4-11  dequant (two per SM sub-partition: one warp per sub-partition left the pipeline at ~0.53 us per
//               64-k stage whatever the token block, the tensor pipe needs 0.13 .. 0.27 us):
//               packed smem -> registers -> fp16/bf16 tile in the canonical MN-major
//               SWIZZLE_128B layout (reference rounding: (q-8) exact, one rounding in the multiply by
//               the group scale), fence.proxy.async, mbarrier
//   warp 13     one lane issues tcgen05.mma (cta_group::1, kind::f16, fp32 accumulators in TMEM),
//               tcgen05.commit releases the smem stages; owns the TMEM allocation (2 accumulators)
//   warps 0-3   epilogue: tcgen05.ld (32 lanes x 32 bit) -> round to T -> (+ bias, second rounding)
//               -> C, overlapped with the next tile's MMAs through the second accumulator
//
// The int8 variant (chatglm_q/int8/triton_ops.py:13-84) shares the skeleton: its weight is [N, K]
// K-contiguous, so the dequantised tile (round_T(q * scale[n]), int8/qlinear.py:38) is a K-major A
// operand; the packed stage is [128 n x 64 k] bytes and the per-channel scales come from global memory.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace cgq {
namespace {

constexpr int TN = 128;          // weight columns per tile  (UMMA M)
constexpr int BK = 64;           // k per pipeline stage      (4 UMMA K-steps)
constexpr int MAX_STAGES = 12;   // TMA ring depth: as many stages as fit in shared memory, at most this
constexpr int ABUFS = 4;         // dequantised-A ring: the dequant warps run up to 4 stages ahead of the MMAs (with 2, the
                                 // store -> fence -> MMA -> commit -> mbarrier -> next store chain paced every stage at ~0.5 us)
constexpr int kDeqWarps = 8;         // dequant warps (two per SM sub-partition)
constexpr int kTmaWarp = 4 + kDeqWarps, kMmaWarp = kTmaWarp + 1;
constexpr int kThreads = (kMmaWarp + 1) * 32;
constexpr int P_BYTES = (BK / 2) * TN;       // 4096: packed int4 tile
constexpr int S_BYTES = (BK / 32) * TN * 2;  // 512: scale tile
constexpr int PS_BYTES = 5120;               // packed + scales, padded to keep B tiles 1024-aligned
constexpr int P8_BYTES = TN * BK;            // 8192: int8 tile [128 n x 64 k]
constexpr int A_BYTES = TN * BK * 2;         // 16384: dequantised tile
constexpr int ATOM = 1024;                   // 8 rows x 128 B swizzle atom

struct Params {
  const void* bias;
  void* C;
  int64_t ldc;
  int M, N, K;
  int MB;        // tokens per tile (UMMA N)
  int n_tiles, m_blocks, k_stages;
  uint32_t idesc;
  const void* scale8;   // int8 variant: per-channel scales [N]
  // split-K (small M: fewer (n-tile, token-block) tiles than SMs): a work item is (tile, k-split); the splits of a
  // tile run on different CTAs at the same time, write fp32 partial tiles to the workspace and the LAST one to arrive
  // (per-tile counter) adds them in split order -- deterministic -- and runs the epilogue
  int stages;           // TMA ring depth of this launch
  int k_splits;
  float* ws_part;       // [items][128][MB] fp32
  int* ws_ctr;          // one counter per tile (128-byte stride), self-cleaning
};

// 64-bit shared-memory matrix descriptor (SWIZZLE_128B, Blackwell version 1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) |
         (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
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
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// One 32-bit word = 4 packed bytes = columns (c0, c1, c2, c3) of one packed row (k = 2r, 2r+1).
// Produces round_T((q - 8) * s) for the column PAIRS (c0, c2) and (c1, c3) of both k values;
// s02 / s13 hold the matching scale pairs.  Bit-exact with chatglm_q/int4/qlinear.py:29-32.
template <typename T>
struct Deq;
template <>
struct Deq<__half> {
  __device__ static __forceinline__ void run(uint32_t w, uint32_t s02, uint32_t s13, uint32_t& e02,
                                             uint32_t& e13, uint32_t& o02, uint32_t& o13) {
    const uint32_t w8 = w >> 8;
    e02 = h2_mul(h2_sub(ptx::and_or(w, 0x000F000Fu, 0x64006400u), 0x64086408u), s02);   // (1024+q)-1032
    e13 = h2_mul(h2_sub(ptx::and_or(w8, 0x000F000Fu, 0x64006400u), 0x64086408u), s13);
    o02 = h2_mul(h2_fma(ptx::and_or(w, 0x00F000F0u, 0x64006400u), 0x2C002C00u, 0xD480D480u), s02);  // (1024+16q)/16-72
    o13 = h2_mul(h2_fma(ptx::and_or(w8, 0x00F000F0u, 0x64006400u), 0x2C002C00u, 0xD480D480u), s13);
  }
};
template <>
struct Deq<__nv_bfloat16> {
  __device__ static __forceinline__ void run(uint32_t w, uint32_t s02, uint32_t s13, uint32_t& e02,
                                             uint32_t& e13, uint32_t& o02, uint32_t& o13) {
    // bf16 keeps 8 significant bits: 128 + q is exact, the high nibble is shifted down first
    e02 = bf2_mul(bf2_sub(ptx::and_or(w, 0x000F000Fu, 0x43004300u), 0x43084308u), s02);  // (128+q)-136
    e13 = bf2_mul(bf2_sub(ptx::and_or(w >> 8, 0x000F000Fu, 0x43004300u), 0x43084308u), s13);
    o02 = bf2_mul(bf2_sub(ptx::and_or(w >> 4, 0x000F000Fu, 0x43004300u), 0x43084308u), s02);
    o13 = bf2_mul(bf2_sub(ptx::and_or(w >> 12, 0x000F000Fu, 0x43004300u), 0x43084308u), s13);
  }
};

// word = 4 consecutive int8 (k..k+3) of one weight row -> round_T(q * s) pairs (k,k+1), (k+2,k+3);
// s2 = (s, s).  int8 -> T is exact, the multiply rounds once (int8/qlinear.py:38).
template <typename T>
struct Deq8;
template <>
struct Deq8<__half> {
  __device__ static __forceinline__ void run(uint32_t w, uint32_t s2, uint32_t& p01, uint32_t& p23) {
    const uint32_t x = w ^ 0x80808080u;  // q + 128 as unsigned bytes
    p01 = h2_mul(h2_sub(__byte_perm(x, 0x64646464u, 0x4140), 0x64806480u), s2);  // (1024+128+q) - 1152
    p23 = h2_mul(h2_sub(__byte_perm(x, 0x64646464u, 0x4342), 0x64806480u), s2);
  }
};
template <>
struct Deq8<__nv_bfloat16> {
  __device__ static __forceinline__ uint32_t one(uint32_t x, int sel) {
    return __float_as_uint(__uint_as_float(__byte_perm(x, 0x4B000000u, sel)) - 8388736.f);  // exact q
  }
  __device__ static __forceinline__ void run(uint32_t w, uint32_t s2, uint32_t& p01, uint32_t& p23) {
    const uint32_t x = w ^ 0x80808080u;
    p01 = bf2_mul(__byte_perm(one(x, 0x7440), one(x, 0x7441), 0x7632), s2);  // high halves = bf16(q)
    p23 = bf2_mul(__byte_perm(one(x, 0x7442), one(x, 0x7443), 0x7632), s2);
  }
};

template <typename T, bool kW8>
__global__ void __launch_bounds__(kThreads, 1)
    wq_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmS,
                      const __grid_constant__ CUtensorMap tmA, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int MB = p.MB;
  const uint32_t b_bytes = static_cast<uint32_t>(MB) * 128u;   // [MB tokens x 64 k] 16-bit
  const uint32_t stage_bytes = b_bytes + (kW8 ? P8_BYTES : PS_BYTES);
  // layout: A buffers | stages { B tile | packed | scales } | barriers | tmem ptr
  const uint32_t off_stage = ABUFS * A_BYTES;
  const int STAGES = p.stages;
  const uint32_t off_bar = off_stage + STAGES * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gen + off_bar);
  uint64_t* full_tma = bars;                    // [STAGES]
  uint64_t* empty_tma = full_tma + STAGES;      // [STAGES]
  uint64_t* a_full = empty_tma + STAGES;        // [ABUFS]
  uint64_t* a_empty = a_full + ABUFS;           // [ABUFS]
  uint64_t* acc_full = a_empty + ABUFS;         // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.n_tiles * p.m_blocks * p.k_splits;     // work items
  const uint32_t tmem_cols = (2 * MB <= 32) ? 32u : (2 * MB <= 64) ? 64u : (2 * MB <= 128) ? 128u
                             : (2 * MB <= 256) ? 256u : 512u;

  if (threadIdx.x == kTmaWarp * 32) {
    ptx::prefetch_tmap(&tmP);
    if (!kW8) ptx::prefetch_tmap(&tmS);
    ptx::prefetch_tmap(&tmA);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_tma[s], 1);
      ptx::mbar_init(&empty_tma[s], 1);
    }
    for (int b = 0; b < ABUFS; ++b) {
      ptx::mbar_init(&a_full[b], 4);      // the four warps of the group that owns the stage
      ptx::mbar_init(&a_empty[b], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&acc_full[a], 1);
      ptx::mbar_init(&acc_empty[a], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    // ===================================== TMA producer =====================================
    {   // every lane runs the loop, lane 0 issues (see the MMA issuer below)
      const uint32_t issue = lane == 0 ? 1u : 0u;
      const uint64_t pol_w = ptx::policy_evict_first();   // weights: streamed (re-use is in L2 window)
      const uint64_t pol_a = ptx::policy_evict_last();    // activations: re-read by every n-tile
      int s = 0, ph = 0;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / p.k_splits, sp = item - tile * p.k_splits;
        const int nt = tile / p.m_blocks, mb = tile - nt * p.m_blocks;
        const int ks0 = p.k_stages * sp / p.k_splits, ks1 = p.k_stages * (sp + 1) / p.k_splits;
        for (int ks = ks0; ks < ks1; ++ks) {
          ptx::mbar_wait(&empty_tma[s], ph ^ 1);
          const uint32_t st = base + off_stage + s * stage_bytes;
          ptx::mbar_expect_tx_warp(&full_tma[s], b_bytes + (kW8 ? P8_BYTES : P_BYTES + S_BYTES), issue);
          ptx::tma_load_2d_warp(st, &tmA, ks * BK, mb * MB, &full_tma[s], pol_a, issue);
          if (kW8) {
            ptx::tma_load_2d_warp(st + b_bytes, &tmP, ks * BK, nt * TN, &full_tma[s], pol_w, issue);
          } else {
            ptx::tma_load_2d_warp(st + b_bytes, &tmP, nt * TN, ks * (BK / 2), &full_tma[s], pol_w, issue);
            ptx::tma_load_2d_warp(st + b_bytes + P_BYTES, &tmS, nt * TN, ks * (BK / 32), &full_tma[s], pol_w, issue);
          }
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    // EVERY lane of the warp runs this loop (converged), lane 0 issues (predicate inside the asm): inside a divergent
    // `if (lane == 0)` region the compiler wraps each tcgen05 instruction in an ELECT / R2UR / BRA.U.ANY waterfall
    // loop -- ~75 dependent instructions per stage for the single issuing thread (tools/umma_rate.cu: 1005 -> 509
    // cycles per stage of four m128 k16 MMAs + commit, i.e. the 4 x 127 cycles the MMAs take whatever their N).
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;
      int s = 0, ph = 0, ab = 0, aph = 0, acc = 0, cph = 0;
      // The issuer is ONE thread: everything it executes per stage is serial latency in front of the tensor pipe (the
      // per-stage descriptor construction -- eight 64-bit shift / mask chains -- was what paced the pipeline at ~780
      // cycles per 64-k stage whatever the token block).  Descriptors are built once; a stage adds its offset to the
      // 14-bit start-address field (shared memory is < 256 KB: the field cannot carry).
      // int4 A: MN-major, 64-column halves 8 KB apart (LBO), 8-k atoms 1 KB apart (SBO); K-step = 2 atoms
      // int8 A: K-major rows of 128 B like B; B: K-major rows of 128 B, 8-row atoms 1 KB apart; K-step = 32 B
      const uint64_t adesc0 = kW8 ? make_desc(base, 16, ATOM) : make_desc(base, (BK / 8) * ATOM, ATOM);
      const uint64_t bdesc0 = make_desc(base + off_stage, 16, ATOM);
      constexpr uint32_t a_kstep = kW8 ? (32u >> 4) : ((2u * ATOM) >> 4);
      const uint32_t b_stage = stage_bytes >> 4;
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int sp = item % p.k_splits;
        const int ks0 = p.k_stages * sp / p.k_splits, ks1 = p.k_stages * (sp + 1) / p.k_splits;
        ptx::mbar_wait(&acc_empty[acc], cph ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * MB);
        for (int ks = ks0; ks < ks1; ++ks) {
          ptx::mbar_wait(&full_tma[s], ph);
          ptx::mbar_wait(&a_full[ab], aph);
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + static_cast<uint32_t>(ab) * (static_cast<uint32_t>(A_BYTES) >> 4);
          const uint64_t bdesc = bdesc0 + static_cast<uint32_t>(s) * b_stage;
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4)
            ptx::umma_f16_ss_warp(d_tmem, adesc + k4 * a_kstep, bdesc + k4 * 2u, p.idesc, (ks != ks0 || k4 != 0) ? 1u : 0u,
                                  issue);
          {
          }
          ptx::umma_commit_warp(&empty_tma[s], issue);   // stage (activations + packed) reusable when MMAs retire
          ptx::umma_commit_warp(&a_empty[ab], issue);
          if (++s == STAGES) {
            s = 0;
            ph ^= 1;
          }
          if (++ab == ABUFS) {
            ab = 0;
            aph ^= 1;
          }
        }
        ptx::umma_commit_warp(&acc_full[acc], issue);
        if (++acc == 2) {
          acc = 0;
          cph ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================== dequant warps =====================================
    // thread -> (16-byte output chunk c of 8 columns, packed rows r0 + 8 i); 128 threads cover
    // 32 packed rows x 16 chunks.  Chunk element order is (c0,c2,c1,c3,c4,c6,c5,c7): the epilogue
    // un-permutes the TMEM lanes.
    // Two groups of four warps take ALTERNATE stages: one group's store -> fence -> arrive chain of a stage overlaps
    // the other group's loads and conversions of the next one (a single group paced the pipeline at ~0.5 us per
    // stage whatever the token block).  Inside a group: thread -> (16-byte output chunk c of 8 columns, packed rows
    // r0 + 8 i); 128 threads cover 32 packed rows x 16 chunks.  Chunk element order is (c0,c2,c1,c3,c4,c6,c5,c7):
    // the epilogue un-permutes the TMEM lanes.
    const int grp = (warp - 4) >> 2;
    const int t = (threadIdx.x - 128) & 127;
    int s = 0, ph = 0, ab = 0, aph = 0;
    unsigned stage_no = 0;
    auto advance = [&]() {
      ++stage_no;
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
      if (++ab == ABUFS) {
        ab = 0;
        aph ^= 1;
      }
    };
    if constexpr (kW8) {
      // thread -> (16-byte chunk j = 8 consecutive k, weight rows r0 + 16 i): 128 rows x 8 chunks
      const int j = t & 7, r0 = t >> 3;
      const T* sc = static_cast<const T*>(p.scale8);
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int tile = item / p.k_splits, sp = item - tile * p.k_splits;
        const int nt = tile / p.m_blocks;
        const int n_ks = p.k_stages * (sp + 1) / p.k_splits - p.k_stages * sp / p.k_splits;
        uint32_t s2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = nt * TN + r0 + 16 * i;
          union {
            uint32_t u;
            T h[2];
          } cv;
          cv.h[0] = cv.h[1] = (n < p.N) ? sc[n] : DT<T>::from_f(0.f);
          s2[i] = cv.u;
        }
        for (int ks = 0; ks < n_ks; ++ks, advance()) {
          if (static_cast<int>(stage_no & 1u) != grp) continue;
          ptx::mbar_wait(&full_tma[s], ph);
          const uint32_t st = base + off_stage + s * stage_bytes + b_bytes;
          uint32_t out[8][4];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint2 pk = ptx::lds64(st + (r0 + 16 * i) * BK + j * 8);
            Deq8<T>::run(pk.x, s2[i], out[i][0], out[i][1]);
            Deq8<T>::run(pk.y, s2[i], out[i][2], out[i][3]);
          }
          ptx::mbar_wait(&a_empty[ab], aph ^ 1);
          const uint32_t a_addr = base + ab * A_BYTES;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = r0 + 16 * i;
            ptx::sts128(a_addr + (r >> 3) * ATOM + (r & 7) * 128 + ((j ^ (r & 7)) << 4),
                        make_uint4(out[i][0], out[i][1], out[i][2], out[i][3]));
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[ab]);
        }
      }
    } else {
      const int c = t & 15;          // column chunk: columns 8c .. 8c+7
      const int r0 = t >> 4;         // 0..7
      for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
        const int sp = item % p.k_splits;
        const int n_ks = p.k_stages * (sp + 1) / p.k_splits - p.k_stages * sp / p.k_splits;
        for (int ks = 0; ks < n_ks; ++ks, advance()) {
          if (static_cast<int>(stage_no & 1u) != grp) continue;
          ptx::mbar_wait(&full_tma[s], ph);
          const uint32_t st = base + off_stage + s * stage_bytes + b_bytes;
          uint32_t out[4][8];  // [row i][k parity * 4 + word]
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint4 sv = ptx::lds128(st + P_BYTES + g * (TN * 2) + c * 16);
            const uint32_t s02a = __byte_perm(sv.x, sv.y, 0x5410), s13a = __byte_perm(sv.x, sv.y, 0x7632);
            const uint32_t s02b = __byte_perm(sv.z, sv.w, 0x5410), s13b = __byte_perm(sv.z, sv.w, 0x7632);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int i = 2 * g + h;                 // packed row r0 + 8 i  (rows 0..15 = group 0)
              const uint2 pk = ptx::lds64(st + (r0 + 8 * i) * TN + c * 8);
              Deq<T>::run(pk.x, s02a, s13a, out[i][0], out[i][1], out[i][4], out[i][5]);
              Deq<T>::run(pk.y, s02b, s13b, out[i][2], out[i][3], out[i][6], out[i][7]);
            }
          }
          ptx::mbar_wait(&a_empty[ab], aph ^ 1);
          const uint32_t a_addr = base + ab * A_BYTES + (c >> 3) * ((BK / 8) * ATOM);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              const int k = 2 * (r0 + 8 * i) + par;    // k row inside the stage
              const uint32_t addr = a_addr + (k >> 3) * ATOM + (k & 7) * 128 + (((c & 7) ^ (k & 7)) << 4);
              ptx::sts128(addr, make_uint4(out[i][4 * par], out[i][4 * par + 1], out[i][4 * par + 2],
                                           out[i][4 * par + 3]));
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[ab]);
        }
      }
    }
  } else {
    // ===================================== epilogue warps =====================================
    T* Cp = static_cast<T*>(p.C);
    const T* bias = static_cast<const T*>(p.bias);
    const int row = warp * 32 + lane;                       // TMEM lane = permuted column of the tile
    // int4: (0,2,1,3) un-permute of the dequant chunk order; int8: identity
    const int col_in_tile = kW8 ? row : ((row & ~3) | ((row & 1) << 1) | ((row >> 1) & 1));
    int acc = 0, cph = 0;
    int* last_flag = reinterpret_cast<int*>(tmem_slot + 1);      // epilogue warps' broadcast word (shared memory)
    for (int item = blockIdx.x; item < total_tiles; item += gridDim.x) {
      const int tile = item / p.k_splits;
      const int nt = tile / p.m_blocks, mb = tile - nt * p.m_blocks;
      const int n = nt * TN + col_in_tile;
      const int m0 = mb * MB;
      ptx::mbar_wait(&acc_full[acc], cph);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) +
                             static_cast<uint32_t>(acc * MB);
      // partial tile layout [item][16-token chunk][128 rows][16 floats]: a warp writes 2 KB contiguous per chunk
      float* part = p.k_splits > 1 ? p.ws_part + static_cast<size_t>(item) * TN * MB + row * 16 : nullptr;
      for (int c0 = 0; c0 < MB; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (part != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(part + c0 * TN + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int m = m0 + c0 + j;
            if (m < p.M) Cp[static_cast<int64_t>(m) * p.ldc + n] = epilogue<T>(__uint_as_float(v[j]), bias, n);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        cph ^= 1;
      }
      if (p.k_splits > 1) {
        // the last split of this tile to arrive adds all partial tiles in split order and stores the result
        __threadfence();
        ptx::named_bar_sync(1, 128);
        if (threadIdx.x == 0) {
          int* ctr = p.ws_ctr + static_cast<size_t>(tile) * kCounterStride;
          const int old = atomicAdd(ctr, 1);
          if (old == p.k_splits - 1) *ctr = 0;      // self-cleaning for the next launch
          *last_flag = old;
        }
        ptx::named_bar_sync(1, 128);
        const bool last = *last_flag == p.k_splits - 1;
        ptx::named_bar_sync(1, 128);                // (the flag word is rewritten by the next item)
        if (last) {
          __threadfence();
          const float* p0 = p.ws_part + static_cast<size_t>(tile) * p.k_splits * TN * MB + row * 16;
          for (int c0 = 0; c0 < MB; c0 += 16) {
            // all splits' 16 values of this row in flight at once (16 independent 16-byte loads at 4 splits)
            float4 t[4][4];
#pragma unroll
            for (int sp = 0; sp < 4; ++sp) {
              if (sp < p.k_splits) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                               : "=f"(t[sp][q].x), "=f"(t[sp][q].y), "=f"(t[sp][q].z), "=f"(t[sp][q].w)
                               : "l"(p0 + static_cast<size_t>(sp) * TN * MB + c0 * TN + 4 * q));
              }
            }
            float vv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) vv[j] = 0.f;
#pragma unroll
            for (int sp = 0; sp < 4; ++sp) {
              if (sp < p.k_splits) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  vv[4 * q] += t[sp][q].x;
                  vv[4 * q + 1] += t[sp][q].y;
                  vv[4 * q + 2] += t[sp][q].z;
                  vv[4 * q + 3] += t[sp][q].w;
                }
              }
            }
            if (n < p.N) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int m = m0 + c0 + j;
                if (m < p.M) Cp[static_cast<int64_t>(m) * p.ldc + n] = epilogue<T>(vv[j], bias, n);
              }
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

template <typename T, bool kW8>
int launch_t(const GemmArgs& a) {
  const int MB = a.M > 128 ? 256 : a.M > 64 ? 128 : a.M > 32 ? 64 : 32;
  Params prm;
  prm.bias = a.bias;
  prm.C = a.C;
  prm.ldc = a.ldc;
  prm.M = a.M;
  prm.N = a.N;
  prm.K = a.K;
  prm.MB = MB;
  prm.n_tiles = (a.N + TN - 1) / TN;
  prm.m_blocks = (a.M + MB - 1) / MB;
  prm.k_stages = (a.K + BK - 1) / BK;
  prm.scale8 = a.scale;
  // split-K when the tiles alone leave SMs idle: the smallest split count (<= 8, >= 4 k-stages per split) whose
  // last wave is >= 85 % full (at most 4: the fix-up keeps all splits' values in registers), bounded by the workspace
  prm.k_splits = 1;
  prm.ws_part = nullptr;
  prm.ws_ctr = nullptr;
  {
    static const int forced = [] {
      const char* e = getenv("CGQ_TC_SPLITS");
      return e != nullptr ? atoi(e) : 0;
    }();
    const int tiles = prm.n_tiles * prm.m_blocks, sms = sm_count();
    const size_t tile_bytes = static_cast<size_t>(TN) * MB * sizeof(float);
    auto eff = [&](int ks) {
      const int items = tiles * ks, waves = (items + sms - 1) / sms;
      return static_cast<double>(items) / (static_cast<double>(waves) * sms);
    };
    int best = 1;
    if (a.workspace != nullptr && tiles <= kMaxTiles && eff(1) < 0.5) {      // (measured: with a wave >= half full the
                                                                             // fix-up costs more than the split buys)
      for (int ks = 2; ks <= 4; ++ks) {
        if (prm.k_stages / ks < 4 || static_cast<size_t>(tiles) * ks * tile_bytes > kPartialBytes) break;
        if (eff(ks) > eff(best) + 0.02) best = ks;
        if (eff(best) >= 0.9) break;
      }
    }
    if (forced >= 1 && forced <= 4 && a.workspace != nullptr &&
        static_cast<size_t>(tiles) * forced * tile_bytes <= kPartialBytes && forced <= prm.k_stages)
      best = forced;
    if (best > 1) {
      prm.k_splits = best;
      prm.ws_ctr = static_cast<int*>(a.workspace);
      prm.ws_part = reinterpret_cast<float*>(static_cast<char*>(a.workspace) + kPartialOffset);
    }
  }
  const uint32_t fmt = (DT<T>::code == CGQ_DTYPE_F16) ? 0u : 1u;
  // c=F32 | a,b format | A major (int4: MN, int8: K) | B K-major | N = MB | M = 128
  prm.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((kW8 ? 0u : 1u) << 15) | (0u << 16) |
              (static_cast<uint32_t>(MB >> 3) << 17) | (static_cast<uint32_t>(TN >> 4) << 24);

  const CUtensorMapDataType dt16 = (DT<T>::code == CGQ_DTYPE_F16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUtensorMap tmP, tmS, tmA;
  int rc;
  if (kW8) {
    TmapKey kp{a.Wq, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N),
               static_cast<uint64_t>(a.K), BK, TN, CU_TENSOR_MAP_DATA_TYPE_UINT8,
               CU_TENSOR_MAP_SWIZZLE_NONE};
    rc = get_tmap_2d(kp, &tmP);
    if (rc != CGQ_OK) return rc;
    tmS = tmP;  // unused
  } else {
    TmapKey kp{a.Wq, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K / 2),
               static_cast<uint64_t>(a.N), TN, BK / 2, CU_TENSOR_MAP_DATA_TYPE_UINT8,
               CU_TENSOR_MAP_SWIZZLE_NONE};
    rc = get_tmap_2d(kp, &tmP);
    if (rc != CGQ_OK) return rc;
    TmapKey ks{a.scale, static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K / 32),
               static_cast<uint64_t>(a.N) * 2, TN, BK / 32, dt16, CU_TENSOR_MAP_SWIZZLE_NONE};
    rc = get_tmap_2d(ks, &tmS);
    if (rc != CGQ_OK) return rc;
  }
  TmapKey ka{a.A, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.M),
             static_cast<uint64_t>(a.lda) * 2, BK, static_cast<uint32_t>(MB), dt16,
             CU_TENSOR_MAP_SWIZZLE_128B};
  rc = get_tmap_2d(ka, &tmA);
  if (rc != CGQ_OK) return rc;

  // the per-CTA pipeline is bound by (stages in flight) / (TMA round-trip latency): take every stage that fits
  const size_t stage_b = static_cast<size_t>(MB) * 128 + (kW8 ? P8_BYTES : PS_BYTES);
  const size_t fixed_b = 1024 + ABUFS * A_BYTES + 8 * (2 * MAX_STAGES + 2 * ABUFS + 4) + 32;
  int stages = static_cast<int>((232448 - fixed_b) / stage_b);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  static const int st_env = [] {
    const char* e = getenv("CGQ_TC_STAGES");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (st_env >= 2 && st_env < stages) stages = st_env;
  prm.stages = stages;
  const size_t smem = fixed_b + static_cast<size_t>(stages) * stage_b;
  auto kern = wq_gemm_tc_kernel<T, kW8>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && smem > configured[dev]) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    configured[dev] = smem;
  }
  int grid = prm.n_tiles * prm.m_blocks * prm.k_splits;
  if (grid > sm_count()) grid = sm_count();
  kern<<<grid, kThreads, smem, a.stream>>>(tmP, tmS, tmA, prm);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

}  // namespace

bool w4_tc_supported(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.M >= 1 && a.K % 32 == 0 && a.N % 16 == 0 && al16(a.Wq) && al16(a.scale) && al16(a.A) &&
         a.lda % 8 == 0;
}
bool w8_tc_supported(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.M >= 1 && a.K % 16 == 0 && al16(a.Wq) && al16(a.A) && a.lda % 8 == 0 &&
         (reinterpret_cast<uintptr_t>(a.scale) & 1) == 0;
}

int launch_w4_tc(const GemmArgs& a) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half, false>(a) : launch_t<__nv_bfloat16, false>(a);
}
int launch_w8_tc(const GemmArgs& a) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half, true>(a) : launch_t<__nv_bfloat16, true>(a);
}

}  // namespace cgq
