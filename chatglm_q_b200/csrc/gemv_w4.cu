// int4g32 decode kernel (M <= 8 token rows): C[M,N] = A[M,K] · ((nib(Wq) - 8) * scale).
// Replaces _dynamic_quant_matmul_s4_kernel (chatglm_q/int4/triton_ops.py:18-87) for decode.
//
// HBM-bound design (B200: 148 SMs, ~6.5 TB/s measured):
//   * work split = (128-column tile) x (Z contiguous k-bands); one CTA per (tile, band), the Z CTAs
//     of a tile form a thread-block CLUSTER.  All CTAs start at the top of their band and walk down
//     in lockstep, so at any moment the chip reads a few row bands of the [K/2, N] byte matrix that
//     are contiguous across ALL column tiles: the access pattern DRAM likes (measured: scattered
//     stream-K strips cap at ~4.5 TB/s with no compute at all, lockstep bands reach 5.8 TB/s);
//   * the Z band sums of a tile are reduced through DISTRIBUTED SHARED MEMORY (st.shared::cluster
//     into rank 0 + barrier.cluster), in rank order: deterministic, no workspace, no global
//     round trips on the critical path of a 1.5 us layer;
//   * small CTAs (4 consumer warps + 1 producer warp, ~45 KB of shared memory at M=1), at most two
//     per SM per launch, so that the CTAs of the NEXT launch fit beside them: with programmatic
//     dependent launch the next kernel's producers fill their rings with weights (which do not
//     depend on the previous kernel) while this kernel is still computing;
//   * one producer lane per CTA feeds an S-deep shared-memory ring with TMA: a [64 x 128] byte
//     tile of packed weights (128-byte swizzle), the [4 x 128] scale tile, and the matching
//     128-k slice of each activation row (cp.async.bulk), all landing on one mbarrier;
//   * 4 consumer warps each own one 32-k quantisation group of the stage.  A lane reads two
//     16-byte runs (16 columns, packed rows 2t and 2t+1), and one PRMT per column makes the
//     32-bit word [byte(r), -, byte(r+1), -], whose masked halves are exactly the (k, k') pairs
//     of an m16n8k16 A fragment with the weight COLUMN as the MMA row.  Activations are the
//     B fragment (<= 8 tokens), accumulation is fp32 in the tensor core;
//   * the group scale is applied to the group's fp32 partial sum:  out = Σ_g s_g · Σ_{k∈g} a_k (q_k-8)
//     (more accurate than the reference's per-element fp16 rounding; within the 1e-2 parity bar);
//   * fp16 fast variant ("trick"): the masked nibble IS an fp16 subnormal q·2^-24 (q·2^-20 for
//     the high nibble, compensated by scaling the odd-k activations by 2^-4), so no int->fp
//     conversion is executed at all; the -8 offset becomes -8·Σ_{k∈g} a_k, obtained from one
//     extra MMA against a constant fragment.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"

namespace cgq {
namespace {

constexpr int BN = 128;            // columns per tile (TMA inner box, bytes)
constexpr int CW = 4;              // consumer warps == quantisation groups per stage
constexpr int ROWS = 16 * CW;      // packed byte rows per stage
constexpr int KSTAGE = 32 * CW;    // k values per stage
constexpr int MMAX = 8;            // token rows (MMA n)
constexpr int A_STRIDE = KSTAGE * 2 + 32;  // bytes per token row of the activation slice (+32: banks)
constexpr int W_BYTES = ROWS * BN;
constexpr int S_BYTES = CW * BN * 2;
constexpr int kThreads = (CW + 1) * 32;

template <bool kM1>
struct Cfg {
  static constexpr int MR = kM1 ? 1 : MMAX;           // token rows staged / reduced
  static constexpr int A_BYTES = MR * A_STRIDE;
  static constexpr int RED_BYTES = CW * MR * BN * 4;
  static constexpr int XRED_BYTES = 8 * MR * BN * 4;   // band sums of up to 8 cluster ranks (rank 0)
  static constexpr int STAGE_BYTES = W_BYTES + S_BYTES + A_BYTES;
};

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

struct Params {
  const void* A;
  int64_t lda;
  const void* bias;
  void* C;
  int64_t ldc;
  int M, N, K;
  int SPT;     // k-stages per column tile
  int Z;       // k-bands per tile == cluster size
  int S;       // ring depth
  unsigned long long* trace;  // optional timeline (cgq_debug_trace), 8 words per CTA
};

__device__ __forceinline__ void stamp(const Params& p, int slot) {
  if (p.trace != nullptr && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + slot] = t;
  }
}

template <typename T, bool kTrick, bool kM1>
__global__ void __launch_bounds__(kThreads, kM1 ? 4 : 2)
    w4_gemv_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmS,
                   const Params p) {
  using C = Cfg<kM1>;
  constexpr int MR = C::MR;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));  // generic pointer to aligned base
  const int S = p.S;
  const uint32_t Wsm = base;
  const uint32_t Ssm = Wsm + S * W_BYTES;
  const uint32_t Asm = Ssm + S * S_BYTES;
  const uint32_t off_red = S * C::STAGE_BYTES;
  float* red = reinterpret_cast<float*>(gen + off_red);
  float* xred = reinterpret_cast<float*>(gen + off_red + C::RED_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + C::RED_BYTES + C::XRED_BYTES);
  uint64_t* empty = full + S;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Z = p.Z;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;   // z == rank in the cluster
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;      // this CTA's k-stages of the tile
  const int n_units = u1 - u0;
  const T* A = static_cast<const T*>(p.A);

  if (threadIdx.x == 0) stamp(p, 0);
  // ---- prologue: barriers, zeroed activation ring (k >= K of a ragged last stage must read 0)
  if (threadIdx.x == CW * 32) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmS);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < S * C::A_BYTES / 16; i += kThreads)
    ptx::sts128(Asm + i * 16, make_uint4(0, 0, 0, 0));
  ptx::fence_proxy_async_smem();
  __syncthreads();
  // Let the next kernel in the stream start its own prologue / weight prefetch (PDL).
  ptx::pdl_launch_dependents();

  if (warp == CW) {
    // =========================== producer: one lane drives TMA ===========================
    if (lane == 0) {
      stamp(p, 1);
      const uint64_t pol = ptx::policy_evict_first();
      auto issue_w = [&](int i, int slot) {
        const int ks = u0 + i;
        const int kvalid = min(KSTAGE, p.K - ks * KSTAGE);
        ptx::mbar_expect_tx(&full[slot], W_BYTES + S_BYTES + p.M * kvalid * 2);
        ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, tile * BN, ks * ROWS, &full[slot], pol);
        ptx::tma_load_2d(gen + S * W_BYTES + slot * S_BYTES, &tmS, tile * BN, ks * CW, &full[slot],
                         pol);
      };
      auto issue_a = [&](int i, int slot) {
        const int ks = u0 + i;
        const int kvalid = min(KSTAGE, p.K - ks * KSTAGE);
        uint8_t* dst = gen + S * (W_BYTES + S_BYTES) + slot * C::A_BYTES;
        if (kM1) {
          ptx::bulk_load_1d(dst, A + ks * KSTAGE, kvalid * 2, &full[slot]);
        } else {
          for (int m = 0; m < p.M; ++m)
            ptx::bulk_load_1d(dst + m * A_STRIDE, A + m * p.lda + ks * KSTAGE, kvalid * 2,
                              &full[slot]);
        }
      };
      const int prefill = min(n_units, S);
      // weights do not depend on the previous kernel: start streaming them before the PDL wait
      for (int i = 0; i < prefill; ++i) issue_w(i, i);
      ptx::pdl_wait_prior_grid();
      for (int i = 0; i < prefill; ++i) issue_a(i, i);
      int slot = 0, phase = 1;
      for (int i = prefill; i < n_units; ++i) {
        ptx::mbar_wait(&empty[slot], phase ^ 1);
        issue_w(i, slot);
        issue_a(i, slot);
        if (++slot == S) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
  // =========================== consumers ===========================
  ptx::pdl_wait_prior_grid();
  if (threadIdx.x == 0) stamp(p, 2);
  const int g = lane >> 2, tig = lane & 3;
  constexpr int NT = kM1 ? 2 : 4;  // accumulator registers kept per MMA tile
  float tot[8][NT];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < NT; ++i) tot[j][i] = 0.f;

  const bool has_tok = kM1 ? (g == 0) : (g < p.M);
  int slot = 0, phase = 0;
  for (int it = 0; it < n_units; ++it) {
    ptx::mbar_wait(&full[slot], phase);
    if (it == 0 && threadIdx.x == 0) stamp(p, 3);
    const uint32_t wrow = Wsm + slot * W_BYTES + (16 * warp) * BN;
    const uint32_t srow = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 32;
    const uint32_t arow = Asm + slot * C::A_BYTES + (kM1 ? 0 : g * A_STRIDE) + (32 * warp + 4 * tig) * 2;

    float grp[8][4];
    float ag[4];
    const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#ifdef CGQ_HACK_NO_COMPUTE
#pragma unroll
    for (int j = 0; j < 8; ++j) grp[j][0] = grp[j][1] = grp[j][2] = grp[j][3] = 1.f;
    ag[0] = ag[1] = 0.f;
    for (int b = 0; b < 0; ++b) {
#else
#pragma unroll
    for (int b = 0; b < 2; ++b) {
#endif
      const int r = 8 * b + 2 * tig;
      const uint4 q = ptx::lds128(wrow + r * BN + ((g ^ (2 * tig)) << 4));
      const uint4 pp = ptx::lds128(wrow + (r + 1) * BN + ((g ^ (2 * tig + 1)) << 4));
      uint32_t b0 = 0, b1 = 0;
      if (has_tok) {
        const uint2 av = ptx::lds64(arow + 32 * b);  // a[k0+4t .. k0+4t+3]
        b0 = __byte_perm(av.x, av.y, 0x5410);        // (a[+0], a[+2]) <-> low nibbles of rows r, r+1
        b1 = __byte_perm(av.x, av.y, 0x7632);        // (a[+1], a[+3]) <-> high nibbles
        if (kTrick) b1 = h2_mul(b1, 0x2C002C00u);    // * 2^-4: high nibble enters as q * 2^-20
      }
      if (kTrick) {
        const uint32_t ones[4] = {0x3C003C00u, 0x3C003C00u, 0x4C004C00u, 0x4C004C00u};  // 1,1 | 16,16
        if (b == 0)
          ptx::mma_16816(ag, ones, b0, b1, zero4, T());
        else
          ptx::mma_16816(ag, ones, b0, b1, ag, T());
      }
      const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
      const uint32_t pw[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t x = qw[j >> 1], y = pw[j >> 1];
        const uint32_t v0 = (j & 1) ? __byte_perm(x, y, 0x6622) : __byte_perm(x, y, 0x4400);
        const uint32_t v1 = (j & 1) ? __byte_perm(x, y, 0x7733) : __byte_perm(x, y, 0x5511);
#ifdef CGQ_HACK_HALF_ALU
        const uint32_t a[4] = {Nib<T, kTrick>::lo(v0), v0, Nib<T, kTrick>::hi(v0), v0};
#else
        const uint32_t a[4] = {Nib<T, kTrick>::lo(v0), Nib<T, kTrick>::lo(v1),
                               Nib<T, kTrick>::hi(v0), Nib<T, kTrick>::hi(v1)};
#endif
        if (b == 0)
          ptx::mma_16816(grp[j], a, b0, b1, zero4, T());
        else
          ptx::mma_16816(grp[j], a, b0, b1, grp[j], T());
      }
    }
    // group scales of this lane's 16 columns
    const uint4 sv0 = ptx::lds128(srow), sv1 = ptx::lds128(srow + 16);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[slot]);  // all reads of this slot are in registers
    const uint32_t sw[8] = {sv0.x, sv0.y, sv0.z, sv0.w, sv1.x, sv1.y, sv1.z, sv1.w};
    float c0 = 0.f, c1 = 0.f;
    if (kTrick) {
      c0 = -8.f * ag[0];
      c1 = -8.f * ag[1];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // columns 2j (MMA row g -> c0,c1) and 2j+1 (MMA row g+8 -> c2,c3)
      union {
        uint32_t u;
        T h[2];
      } cv;
      cv.u = sw[j];
      const float sa = DT<T>::to_f(cv.h[0]), sb = DT<T>::to_f(cv.h[1]);
      if (kM1) {
        const float t0 = kTrick ? fmaf(grp[j][0], 16777216.f, c0) : grp[j][0];
        const float t2 = kTrick ? fmaf(grp[j][2], 16777216.f, c0) : grp[j][2];
        tot[j][0] = fmaf(sa, t0, tot[j][0]);
        tot[j][1] = fmaf(sb, t2, tot[j][1]);
      } else {
        const float t0 = kTrick ? fmaf(grp[j][0], 16777216.f, c0) : grp[j][0];
        const float t1 = kTrick ? fmaf(grp[j][1], 16777216.f, c1) : grp[j][1];
        const float t2 = kTrick ? fmaf(grp[j][2], 16777216.f, c0) : grp[j][2];
        const float t3 = kTrick ? fmaf(grp[j][3], 16777216.f, c1) : grp[j][3];
        tot[j][0] = fmaf(sa, t0, tot[j][0]);
        tot[j][1] = fmaf(sa, t1, tot[j][1]);
        tot[j][2] = fmaf(sb, t2, tot[j][2]);
        tot[j][3] = fmaf(sb, t3, tot[j][3]);
      }
    }
    if (++slot == S) {
      slot = 0;
      phase ^= 1;
    }

  }
  if (threadIdx.x == 0) stamp(p, 4);

  // ---------------- band sum of this CTA: cross-warp (k-group) reduction through shared memory
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int tok = kM1 ? 0 : 2 * tig + (i & 1);
      const int col = 16 * g + 2 * j + (kM1 ? i : (i >> 1));
      const bool ok = kM1 ? (tig == 0) : (tok < p.M);
      if (ok) red[(warp * MR + tok) * BN + col] = tot[j][i];
    }
  }
  ptx::named_bar_sync(1, CW * 32);
  {
    const int t = threadIdx.x;  // column within the tile
    float v[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      v[m] = 0.f;
      if (m < p.M) {
#pragma unroll
        for (int w = 0; w < CW; ++w) v[m] += red[(w * MR + m) * BN + t];
      }
    }
    if (Z == 1) {
      const int n = tile * BN + t;
      if (n < p.N) {
        T* Cp = static_cast<T*>(p.C);
        const T* bias = static_cast<const T*>(p.bias);
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < p.M) Cp[m * p.ldc + n] = epilogue<T>(v[m], bias, n);
      }
    } else {
      // push the band sum into rank 0's shared memory (DSMEM); rank 0 adds them in rank order
      const uint32_t local = ptx::smem_u32(xred) + static_cast<uint32_t>((z * MR) * BN + t) * 4u;
      const uint32_t remote = ptx::mapa_rank(local, 0);
#pragma unroll
      for (int m = 0; m < MR; ++m)
        if (m < p.M) ptx::st_cluster_f32(remote + m * BN * 4, v[m]);
    }
  }
  }  // consumers
  if (Z > 1) {
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
    if (z == 0 && threadIdx.x < BN) {
      const int t = threadIdx.x, n = tile * BN + t;
      if (n < p.N) {
        T* Cp = static_cast<T*>(p.C);
        const T* bias = static_cast<const T*>(p.bias);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          if (m < p.M) {
            float acc = 0.f;
            for (int zz = 0; zz < Z; ++zz) acc += xred[(zz * MR + m) * BN + t];
            Cp[m * p.ldc + n] = epilogue<T>(acc, bias, n);
          }
        }
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

template <typename T, bool kTrick, bool kM1>
int launch_inst(const GemmArgs& a, const CUtensorMap& tmW, const CUtensorMap& tmS, Params prm,
                int grid, int stages, bool pdl) {
  using C = Cfg<kM1>;
  const size_t smem = 1024 + static_cast<size_t>(stages) * C::STAGE_BYTES + C::RED_BYTES +
                      C::XRED_BYTES + 16 * stages + 16;
  auto kern = w4_gemv_kernel<T, kTrick, kM1>;
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
  if (prm.Z > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(prm.Z);
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

template <typename T>
int launch_t(const GemmArgs& a, bool exact) {
  const int G = a.K / 32;
  const int SPT = (G + CW - 1) / CW;
  const int tiles = (a.N + BN - 1) / BN;
  static const int stages_env = env_int("CGQ_GEMV_STAGES", 0, 0, 16);
  static const int z_env = env_int("CGQ_GEMV_Z", 0, 0, 8);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  // k-bands per tile: fill the CTA slots of the SMs in one wave, powers of two up to the portable
  // cluster size, never more bands than k-stages
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 4, 1, 4);
  const int slots = cps * sm_count();
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2) Z *= 2;
  if (z_env > 0) Z = z_env;
  if (Z > SPT) Z = 1;
  const int grid = tiles * Z;
  // fewer CTAs than slots -> deeper rings keep the same number of bytes in flight
  int stages = stages_env > 0 ? stages_env : (grid * 4 <= slots * 3 ? 6 : 4);
  const int per_cta = (SPT + Z - 1) / Z;
  if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;

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
  prm.lda = a.lda;
  prm.bias = a.bias;
  prm.C = a.C;
  prm.ldc = a.ldc;
  prm.M = a.M;
  prm.N = a.N;
  prm.K = a.K;
  prm.SPT = SPT;
  prm.Z = Z;
  prm.S = stages;
  prm.trace = static_cast<unsigned long long*>(take_trace_buffer());

  constexpr bool kIsHalf = (DT<T>::code == CGQ_DTYPE_F16);
  const bool trick = kIsHalf && !exact;
  if (a.M == 1) {
    if (trick) return launch_inst<T, kIsHalf, true>(a, tmW, tmS, prm, grid, stages, pdl);
    return launch_inst<T, false, true>(a, tmW, tmS, prm, grid, stages, pdl);
  }
  if (trick) return launch_inst<T, kIsHalf, false>(a, tmW, tmS, prm, grid, stages, pdl);
  return launch_inst<T, false, false>(a, tmW, tmS, prm, grid, stages, pdl);
}

}  // namespace

bool w4_gemv_supported(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.M >= 1 && a.M <= MMAX && a.K % 32 == 0 && a.N % 16 == 0 && al16(a.Wq) &&
         al16(a.scale) && al16(a.A) && a.lda % 8 == 0 && true;
}

int launch_w4_gemv(const GemmArgs& a, bool exact) {
  static const bool umma_default = env_int("CGQ_GEMV_UMMA", 0, 0, 1) != 0;
  if (umma_default && !exact && a.M == 1) {   // opt-in: integer tcgen05 decode kernel (gemv_w4_umma.cu)
    bool taken = false;
    const int rc = launch_w4_gemv_umma(a, &taken);
    if (rc != CGQ_OK || taken) return rc;
  }
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, exact) : launch_t<__nv_bfloat16>(a, exact);
}

}  // namespace cgq
