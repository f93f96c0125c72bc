// int4g32 decode kernel (M <= 8 token rows): C[M,N] = A[M,K] · ((nib(Wq) - 8) * scale).
// Replaces _dynamic_quant_matmul_s4_kernel (chatglm_q/int4/triton_ops.py:18-87) for decode.
//
// HBM-bound design (B200: 148 SMs, ~6.5 TB/s measured):
//   * persistent stream-K grid: the (column-tile, k-stage) units are cut into gridDim.x equal
//     contiguous ranges, so every SM streams the same number of bytes whatever N is;
//   * small CTAs (4 consumer warps + 1 producer warp, ~50 KB of shared memory at M=1), two per SM
//     per launch, so that the CTAs of the NEXT launch fit beside them: with programmatic dependent
//     launch the next kernel's producers fill their rings with weights (which do not depend on the
//     previous kernel) while this kernel is still computing — the HBM stream does not stop at
//     kernel boundaries, which is what a chain of 1.5 µs GEMVs needs;
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
//     extra MMA against a constant fragment;
//   * tiles cut by a range boundary are reduced deterministically: partial tiles go to a
//     workspace slot, the last arriver (self-cleaning counter) sums the slots in CTA order.
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
  static constexpr int STAGE_BYTES = W_BYTES + S_BYTES + A_BYTES;
};

static_assert(kSlotFloats == MMAX * BN, "workspace slot size");

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
  int U;       // total units = tiles * SPT
  int S;       // ring depth
  int* counters;
  float* partials;
  unsigned long long* trace;  // optional timeline (cgq_debug_trace), 8 words per CTA
};

__device__ __forceinline__ int unit_begin(int U, int P, int c) {
  return static_cast<int>(static_cast<int64_t>(U) * c / P);
}
__device__ __forceinline__ int unit_owner(int U, int P, int u) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * P - 1) / U);
}
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
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + C::RED_BYTES);
  uint64_t* empty = full + S;
  int* flag = reinterpret_cast<int*>(empty + S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = gridDim.x, c = blockIdx.x;
  const int u0 = unit_begin(p.U, P, c), u1 = unit_begin(p.U, P, c + 1);
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
        const int u = u0 + i, tile = u / p.SPT, ks = u - tile * p.SPT;
        const int kvalid = min(KSTAGE, p.K - ks * KSTAGE);
        ptx::mbar_expect_tx(&full[slot], W_BYTES + S_BYTES + p.M * kvalid * 2);
        ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, tile * BN, ks * ROWS, &full[slot], pol);
        ptx::tma_load_2d(gen + S * W_BYTES + slot * S_BYTES, &tmS, tile * BN, ks * CW, &full[slot],
                         pol);
      };
      auto issue_a = [&](int i, int slot) {
        const int u = u0 + i, tile = u / p.SPT, ks = u - tile * p.SPT;
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
    return;
  }

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
  int tile = u0 / p.SPT, ks = u0 - tile * p.SPT;  // tracked incrementally (no division in the loop)
  for (int it = 0; it < n_units; ++it) {
    ptx::mbar_wait(&full[slot], phase);
    if (it == 0 && threadIdx.x == 0) stamp(p, 3);
    const uint32_t wrow = Wsm + slot * W_BYTES + (16 * warp) * BN;
    const uint32_t srow = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 32;
    const uint32_t arow = Asm + slot * C::A_BYTES + (kM1 ? 0 : g * A_STRIDE) + (32 * warp + 4 * tig) * 2;

    float grp[8][4];
    float ag[4];
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int b = 0; b < 2; ++b) {
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
          ptx::mma_16816(ag, ones, b0, b1, z, T());
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
        const uint32_t a[4] = {Nib<T, kTrick>::lo(v0), Nib<T, kTrick>::lo(v1),
                               Nib<T, kTrick>::hi(v0), Nib<T, kTrick>::hi(v1)};
        if (b == 0)
          ptx::mma_16816(grp[j], a, b0, b1, z, T());
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

    // ---------------- end of a column tile (or of this CTA's range): reduce and emit
    if (ks == p.SPT - 1 || it == n_units - 1) {
      if (it == n_units - 1 && threadIdx.x == 0) stamp(p, 4);
      // (1) cross-warp (k-group) reduction through shared memory
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int i = 0; i < NT; ++i) {
          const int tok = kM1 ? 0 : 2 * tig + (i & 1);
          const int col = 16 * g + 2 * j + (kM1 ? i : (i >> 1));
          const bool ok = kM1 ? (tig == 0) : (tok < p.M);
          if (ok) red[(warp * MR + tok) * BN + col] = tot[j][i];
          tot[j][i] = 0.f;
        }
      }
      ptx::named_bar_sync(1, CW * 32);
      const int t = threadIdx.x;  // column within the tile
      const int n = tile * BN + t;
      float v[MR];
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        v[m] = 0.f;
        if (m < p.M) {
#pragma unroll
          for (int w = 0; w < CW; ++w) v[m] += red[(w * MR + m) * BN + t];
        }
      }
      // (2) is the tile cut by a range boundary?
      const int t_first = tile * p.SPT, t_last = t_first + p.SPT - 1;
      const bool whole = (u0 <= t_first) && (u1 > t_last);
      bool write_out = whole;
      if (!whole) {
        const int my_slot = c * 2 + ((tile == u0 / p.SPT) ? 0 : 1);
        float* mine = p.partials + static_cast<size_t>(my_slot) * kSlotFloats;
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < p.M) mine[m * BN + t] = v[m];
        ptx::named_bar_sync(1, CW * 32);  // every partial of this CTA is issued ...
        const int c_first = unit_owner(p.U, P, t_first), c_last = unit_owner(p.U, P, t_last);
        if (t == 0) {
          ptx::fence_acq_rel_gpu();       // ... and made visible (cumulative) before the count
          const int old = atomicAdd(&p.counters[tile * kCounterStride], 1);
          const int last = (old == c_last - c_first) ? 1 : 0;
          if (last) {
            p.counters[tile * kCounterStride] = 0;         // self-cleaning: every contributor has arrived
            ptx::fence_acq_rel_gpu();     // acquire side: the others' partials are visible
          }
          *flag = last;
          if (it == n_units - 1) stamp(p, 6);
        }
        ptx::named_bar_sync(1, CW * 32);
        write_out = (*flag != 0);
        if (write_out) {
          // Contributor cc > c_first starts inside this tile (its first tile -> slot 0); c_first
          // uses slot 1 iff its range began in an earlier tile.
          const int first_slot = (unit_begin(p.U, P, c_first) < t_first) ? 1 : 0;
#pragma unroll
          for (int m = 0; m < MR; ++m) v[m] = 0.f;
          // fixed order -> deterministic sum; loads are issued in batches so that their L2
          // latencies overlap instead of adding up (the tile has up to ~10 contributors)
          constexpr int CH = kM1 ? 8 : 2;
          for (int cb = c_first; cb <= c_last; cb += CH) {
            float ld[CH][MR];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const int cc = cb + i;
              const int sl = cc * 2 + (cc == c_first ? first_slot : 0);
              const float* src = p.partials + static_cast<size_t>(sl) * kSlotFloats;
#pragma unroll
              for (int m = 0; m < MR; ++m)
                ld[i][m] = (cc <= c_last && m < p.M) ? ptx::ldcg_f32(src + m * BN + t) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < CH; ++i)
#pragma unroll
              for (int m = 0; m < MR; ++m) v[m] += ld[i][m];
          }
          if (t == 0) stamp(p, 7);
        }
      }
      if (write_out && n < p.N) {
        T* Cp = static_cast<T*>(p.C);
        const T* bias = static_cast<const T*>(p.bias);
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < p.M) Cp[m * p.ldc + n] = epilogue<T>(v[m], bias, n);
      }
      ptx::named_bar_sync(1, CW * 32);  // red[] / flag may be reused
    }
    if (++ks == p.SPT) {
      ks = 0;
      ++tile;
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
  const size_t smem =
      1024 + static_cast<size_t>(stages) * C::STAGE_BYTES + C::RED_BYTES + 16 * stages + 16;
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
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmW, tmS, prm));
  return CGQ_OK;
}

template <typename T>
int launch_t(const GemmArgs& a, bool exact) {
  const int G = a.K / 32;
  const int SPT = (G + CW - 1) / CW;
  const int tiles = (a.N + BN - 1) / BN;
  const int U = tiles * SPT;
  static const int stages = env_int("CGQ_GEMV_STAGES", 4, 2, 16);
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 2, 1, 4);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  int grid = sm_count() * cps;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid > U) grid = U;

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
  prm.U = U;
  prm.S = stages;
  prm.counters = static_cast<int*>(a.workspace);
  prm.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + kCounterBytes);
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
  const int tiles = (a.N + BN - 1) / BN;
  return a.M >= 1 && a.M <= MMAX && a.K % 32 == 0 && a.N % 16 == 0 && al16(a.Wq) &&
         al16(a.scale) && al16(a.A) && a.lda % 8 == 0 && tiles <= kMaxTiles;
}

int launch_w4_gemv(const GemmArgs& a, bool exact) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, exact) : launch_t<__nv_bfloat16>(a, exact);
}

}  // namespace cgq
