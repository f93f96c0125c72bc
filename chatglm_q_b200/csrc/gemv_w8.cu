// int8 per-channel decode kernel (M <= 8 token rows): C[M,N] = A[M,K] · (Wq[N,K]^T * scale[N]).
// Replaces _dynamic_quant_matmul_kernel (chatglm_q/int8/triton_ops.py:13-84) for decode.
//
// Same skeleton as gemv_w4.cu: persistent stream-K grid over (64-column tile, 128-k stage) units,
// TMA ring (a [64 n x 128 k] byte tile with 128-byte swizzle + the activation slice), 4 consumer
// warps.  The weight is K-contiguous, so a 16-byte run of one row is 8 ready-made (k, k+1) pairs:
// int8 -> fp16 is `xor 0x80`, one PRMT against 0x64 (=> 1024+128+q) and one exact fp16 subtract.
// Each warp owns 16 output columns (MMA rows) of the tile, so no cross-warp reduction is needed;
// the per-channel scale multiplies the fp32 sum once in the epilogue (the reference rounds
// q*scale to fp16 per element: ours is the more accurate side of the 1e-2 parity bar).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "w4_dev.cuh"

namespace cgq {
namespace {

constexpr int BN8 = 64;             // output columns (weight rows) per tile
constexpr int CW = 4;               // consumer warps, 16 columns each
constexpr int KSTAGE = 128;         // k bytes per stage (TMA inner box)
constexpr int MMAX = 8;
constexpr int A_STRIDE = KSTAGE * 2 + 32;
constexpr int W_BYTES = BN8 * KSTAGE;
constexpr int A_BYTES = MMAX * A_STRIDE;
constexpr int RED_BYTES = MMAX * BN8 * 4;
constexpr int kThreads = (CW + 1) * 32;

__device__ __forceinline__ uint32_t h2_sub(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// word = 4 consecutive int8 (k..k+3) -> (k,k+1) and (k+2,k+3) as packed T pairs, exact.
template <typename T>
struct Cvt8;
template <>
struct Cvt8<__half> {
  __device__ static __forceinline__ void run(uint32_t w, uint32_t& p01, uint32_t& p23) {
    const uint32_t x = w ^ 0x80808080u;  // q + 128 as unsigned bytes
    p01 = h2_sub(__byte_perm(x, 0x64646464u, 0x4140), 0x64806480u);  // (1024 + 128 + q) - 1152
    p23 = h2_sub(__byte_perm(x, 0x64646464u, 0x4342), 0x64806480u);
  }
};
template <>
struct Cvt8<__nv_bfloat16> {
  // bf16 has 8 significant bits: go through fp32 (2^23 + 128 + q) - (2^23 + 128), keep the top half
  __device__ static __forceinline__ uint32_t one(uint32_t x, int sel) {
    const float f = __uint_as_float(__byte_perm(x, 0x4B000000u, sel)) - 8388736.f;
    return __float_as_uint(f);
  }
  __device__ static __forceinline__ void run(uint32_t w, uint32_t& p01, uint32_t& p23) {
    const uint32_t x = w ^ 0x80808080u;
    const uint32_t f0 = one(x, 0x7440), f1 = one(x, 0x7441), f2 = one(x, 0x7442),
                   f3 = one(x, 0x7443);
    p01 = __byte_perm(f0, f1, 0x7632);  // high halves: (bf16(q0), bf16(q1))
    p23 = __byte_perm(f2, f3, 0x7632);
  }
};

struct Params {
  const void* A;
  int64_t lda;
  const void* scale;
  const void* bias;
  void* C;
  int64_t ldc;
  int M, N, K;
  int SPT, U, S;
  int* counters;
  float* partials;
};

__device__ __forceinline__ int unit_begin(int U, int P, int c) {
  return static_cast<int>(static_cast<int64_t>(U) * c / P);
}
__device__ __forceinline__ int unit_owner(int U, int P, int u) {
  return static_cast<int>((static_cast<int64_t>(u + 1) * P - 1) / U);
}

template <typename T, bool kM1>
__global__ void __launch_bounds__(kThreads)
    w8_gemv_kernel(const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int S = p.S;
  const uint32_t Wsm = base;
  const uint32_t Asm = Wsm + S * W_BYTES;
  const uint32_t off_red = S * (W_BYTES + A_BYTES);
  float* red = reinterpret_cast<float*>(gen + off_red);
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + RED_BYTES);
  uint64_t* empty = full + S;
  int* flag = reinterpret_cast<int*>(empty + S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = gridDim.x, c = blockIdx.x;
  const int u0 = unit_begin(p.U, P, c), u1 = unit_begin(p.U, P, c + 1);
  const int n_units = u1 - u0;
  const T* A = static_cast<const T*>(p.A);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < S * A_BYTES / 16; i += kThreads)
    ptx::sts128(Asm + i * 16, make_uint4(0, 0, 0, 0));
  ptx::fence_proxy_async_smem();
  __syncthreads();
  ptx::pdl_launch_dependents();

  if (warp == CW) {
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      auto issue_w = [&](int i, int slot) {
        const int u = u0 + i, tile = u / p.SPT, ks = u - tile * p.SPT;
        const int kvalid = min(KSTAGE, p.K - ks * KSTAGE);
        ptx::mbar_expect_tx(&full[slot], W_BYTES + p.M * kvalid * 2);
        ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, ks * KSTAGE, tile * BN8, &full[slot], pol);
      };
      auto issue_a = [&](int i, int slot) {
        const int u = u0 + i, tile = u / p.SPT, ks = u - tile * p.SPT;
        const int kvalid = min(KSTAGE, p.K - ks * KSTAGE);
        uint8_t* dst = gen + S * W_BYTES + slot * A_BYTES;
        for (int m = 0; m < p.M; ++m)
          ptx::bulk_load_1d(dst + m * A_STRIDE, A + m * p.lda + ks * KSTAGE, kvalid * 2,
                            &full[slot]);
      };
      const int prefill = min(n_units, S);
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

  ptx::pdl_wait_prior_grid();
  const int g = lane >> 2, tig = lane & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float acc2[4] = {0.f, 0.f, 0.f, 0.f};  // second chain: halves the HMMA dependency depth
  const bool has_tok = kM1 ? (g == 0) : (g < p.M);
  int slot = 0, phase = 0;
  int tile = u0 / p.SPT, ks = u0 - tile * p.SPT;  // tracked incrementally (no division in the loop)
  for (int it = 0; it < n_units; ++it) {
    ptx::mbar_wait(&full[slot], phase);
    const uint32_t wa = Wsm + slot * W_BYTES + (16 * warp + g) * KSTAGE;  // row g of this warp
    const uint32_t wb = wa + 8 * KSTAGE;                                   // row g + 8
    const uint32_t arow = Asm + slot * A_BYTES + g * A_STRIDE + (32 * tig) * 2;
    uint32_t ld_dep = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int chunk = 2 * tig + h;  // 16-byte run: k = 16*chunk .. +15
      const uint4 qa = ptx::lds128(wa + ((chunk ^ g) << 4));
      const uint4 qb = ptx::lds128(wb + ((chunk ^ g) << 4));
      uint4 a0 = make_uint4(0, 0, 0, 0), a1 = make_uint4(0, 0, 0, 0);
      if (has_tok) {
        a0 = ptx::lds128(arow + 32 * h);
        a1 = ptx::lds128(arow + 32 * h + 16);
      }
      ld_dep = qa.x | qb.x | a0.x | a1.x;
      const uint32_t wqa[4] = {qa.x, qa.y, qa.z, qa.w};
      const uint32_t wqb[4] = {qb.x, qb.y, qb.z, qb.w};
      const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t fr[4];
        Cvt8<T>::run(wqa[e], fr[0], fr[2]);  // row g:   slots (2t,2t+1) and (2t+8,2t+9)
        Cvt8<T>::run(wqb[e], fr[1], fr[3]);  // row g+8
        if (e & 1)
          ptx::mma_16816(acc2, fr, av[2 * e], av[2 * e + 1], acc2, T());
        else
          ptx::mma_16816(acc, fr, av[2 * e], av[2 * e + 1], acc, T());
      }
    }
    __syncwarp();
    // release the slot only once every load from it has returned (ptx::mbar_arrive_after_loads)
    if (lane == 0)
      ptx::mbar_arrive_after_loads(&empty[slot], ld_dep, static_cast<uint32_t>(p.K) >> 31);
    if (++slot == S) {
      slot = 0;
      phase ^= 1;
    }

    if (ks == p.SPT - 1 || it == n_units - 1) {
      // D fragment: acc[0..1] = (column 16w+g, tokens 2t, 2t+1), acc[2..3] = column 16w+g+8
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = 2 * tig + (i & 1);
        const int col = 16 * warp + g + 8 * (i >> 1);
        if (tok < p.M) red[tok * BN8 + col] = acc[i] + acc2[i];
        acc[i] = 0.f;
        acc2[i] = 0.f;
      }
      ptx::named_bar_sync(1, CW * 32);
      const int t = threadIdx.x;
      const bool colthread = t < BN8;
      const int n = tile * BN8 + t;
      float v[MMAX];
#pragma unroll
      for (int m = 0; m < MMAX; ++m) v[m] = (colthread && m < p.M) ? red[m * BN8 + t] : 0.f;

      const int t_first = tile * p.SPT, t_last = t_first + p.SPT - 1;
      const bool whole = (u0 <= t_first) && (u1 > t_last);
      bool write_out = whole;
      if (!whole) {
        const int my_slot = c * 2 + ((tile == u0 / p.SPT) ? 0 : 1);
        float* mine = p.partials + static_cast<size_t>(my_slot) * kSlotFloats;
        if (colthread) {
#pragma unroll
          for (int m = 0; m < MMAX; ++m)
            if (m < p.M) mine[m * BN8 + t] = v[m];
        }
        __threadfence();
        ptx::named_bar_sync(1, CW * 32);
        const int c_first = unit_owner(p.U, P, t_first), c_last = unit_owner(p.U, P, t_last);
        if (t == 0) {
          const int old = atomicAdd(&p.counters[tile * kCounterStride], 1);
          const int last = (old == c_last - c_first) ? 1 : 0;
          if (last) p.counters[tile * kCounterStride] = 0;
          *flag = last;
        }
        ptx::named_bar_sync(1, CW * 32);
        write_out = (*flag != 0);
        if (write_out && colthread) {
          __threadfence();
#pragma unroll
          for (int m = 0; m < MMAX; ++m) v[m] = 0.f;
          for (int cc = c_first; cc <= c_last; ++cc) {
            const int sl = cc * 2 + ((tile == unit_begin(p.U, P, cc) / p.SPT) ? 0 : 1);
            const float* src = p.partials + static_cast<size_t>(sl) * kSlotFloats;
#pragma unroll
            for (int m = 0; m < MMAX; ++m)
              if (m < p.M) v[m] += ptx::ldcg_f32(src + m * BN8 + t);
          }
        }
      }
      if (write_out && colthread && n < p.N) {
        T* Cp = static_cast<T*>(p.C);
        const T* bias = static_cast<const T*>(p.bias);
        const float s = DT<T>::to_f(static_cast<const T*>(p.scale)[n]);
#pragma unroll
        for (int m = 0; m < MMAX; ++m)
          if (m < p.M) Cp[m * p.ldc + n] = epilogue<T>(v[m] * s, bias, n);
      }
      ptx::named_bar_sync(1, CW * 32);
    }
    if (++ks == p.SPT) {
      ks = 0;
      ++tile;
    }
  }
}

int env_int(const char* name, int dflt, int lo, int hi) {
  const char* s = getenv(name);
  if (s == nullptr || *s == 0) return dflt;
  int v = atoi(s);
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return v;
}

template <typename T, bool kM1>
int launch_inst(const GemmArgs& a, const CUtensorMap& tmW, Params prm, int grid, size_t smem,
                bool pdl) {
  auto kern = w8_gemv_kernel<T, kM1>;
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
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmW, prm));
  return CGQ_OK;
}

template <typename T>
int launch_t(const GemmArgs& a) {
  const int SPT = (a.K + KSTAGE - 1) / KSTAGE;
  const int tiles = (a.N + BN8 - 1) / BN8;
  const int U = tiles * SPT;
  static const int stages = env_int("CGQ_GEMV_STAGES", 8, 2, 16);
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 2, 1, 4);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  int grid = sm_count() * cps;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid > U) grid = U;

  CUtensorMap tmW;
  TmapKey kw{a.Wq, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N),
             static_cast<uint64_t>(a.K), KSTAGE, BN8, CU_TENSOR_MAP_DATA_TYPE_UINT8,
             CU_TENSOR_MAP_SWIZZLE_128B};
  int rc = get_tmap_2d(kw, &tmW);
  if (rc != CGQ_OK) return rc;

  Params prm;
  prm.A = a.A;
  prm.lda = a.lda;
  prm.scale = a.scale;
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
  const size_t smem =
      1024 + static_cast<size_t>(stages) * (W_BYTES + A_BYTES) + RED_BYTES + 16 * stages + 16;
  if (a.M == 1) return launch_inst<T, true>(a, tmW, prm, grid, smem, pdl);
  return launch_inst<T, false>(a, tmW, prm, grid, smem, pdl);
}


// =====================================================================================================
// M == 1 int8 decode kernel on gemv_w4.cu's skeleton (round 2): (64-column tile) x (Z contiguous k-bands), the Z
// CTAs of a tile form a thread-block cluster and reduce their band sums through distributed shared memory; the
// TMA ring carries weights only, the activation band is staged once per CTA by the consumer warps through the
// fused prologue (RMSNorm / SiLU*gate), the residual is added in the epilogue (cgq_w8a16_gemv_fused).  Replaces
// the stream-K workspace fix-up for one token: qkv / o_proj took 11.6 - 11.9 us with it.
namespace m1 {

using w4::ldcg128;
using w4::ldnc128;
using w4::PRO_NONE;
using w4::PRO_RMSNORM;
using w4::PRO_SILU_GATE;

struct P1 {
  const void* A;
  const void* scale;
  const void* bias;
  const void* resid;
  const void* norm_w;
  void* C;
  int N, K, SPT, Z, S, band_units;
  float eps;
};

template <typename T, int kPro>
__global__ void __launch_bounds__(kThreads, 4)
    w8_gemv_m1_kernel(const __grid_constant__ CUtensorMap tmW, const P1 p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int S = p.S, Z = p.Z;
  const uint32_t Wsm = base;
  const uint32_t off_x = S * W_BYTES;
  float* xred = reinterpret_cast<float*>(gen + off_x);                      // [Z][64] band sums (used on rank 0)
  float* sred = xred + 8 * BN8;                                             // [CW] sum(x^2) partials
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_x + 8 * BN8 * 4 + 64);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;   // rank 0: completes when the band sums of ranks 1 .. Z-1 have landed in xred
  const uint32_t Aband = base + ((off_x + 8 * BN8 * 4 + 64 + 16 * S + 16 + 15) & ~15u);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;
  const int n_units = u1 - u0;
  const T* A = static_cast<const T*>(p.A);

  if (threadIdx.x == CW * 32) {
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    if (Z > 1 && z == 0) {     // the remote st.async stores carry the arrival: two columns per consumer quad
      ptx::mbar_init(xbar, 1);
      ptx::mbar_expect_tx(xbar, static_cast<uint32_t>(Z - 1) * BN8 * 4u);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  __syncthreads();
  if (Z > 1) ptx::cluster_arrive_release();      // phase A: waited for before the first DSMEM store
  ptx::pdl_launch_dependents();

  if (warp == CW) {
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      auto issue = [&](int i, int slot) {
        ptx::mbar_expect_tx(&full[slot], W_BYTES);
        ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, (u0 + i) * KSTAGE, tile * BN8, &full[slot], pol);
      };
      const int prefill = min(n_units, S);
      for (int i = 0; i < prefill; ++i) issue(i, i);        // weights do not depend on the previous kernel
      int slot = 0, phase = 1;
      for (int i = prefill; i < n_units; ++i) {
        ptx::mbar_wait(&empty[slot], phase ^ 1);
        issue(i, slot);
        if (++slot == S) {
          slot = 0;
          phase ^= 1;
        }
      }
    }
    __syncwarp();
    if (Z > 1) ptx::cluster_wait_acquire();
  } else {
    // ---- stage the activation band (16-byte chunks of 8 k, zero beyond K) through the fused prologue
    const int tid = threadIdx.x;
    const int nchunk = p.K >> 3;
    const int c_lo = u0 * (KSTAGE / 8), c_hi = u1 * (KSTAGE / 8);
    const uint4 zero = make_uint4(0, 0, 0, 0);
    if (kPro == PRO_RMSNORM) {
      const T* nw = static_cast<const T*>(p.norm_w);
      ptx::pdl_wait_prior_grid();
      float ss = 0.f;
      for (int c = tid; c < nchunk; c += CW * 32) ss += w4::sumsq8<T>(ldcg128(A + c * 8));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) sred[warp] = ss;
      ptx::named_bar_sync(1, CW * 32);
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < CW; ++w) tot += sred[w];
      const float rstd = rsqrtf(tot / static_cast<float>(p.K) + p.eps);
      for (int c = c_lo + tid; c < c_hi; c += CW * 32)
        ptx::sts128(Aband + (c - c_lo) * 16,
                    c < nchunk ? w4::rmsnorm8<T>(ldcg128(A + c * 8), ldnc128(nw + c * 8), rstd) : zero);
    } else {
      ptx::pdl_wait_prior_grid();
      for (int c = c_lo + tid; c < c_hi; c += CW * 32) {
        uint4 v = zero;
        if (c < nchunk) {
          v = ldcg128(A + c * 8);
          if (kPro == PRO_SILU_GATE) v = w4::silu_gate8<T>(v, ldcg128(A + p.K + c * 8));
        }
        ptx::sts128(Aband + (c - c_lo) * 16, v);
      }
    }
    // scale / bias / residual of the column this thread will store (threads 0 .. 63 of the rank that stores), requested
    // now: the tail of a 4 us launch should not wait for three more L2 round trips
    float pre_s = 0.f;
    T pre_b = DT<T>::from_f(0.f), pre_r = DT<T>::from_f(0.f);
    if ((Z == 1 || z == 0) && threadIdx.x < BN8 && tile * BN8 + static_cast<int>(threadIdx.x) < p.N) {
      const int n = tile * BN8 + threadIdx.x;
      pre_s = DT<T>::to_f(static_cast<const T*>(p.scale)[n]);
      if (p.bias != nullptr) pre_b = static_cast<const T*>(p.bias)[n];
      if (p.resid != nullptr) pre_r = w4::ldcg_t<T>(static_cast<const T*>(p.resid) + n);
    }
    ptx::named_bar_sync(1, CW * 32);

    const int g = lane >> 2, tig = lane & 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float acc2[4] = {0.f, 0.f, 0.f, 0.f};
    int slot = 0, phase = 0;
    for (int it = 0; it < n_units; ++it) {
      ptx::mbar_wait(&full[slot], phase);
      const uint32_t wa = Wsm + slot * W_BYTES + (16 * warp + g) * KSTAGE;  // row g of this warp
      const uint32_t wb = wa + 8 * KSTAGE;                                   // row g + 8
      const uint32_t arow = Aband + (it * KSTAGE + 32 * tig) * 2;
      uint32_t ld_dep = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int chunk = 2 * tig + h;  // 16-byte run: k = 16*chunk .. +15
        const uint4 qa = ptx::lds128(wa + ((chunk ^ g) << 4));
        const uint4 qb = ptx::lds128(wb + ((chunk ^ g) << 4));
        uint4 a0 = zero, a1 = zero;
        if (g == 0) {                       // one token: only MMA column 0 carries data
          a0 = ptx::lds128(arow + 32 * h);
          a1 = ptx::lds128(arow + 32 * h + 16);
        }
        ld_dep |= qa.x | qb.x;
        const uint32_t wqa[4] = {qa.x, qa.y, qa.z, qa.w};
        const uint32_t wqb[4] = {qb.x, qb.y, qb.z, qb.w};
        const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t fr[4];
          Cvt8<T>::run(wqa[e], fr[0], fr[2]);
          Cvt8<T>::run(wqb[e], fr[1], fr[3]);
          if (e & 1)
            ptx::mma_16816(acc2, fr, av[2 * e], av[2 * e + 1], acc2, T());
          else
            ptx::mma_16816(acc, fr, av[2 * e], av[2 * e + 1], acc, T());
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_after_loads(&empty[slot], ld_dep, static_cast<uint32_t>(p.K) >> 31);
      if (++slot == S) {
        slot = 0;
        phase ^= 1;
      }
    }
    // D fragment, token 0 (lanes tig == 0): acc[0] = column 16 w + g, acc[2] = column 16 w + g + 8
    const float v0 = acc[0] + acc2[0], v1 = acc[2] + acc2[2];
    const int c0 = 16 * warp + g, c1 = c0 + 8;
    if (Z > 1) ptx::cluster_wait_acquire();            // every CTA of the cluster runs (phase A)
    if (Z > 1 && z != 0) {
      // band sums meet in rank 0's shared memory: ranks 1 .. Z-1 push theirs with st.async onto rank 0's mbarrier (the
      // data is its own arrival signal: no closing cluster barrier, the pushing CTAs exit at once), as gemv_w4.cu does
      if (tig == 0) {
        const uint32_t remote = ptx::mapa_rank(ptx::smem_u32(xred) + static_cast<uint32_t>(z * BN8) * 4u, 0);
        const uint32_t rbar = ptx::mapa_rank(ptx::smem_u32(xbar), 0);
        ptx::st_async_cluster_f32(remote + c0 * 4, v0, rbar);
        ptx::st_async_cluster_f32(remote + c1 * 4, v1, rbar);
      }
    } else {
      if (tig == 0) {
        xred[c0] = v0;
        xred[c1] = v1;
      }
      ptx::named_bar_sync(1, CW * 32);
      if (threadIdx.x < BN8) {
        if (Z > 1) ptx::mbar_wait(xbar, 0);
        const int t = threadIdx.x, n = tile * BN8 + t;
        if (n < p.N) {
          float acc = 0.f;
          for (int zz = 0; zz < Z; ++zz) acc += xred[zz * BN8 + t];       // rank order: deterministic
          // round(acc * scale) -> (+ bias, rounded) -> (+ residual, rounded), on the values requested after staging
          T val = DT<T>::from_f(acc * pre_s);
          if (p.bias != nullptr) val = DT<T>::from_f(DT<T>::to_f(val) + DT<T>::to_f(pre_b));
          if (p.resid != nullptr) val = DT<T>::from_f(DT<T>::to_f(pre_r) + DT<T>::to_f(val));
          static_cast<T*>(p.C)[n] = val;
        }
      }
    }
  }
}

template <typename T, int kPro>
int launch_pro(const GemmArgs& a, const CUtensorMap& tmW, const P1& prm, int grid, bool pdl) {
  const size_t smem = 1024 + static_cast<size_t>(prm.S) * W_BYTES + 8 * BN8 * 4 + 64 + 16 * prm.S + 48 +
                      static_cast<size_t>(prm.band_units) * KSTAGE * 2;
  auto kern = w8_gemv_m1_kernel<T, kPro>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && smem > configured[dev]) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmW, prm));
  return CGQ_OK;
}

template <typename T>
int launch(const GemmArgs& a, const GemvFused* fu) {
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  const int SPT = (a.K + KSTAGE - 1) / KSTAGE;
  const int tiles = (a.N + BN8 - 1) / BN8;
  const int slots = 4 * sm_count();
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2) Z *= 2;
  const int grid = tiles * Z;
  const int per_cta = (SPT + Z - 1) / Z;
  // ring depth by CTAs per SM, as in gemv_w4.cu (launch_t)
  static const int s_four = env_int("CGQ_W8_STAGES_4PERSM", 5, 2, 16);
  static const int s_small = env_int("CGQ_W8_STAGES_SMALL", 6, 2, 16);
  static const int s_two = env_int("CGQ_W8_STAGES_2PERSM", 0, 0, 16);
  int stages = grid * 4 <= slots * 3 ? s_small : s_four;
  if (s_two > 0 && grid * 2 <= slots) stages = s_two;
  if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
  CUtensorMap tmW;
  TmapKey kw{a.Wq, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K), KSTAGE, BN8,
             CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_128B};
  int rc = get_tmap_2d(kw, &tmW);
  if (rc != CGQ_OK) return rc;
  P1 prm;
  prm.A = a.A;
  prm.scale = a.scale;
  prm.bias = a.bias;
  prm.resid = fu != nullptr ? fu->resid : nullptr;
  prm.norm_w = fu != nullptr ? fu->norm_w : nullptr;
  prm.eps = fu != nullptr ? fu->eps : 0.f;
  prm.C = a.C;
  prm.N = a.N;
  prm.K = a.K;
  prm.SPT = SPT;
  prm.Z = Z;
  prm.S = stages;
  prm.band_units = per_cta;
  const int pro = fu != nullptr ? fu->prologue : PRO_NONE;
  switch (pro) {
    case PRO_RMSNORM:
      return launch_pro<T, PRO_RMSNORM>(a, tmW, prm, grid, pdl);
    case PRO_SILU_GATE:
      return launch_pro<T, PRO_SILU_GATE>(a, tmW, prm, grid, pdl);
    default:
      return launch_pro<T, PRO_NONE>(a, tmW, prm, grid, pdl);
  }
}

}  // namespace m1

// =====================================================================================================
// M = 2 .. 8 int8 decode kernel (round 2b) on the same skeleton: (64-column tile) x (Z k-bands) as a cluster, band
// sums through distributed shared memory in rank order.  The activations ride in the TMA ring next to the weights:
// two [8 tokens x 64 k] boxes with the 128-byte swizzle per stage (rows >= M are zero-filled by the tensor map), so a
// lane's B fragment -- token g, 16 consecutive k -- is two conflict-free 16-byte loads.  Same MMA count as one
// token (the eight MMA columns are the tokens).  Replaces the stream-K workspace kernel for bs = 2 .. 8:
// BASELINE config 4's bs = 8 decode step 2.82 ms -> see DESIGN.md §5.0.
namespace mx {

constexpr int AX_BYTES = 2 * 8 * 128;   // activation tile of a stage

struct PX {
  const void* scale;
  const void* bias;
  void* C;
  int64_t ldc;
  int M, N, K, SPT, Z, S;
};

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
    w8_gemv_mx_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const PX p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const int S = p.S, Z = p.Z;
  const uint32_t Wsm = base;
  const uint32_t Asm = base + S * W_BYTES;
  const uint32_t off_x = S * (W_BYTES + AX_BYTES);
  float* xred = reinterpret_cast<float*>(gen + off_x);                      // [Z][8][64] band sums (used on rank 0)
  const uint32_t xbytes = static_cast<uint32_t>(Z) * MMAX * BN8 * 4;
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_x + xbytes);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;   // rank 0: completes when the band sums of ranks 1 .. Z-1 have landed in xred

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;
  const int n_units = u1 - u0;

  if (threadIdx.x == CW * 32) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmA);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    if (Z > 1 && z == 0) {
      ptx::mbar_init(xbar, 1);
      ptx::mbar_expect_tx(xbar, static_cast<uint32_t>(Z - 1) * static_cast<uint32_t>(p.M) * BN8 * 4u);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  __syncthreads();
  if (Z > 1) ptx::cluster_arrive_release();      // phase A: waited for before the first DSMEM store
  ptx::pdl_launch_dependents();

  if (warp == CW) {
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      const uint64_t keep = ptx::policy_evict_last();
      auto issue_w = [&](int i, int slot) {
        ptx::mbar_expect_tx(&full[slot], W_BYTES + AX_BYTES);
        ptx::tma_load_2d(gen + slot * W_BYTES, &tmW, (u0 + i) * KSTAGE, tile * BN8, &full[slot], pol);
      };
      auto issue_a = [&](int i, int slot) {
        uint8_t* dst = gen + S * W_BYTES + slot * AX_BYTES;
        ptx::tma_load_2d(dst, &tmA, (u0 + i) * KSTAGE, 0, &full[slot], keep);
        ptx::tma_load_2d(dst + 8 * 128, &tmA, (u0 + i) * KSTAGE + 64, 0, &full[slot], keep);
      };
      const int prefill = min(n_units, S);
      for (int i = 0; i < prefill; ++i) issue_w(i, i);      // weights do not depend on the previous kernel
      ptx::pdl_wait_prior_grid();                            // the activations do
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
    if (Z > 1) ptx::cluster_wait_acquire();
  } else {
    const int g = lane >> 2, tig = lane & 3;
    // scale / bias of the column this thread will store (threads 0 .. 63 of the storing rank): constants, requested now
    float pre_s = 0.f;
    T pre_b = DT<T>::from_f(0.f);
    if ((Z == 1 || z == 0) && threadIdx.x < BN8 && tile * BN8 + static_cast<int>(threadIdx.x) < p.N) {
      const int n = tile * BN8 + threadIdx.x;
      pre_s = DT<T>::to_f(static_cast<const T*>(p.scale)[n]);
      if (p.bias != nullptr) pre_b = static_cast<const T*>(p.bias)[n];
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float acc2[4] = {0.f, 0.f, 0.f, 0.f};
    int slot = 0, phase = 0;
    for (int it = 0; it < n_units; ++it) {
      ptx::mbar_wait(&full[slot], phase);
      const uint32_t wa = Wsm + slot * W_BYTES + (16 * warp + g) * KSTAGE;  // row g of this warp
      const uint32_t wb = wa + 8 * KSTAGE;                                   // row g + 8
      // token g, k = 32 tig + 16 h .. + 15: box (tig >> 1), 16-byte chunks (4 tig + 2 h) % 8 and + 1, swizzled by g
      const uint32_t arow = Asm + slot * AX_BYTES + (tig >> 1) * (8 * 128) + g * 128;
      uint32_t ld_dep = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int chunk = 2 * tig + h;  // 16-byte run of the weight row: k = 16*chunk .. +15
        const uint4 qa = ptx::lds128(wa + ((chunk ^ g) << 4));
        const uint4 qb = ptx::lds128(wb + ((chunk ^ g) << 4));
        const int ac = (4 * tig + 2 * h) & 7;
        const uint4 a0 = ptx::lds128(arow + ((ac ^ g) << 4));
        const uint4 a1 = ptx::lds128(arow + (((ac + 1) ^ g) << 4));
        ld_dep |= qa.x | qb.x | a0.x | a1.x;
        const uint32_t wqa[4] = {qa.x, qa.y, qa.z, qa.w};
        const uint32_t wqb[4] = {qb.x, qb.y, qb.z, qb.w};
        const uint32_t av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t fr[4];
          Cvt8<T>::run(wqa[e], fr[0], fr[2]);
          Cvt8<T>::run(wqb[e], fr[1], fr[3]);
          if (e & 1)
            ptx::mma_16816(acc2, fr, av[2 * e], av[2 * e + 1], acc2, T());
          else
            ptx::mma_16816(acc, fr, av[2 * e], av[2 * e + 1], acc, T());
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_after_loads(&empty[slot], ld_dep, static_cast<uint32_t>(p.K) >> 31);
      if (++slot == S) {
        slot = 0;
        phase ^= 1;
      }
    }
    // D fragment: tokens 2 tig, 2 tig + 1 of columns 16 w + g (acc[0], acc[1]) and 16 w + g + 8 (acc[2], acc[3])
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = acc[i] + acc2[i];
    const int c0 = 16 * warp + g;
    if (Z > 1) ptx::cluster_wait_acquire();            // every CTA of the cluster runs (phase A)
    if (Z > 1 && z != 0) {
      // ranks 1 .. Z-1 push their band sums with st.async onto rank 0's mbarrier and exit; rank 0 adds in rank order
      const uint32_t remote = ptx::mapa_rank(ptx::smem_u32(xred) + static_cast<uint32_t>(z * MMAX * BN8) * 4u, 0);
      const uint32_t rbar = ptx::mapa_rank(ptx::smem_u32(xbar), 0);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = 2 * tig + (i & 1);
        if (tok < p.M)
          ptx::st_async_cluster_f32(remote + static_cast<uint32_t>(tok * BN8 + c0 + 8 * (i >> 1)) * 4u, v[i], rbar);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int tok = 2 * tig + (i & 1);
        if (tok < p.M) xred[tok * BN8 + c0 + 8 * (i >> 1)] = v[i];
      }
      ptx::named_bar_sync(1, CW * 32);
      if (threadIdx.x < BN8) {
        if (Z > 1) ptx::mbar_wait(xbar, 0);
        const int t = threadIdx.x, n = tile * BN8 + t;
        if (n < p.N) {
          for (int m = 0; m < p.M; ++m) {
            float a = 0.f;
            for (int zz = 0; zz < Z; ++zz) a += xred[(zz * MMAX + m) * BN8 + t];       // rank order: deterministic
            T val = DT<T>::from_f(a * pre_s);
            if (p.bias != nullptr) val = DT<T>::from_f(DT<T>::to_f(val) + DT<T>::to_f(pre_b));
            static_cast<T*>(p.C)[m * p.ldc + n] = val;
          }
        }
      }
    }
  }
}

template <typename T>
int launch(const GemmArgs& a) {
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  static const int s_env = env_int("CGQ_W8_MX_STAGES", 0, 0, 16);
  const int SPT = (a.K + KSTAGE - 1) / KSTAGE;
  const int tiles = (a.N + BN8 - 1) / BN8;
  const int slots = 3 * sm_count();
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2) Z *= 2;
  const int grid = tiles * Z;
  const int per_cta = (SPT + Z - 1) / Z;
  static const int s_big = env_int("CGQ_W8_MX_STAGES_BIG", 6, 2, 16), s_sml = env_int("CGQ_W8_MX_STAGES_SMALL", 6, 2, 16);
  int stages = s_env > 0 ? s_env : (grid * 3 <= slots * 2 ? s_sml : s_big);
  if (stages > per_cta) stages = per_cta < 2 ? 2 : per_cta;
  CUtensorMap tmW, tmA;
  TmapKey kw{a.Wq, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.N), static_cast<uint64_t>(a.K), KSTAGE, BN8,
             CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_128B};
  int rc = get_tmap_2d(kw, &tmW);
  if (rc != CGQ_OK) return rc;
  TmapKey ka{a.A, static_cast<uint64_t>(a.K), static_cast<uint64_t>(a.M), static_cast<uint64_t>(a.lda) * 2, 64, 8,
             a.dtype == CGQ_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
             CU_TENSOR_MAP_SWIZZLE_128B};
  rc = get_tmap_2d(ka, &tmA);
  if (rc != CGQ_OK) return rc;
  PX prm;
  prm.scale = a.scale;
  prm.bias = a.bias;
  prm.C = a.C;
  prm.ldc = a.ldc;
  prm.M = a.M;
  prm.N = a.N;
  prm.K = a.K;
  prm.SPT = SPT;
  prm.Z = Z;
  prm.S = stages;
  const size_t smem = 1024 + static_cast<size_t>(stages) * (W_BYTES + AX_BYTES) +
                      static_cast<size_t>(Z) * MMAX * BN8 * 4 + 16 * stages + 48;
  auto kern = w8_gemv_mx_kernel<T>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && smem > configured[dev]) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
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
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmW, tmA, prm));
  return CGQ_OK;
}

// the TMA box of the activations needs 16-byte aligned rows and a K the 64-element boxes tile exactly enough for
// the tensor map (K % 8); anything else stays on the stream-K kernel
bool supported(const GemmArgs& a) { return a.M >= 2 && a.M <= MMAX && a.K % 8 == 0 && (a.lda * 2) % 16 == 0; }

}  // namespace mx

}  // namespace

bool w8_gemv_supported(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const int tiles = (a.N + BN8 - 1) / BN8;
  return a.M >= 1 && a.M <= MMAX && a.K % 16 == 0 && al16(a.Wq) && al16(a.A) && a.lda % 8 == 0 &&
         tiles <= kMaxTiles;
}

int launch_w8_gemv(const GemmArgs& a) {
  static const bool legacy = env_int("CGQ_W8_STREAMK_M1", 0, 0, 1) != 0;    // the round-1 stream-K path for one token
  if (a.M == 1 && !legacy)
    return a.dtype == CGQ_DTYPE_F16 ? m1::launch<__half>(a, nullptr) : m1::launch<__nv_bfloat16>(a, nullptr);
  static const bool legacy_mx = env_int("CGQ_W8_STREAMK_MX", 0, 0, 1) != 0; // ... and for 2 .. 8 tokens
  if (!legacy_mx && mx::supported(a))
    return a.dtype == CGQ_DTYPE_F16 ? mx::launch<__half>(a) : mx::launch<__nv_bfloat16>(a);
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a) : launch_t<__nv_bfloat16>(a);
}

// M == 1 with a fused prologue (RMSNorm / SiLU-gate on the activation) and residual epilogue.
int launch_w8_gemv_fused(const GemmArgs& a, const GemvFused& fu) {
  return a.dtype == CGQ_DTYPE_F16 ? m1::launch<__half>(a, &fu) : m1::launch<__nv_bfloat16>(a, &fu);
}

}  // namespace cgq
