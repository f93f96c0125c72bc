// Persistent decode "program": a whole chain of batch-1 int4g32 linears (with their fused RMSNorm /
// SiLU-gate prologues and residual epilogues) in ONE launch.
//
// Why: launched one kernel per linear (gemv_w4.cu), a decode token spends ~3 us per launch in
// dependency bubbles -- the next grid can only prefetch weights into shared memory the current grid
// has not taken, then waits for the producer grid to drain, then hands the activation over.  Here the
// CTAs are persistent workers: the producer lane of every worker walks the WHOLE program and keeps
// its TMA ring full with the weights of whatever comes next, across linear boundaries, while the
// consumer warps sit in the grid barrier that separates two dependent linears.  HBM keeps streaming
// through every hand-over.
//
//   * grid = 8 x (max co-resident clusters); cluster = 8 workers; worker = 1 producer warp + 4
//     consumer warps, 4-stage ring of [64 x 128 B] weight tiles + [4 x 128] scale tiles (gemv_w4.cu's
//     stage), <= 4 workers per SM;
//   * linear p is cut like gemv_w4.cu cuts it: (128-column tile) x (Z_p k-bands), the Z_p bands of a
//     tile on Z_p consecutive ranks of ONE cluster; their sums are pushed into the first rank's shared
//     memory (DSMEM) and announced with a remote mbarrier arrive (barrier.cluster cannot be used: the
//     producer warps run ahead and never join);
//   * between dependent linears the consumer warps meet in a grid barrier (one atomic counter in
//     global memory, release / acquire at gpu scope); the producers never do;
//   * the arithmetic of a stage is gemv_w4.cu's M == 1 path, bit for bit.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "w4_dev.cuh"

namespace cgq {
namespace {

using namespace w4;

constexpr int kCluster = 8;
constexpr int kStages = 4;
constexpr int kThreads = (CW + 1) * 32;
constexpr int kMaxBandUnits = 32;                       // k-stages of one band (activation band buffer)
constexpr int kRedBytes = CW * BN * 4;
constexpr int kXredBytes = kCluster * BN * 4;
constexpr int kBandBytes = kMaxBandUnits * KSTAGE * 2;
constexpr int kSmemBytes = 1024 + kStages * (W_BYTES + S_BYTES) + kRedBytes + kXredBytes + 128 + kBandBytes;

struct alignas(64) OpDev {
  CUtensorMap tmW, tmS;
  const void* A;        // activation row ([K], or [2K] for PRO_SILU_GATE)
  const void* bias;     // [N] or null
  const void* norm_w;   // [K] (PRO_RMSNORM)
  const void* resid;    // [N] or null
  void* C;              // [N]
  int N, K, SPT, Z, tiles, prologue;
  float eps;
  int pad;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// remote arrive on an mbarrier in another CTA of the cluster (cluster-scope release)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity)
        : "memory");
  } while (ok == 0);
}

// Consumer warps of every worker meet here between two dependent linears.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, int* failed) {
  ptx::named_bar_sync(1, CW * 32);            // this worker's global stores are issued
  if (threadIdx.x == 0) {
    red_release_gpu(ctr, 1u);
    unsigned spins = 0;
    while (ld_acquire_gpu(ctr) < target) {
      __nanosleep(64);
      // a worker is missing (seconds have passed): give up loudly instead of hanging the device; the flag is
      // sticky, later barriers fall through at once and cgq_program_status reports the failure
      if (++spins > (1u << 21) || *reinterpret_cast<volatile int*>(failed) != 0) {
        *failed = 1;
        break;
      }
    }
  }
  ptx::named_bar_sync(1, CW * 32);
}

template <typename T, bool kTrick>
__global__ void __launch_bounds__(kThreads, 4)
    w4_program_kernel(const OpDev* __restrict__ ops, int n_ops, unsigned* __restrict__ ctr, int* __restrict__ failed,
                      unsigned long long* __restrict__ trace) {
  // optional timeline (cgq_debug_trace): 4 stamps per (op, worker): barrier passed, band staged, loop end, stored
  auto stamp = [&](int op, int slot) {
    if (trace != nullptr && threadIdx.x == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      trace[(static_cast<size_t>(op) * gridDim.x + blockIdx.x) * 4 + slot] = t;
    }
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  constexpr int S = kStages;
  const uint32_t Wsm = base;
  const uint32_t Ssm = Wsm + S * W_BYTES;
  const uint32_t off_red = S * (W_BYTES + S_BYTES);
  float* red = reinterpret_cast<float*>(gen + off_red);
  float* xred = reinterpret_cast<float*>(gen + off_red + kRedBytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + kRedBytes + kXredBytes);
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;                      // band sums of the peers have landed (leader ranks)
  const uint32_t Aband = base + off_red + kRedBytes + kXredBytes + 128;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NC = gridDim.x / kCluster;
  const int cid = blockIdx.x / kCluster, rank = blockIdx.x - cid * kCluster;

  if (threadIdx.x == CW * 32) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    ptx::mbar_init(xbar, kCluster - 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  // every CTA of the cluster must have initialised its barriers before a peer arrives on them
  ptx::cluster_arrive_release();
  ptx::cluster_wait_acquire();

  if (warp == CW) {
    // =========================== producer: one lane walks the whole program ===========================
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      unsigned issued = 0;
      for (int op = 0; op < n_ops; ++op) {
        const OpDev* o = ops + op;
        const int Z = o->Z, SPT = o->SPT, tiles = o->tiles;
        const int gpc = kCluster / Z, groups = NC * gpc;
        const int gi = cid * gpc + rank / Z, z = rank % Z;
        const int u0 = SPT * z / Z, u1 = SPT * (z + 1) / Z;
        for (int tile = gi; tile < tiles; tile += groups) {
          for (int ks = u0; ks < u1; ++ks) {
            const int slot = issued % S;
            if (issued >= S) ptx::mbar_wait(&empty[slot], ((issued / S) - 1) & 1);
            ptx::mbar_expect_tx(&full[slot], W_BYTES + S_BYTES);
            ptx::tma_load_2d(gen + slot * W_BYTES, &o->tmW, tile * BN, ks * ROWS, &full[slot], pol);
            ptx::tma_load_2d(gen + S * W_BYTES + slot * S_BYTES, &o->tmS, tile * BN, ks * CW, &full[slot], pol);
            ++issued;
          }
        }
      }
    }
    return;
  }

  // =========================== consumers ===========================
  const int tid = threadIdx.x;
  if (trace != nullptr && tid == 0) {   // op 0, slot 3 is never stamped by idle workers: record the SM id there
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    trace[static_cast<size_t>(n_ops) * gridDim.x * 4 + blockIdx.x] = smid + 1;
  }
  const int g = lane >> 2, tig = lane & 3;
  const bool has_tok = g == 0;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  unsigned consumed = 0, xphase = 0;
  for (int op = 0; op < n_ops; ++op) {
    const OpDev* o = ops + op;
    const int Z = o->Z, SPT = o->SPT, tiles = o->tiles, K = o->K, N = o->N, pro = o->prologue;
    const int gpc = kCluster / Z, groups = NC * gpc;
    const int gi = cid * gpc + rank / Z, z = rank % Z;
    const int u0 = SPT * z / Z, u1 = SPT * (z + 1) / Z;
    const int n_units = u1 - u0;
    const T* A = static_cast<const T*>(o->A);
    if (op > 0) grid_barrier(ctr, static_cast<unsigned>(op) * gridDim.x, failed);
    stamp(op, 0);

    if (gi < tiles) {
      // ---- stage this worker's k-band of the activation row (gemv_w4.cu's M == 1 prologue)
      const int nchunk = K >> 3;
      const int c_lo = u0 * (KSTAGE / 8), c_hi = u1 * (KSTAGE / 8);
      if (pro == PRO_RMSNORM) {
        const T* nw = static_cast<const T*>(o->norm_w);
        float ss = 0.f;
        for (int c = tid; c < nchunk; c += CW * 32) ss += sumsq8<T>(ldcg128(A + c * 8));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if (lane == 0) red[warp] = ss;
        ptx::named_bar_sync(2, CW * 32);
        float tot_ss = 0.f;
#pragma unroll
        for (int w = 0; w < CW; ++w) tot_ss += red[w];
        const float rstd = rsqrtf(tot_ss / static_cast<float>(K) + o->eps);
        for (int c = c_lo + tid; c < c_hi; c += CW * 32)
          ptx::sts128(Aband + (c - c_lo) * 16,
                      c < nchunk ? rmsnorm8<T>(ldcg128(A + c * 8), ldnc128(nw + c * 8), rstd) : zero);
      } else {
        for (int c = c_lo + tid; c < c_hi; c += CW * 32) {
          uint4 v = zero;
          if (c < nchunk) {
            v = ldcg128(A + c * 8);
            if (pro == PRO_SILU_GATE) v = silu_gate8<T>(v, ldcg128(A + K + c * 8));
          }
          ptx::sts128(Aband + (c - c_lo) * 16, v);
        }
      }
      ptx::named_bar_sync(2, CW * 32);
    }
    stamp(op, 1);

    for (int tile = gi; tile < tiles; tile += groups) {
      float tot[8][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) tot[j][0] = tot[j][1] = 0.f;
      for (int it = 0; it < n_units; ++it) {
        const int slot = consumed % S;
        ptx::mbar_wait(&full[slot], (consumed / S) & 1);
        const uint32_t wrow = Wsm + slot * W_BYTES + (16 * warp) * BN;
        const uint32_t srow = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 32;
        float grp[8][4];
        float ag[4];
        uint32_t w_dep = 0;
        const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int r = 8 * b + 2 * tig;
          const uint4 q = ptx::lds128(wrow + r * BN + ((g ^ (2 * tig)) << 4));
          const uint4 pp = ptx::lds128(wrow + (r + 1) * BN + ((g ^ (2 * tig + 1)) << 4));
          w_dep = q.x | pp.x;
          uint2 av = make_uint2(0u, 0u);
          if (has_tok) av = ptx::lds64(Aband + (it * KSTAGE + 32 * warp + 4 * tig) * 2 + 32 * b);
          uint32_t b0 = __byte_perm(av.x, av.y, 0x5410);
          uint32_t b1 = __byte_perm(av.x, av.y, 0x7632);
          if (kTrick) b1 = h2_mul(b1, 0x2C002C00u);
          if (kTrick) {
            const uint32_t ones[4] = {0x3C003C00u, 0x3C003C00u, 0x4C004C00u, 0x4C004C00u};
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
            const uint32_t a[4] = {Nib<T, kTrick>::lo(v0), Nib<T, kTrick>::lo(v1), Nib<T, kTrick>::hi(v0),
                                   Nib<T, kTrick>::hi(v1)};
            if (b == 0)
              ptx::mma_16816(grp[j], a, b0, b1, zero4, T());
            else
              ptx::mma_16816(grp[j], a, b0, b1, grp[j], T());
          }
        }
        const uint4 sv0 = ptx::lds128(srow), sv1 = ptx::lds128(srow + 16);
        __syncwarp();
        if (lane == 0)
          ptx::mbar_arrive_after_loads(&empty[slot], sv0.x | sv1.x | w_dep, static_cast<uint32_t>(K) >> 31);
        ++consumed;
        const uint32_t sw[8] = {sv0.x, sv0.y, sv0.z, sv0.w, sv1.x, sv1.y, sv1.z, sv1.w};
        const float c0 = kTrick ? -8.f * ag[0] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          union {
            uint32_t u;
            T h[2];
          } cv;
          cv.u = sw[j];
          const float t0 = kTrick ? fmaf(grp[j][0], 16777216.f, c0) : grp[j][0];
          const float t2 = kTrick ? fmaf(grp[j][2], 16777216.f, c0) : grp[j][2];
          tot[j][0] = fmaf(DT<T>::to_f(cv.h[0]), t0, tot[j][0]);
          tot[j][1] = fmaf(DT<T>::to_f(cv.h[1]), t2, tot[j][1]);
        }
      }
      stamp(op, 2);
      // ---- band sum of this worker: cross-warp reduction, then the tile's bands on the leader rank
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
          if (tig == 0) red[warp * BN + 16 * g + 2 * j + i] = tot[j][i];
      ptx::named_bar_sync(2, CW * 32);
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < CW; ++w) v += red[w * BN + tid];
      ptx::named_bar_sync(2, CW * 32);            // `red` may be rewritten (next tile / next prologue)
      bool store = Z == 1;
      if (Z > 1) {
        const uint32_t leader = static_cast<uint32_t>(rank - z);
        if (z != 0) {
          ptx::st_cluster_f32(ptx::mapa_rank(ptx::smem_u32(xred + z * BN + tid), leader), v);
          ptx::named_bar_sync(2, CW * 32);        // all 128 pushes of this worker are issued
          if (tid == 0) mbar_arrive_remote(ptx::mapa_rank(ptx::smem_u32(xbar), leader));
        } else {
          if (tid < kCluster - Z) ptx::mbar_arrive(xbar);   // the barrier always counts 7 arrivals
          mbar_wait_cluster(xbar, xphase & 1);
          ++xphase;
          for (int zz = 1; zz < Z; ++zz) v += xred[zz * BN + tid];   // rank order: deterministic
          store = true;
        }
      }
      if (store) {
        const int n = tile * BN + tid;
        if (n < N)
          static_cast<T*>(o->C)[n] = add_resid<T>(epilogue<T>(v, static_cast<const T*>(o->bias), n),
                                                  static_cast<const T*>(o->resid), n);
      }
      stamp(op, 3);
    }
  }
}

// ------------------------------------------------------------------ host: program objects
struct Program {
  OpDev* d_ops = nullptr;
  unsigned* d_ctr = nullptr;   // [0] grid-barrier counter, [1] failure flag
  int n_ops = 0;
  int dtype = 0;
  int grid = 0;
  int device = 0;
};

std::mutex g_mu;
std::unordered_map<uint64_t, Program> g_programs;
uint64_t g_next_handle = 1;

template <typename T, bool kTrick>
int max_clusters(int* out) {
  auto kern = w4_program_kernel<T, kTrick>;
  CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                    cudaSharedmemCarveoutMaxShared));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kCluster * 148);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  CGQ_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
  *out = n;
  return CGQ_OK;
}

template <typename T, bool kTrick>
int launch_program(const Program& pr, cudaStream_t st) {
  auto kern = w4_program_kernel<T, kTrick>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pr.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CGQ_CUDA_TRY(cudaMemsetAsync(pr.d_ctr, 0, 2 * sizeof(unsigned), st));
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, static_cast<const OpDev*>(pr.d_ops), pr.n_ops, pr.d_ctr,
                                  reinterpret_cast<int*>(pr.d_ctr + 1),
                                  static_cast<unsigned long long*>(take_trace_buffer())));
  return CGQ_OK;
}

}  // namespace
}  // namespace cgq

using namespace cgq;

extern "C" int cgq_program_create(const cgq_linear_op* ops, int n_ops, int dtype, uint64_t* handle) {
  const char* fn = "cgq_program_create";
  if (ops == nullptr || n_ops <= 0 || handle == nullptr) {
    set_error("%s: null / empty program", fn);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16 && dtype != CGQ_DTYPE_BF16) {
    set_error("%s: unsupported dtype code %d", fn, dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  int nc = 0;
  int rc = dtype == CGQ_DTYPE_F16 ? max_clusters<__half, true>(&nc) : max_clusters<__nv_bfloat16, false>(&nc);
  if (rc != CGQ_OK) return rc;
  if (nc < 1) {
    set_error("%s: no cluster of %d workers fits on this device", fn, kCluster);
    return CGQ_ERR_UNSUPPORTED;
  }
  std::vector<OpDev> host(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    const cgq_linear_op& s = ops[i];
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (s.N <= 0 || s.K <= 0 || s.K % 32 != 0 || s.N % 16 != 0 || s.Wq == nullptr || s.scale == nullptr ||
        s.A == nullptr || s.C == nullptr || !al16(s.Wq) || !al16(s.scale) || !al16(s.A) ||
        (s.prologue == CGQ_PRO_RMSNORM && (s.norm_w == nullptr || !al16(s.norm_w))) ||
        (s.prologue != CGQ_PRO_NONE && s.prologue != CGQ_PRO_RMSNORM && s.prologue != CGQ_PRO_SILU_GATE)) {
      set_error("%s: op %d: bad shape / pointer (N=%d K=%d prologue=%d)", fn, i, s.N, s.K, s.prologue);
      return CGQ_ERR_BAD_SHAPE;
    }
    OpDev& d = host[i];
    const int G = s.K / 32;
    d.SPT = (G + CW - 1) / CW;
    d.tiles = (s.N + BN - 1) / BN;
    // k-bands per tile: as many as keep every tile on one pass of the workers, bands of >= 2 stages that fit
    // the activation band buffer; more passes are only allowed without the cross-worker reduction (Z == 1)
    int Z = kCluster;
    while (Z > 1 && (d.tiles > nc * (kCluster / Z) || d.SPT < 2 * Z)) Z /= 2;
    if ((d.SPT + Z - 1) / Z > kMaxBandUnits) {
      set_error("%s: op %d: K=%d needs a k-band of more than %d stages per worker", fn, i, s.K, kMaxBandUnits);
      return CGQ_ERR_BAD_SHAPE;
    }
    d.Z = Z;
    d.A = s.A;
    d.bias = s.bias;
    d.norm_w = s.norm_w;
    d.resid = s.resid;
    d.C = s.C;
    d.N = s.N;
    d.K = s.K;
    d.prologue = s.prologue;
    d.eps = s.eps;
    d.pad = 0;
    TmapKey kw{s.Wq, static_cast<uint64_t>(s.N), static_cast<uint64_t>(s.K / 2), static_cast<uint64_t>(s.N), BN,
               ROWS, CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_128B};
    rc = get_tmap_2d(kw, &d.tmW);
    if (rc != CGQ_OK) return rc;
    TmapKey ks{s.scale, static_cast<uint64_t>(s.N), static_cast<uint64_t>(G), static_cast<uint64_t>(s.N) * 2, BN, CW,
               dtype == CGQ_DTYPE_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
               CU_TENSOR_MAP_SWIZZLE_NONE};
    rc = get_tmap_2d(ks, &d.tmS);
    if (rc != CGQ_OK) return rc;
  }
  Program pr;
  pr.n_ops = n_ops;
  pr.dtype = dtype;
  pr.grid = nc * kCluster;
  CGQ_CUDA_TRY(cudaGetDevice(&pr.device));
  CGQ_CUDA_TRY(cudaMalloc(&pr.d_ops, sizeof(OpDev) * n_ops));
  CGQ_CUDA_TRY(cudaMalloc(&pr.d_ctr, 2 * sizeof(unsigned)));
  CGQ_CUDA_TRY(cudaMemcpy(pr.d_ops, host.data(), sizeof(OpDev) * n_ops, cudaMemcpyHostToDevice));
  std::lock_guard<std::mutex> lk(g_mu);
  *handle = g_next_handle++;
  g_programs[*handle] = pr;
  return CGQ_OK;
}

extern "C" int cgq_program_run(uint64_t handle, void* stream) {
  Program pr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_programs.find(handle);
    if (it == g_programs.end()) {
      set_error("cgq_program_run: unknown program handle");
      return CGQ_ERR_BAD_SHAPE;
    }
    pr = it->second;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return pr.dtype == CGQ_DTYPE_F16 ? launch_program<__half, true>(pr, st)
                                   : launch_program<__nv_bfloat16, false>(pr, st);
}

extern "C" int cgq_program_status(uint64_t handle, int* workers, int* failed) {
  Program pr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_programs.find(handle);
    if (it == g_programs.end()) {
      set_error("cgq_program_status: unknown program handle");
      return CGQ_ERR_BAD_SHAPE;
    }
    pr = it->second;
  }
  unsigned h[2] = {0, 0};
  CGQ_CUDA_TRY(cudaDeviceSynchronize());   // (a plain cudaMemcpy would not wait for non-blocking streams)
  CGQ_CUDA_TRY(cudaMemcpy(h, pr.d_ctr, sizeof(h), cudaMemcpyDeviceToHost));
  if (workers != nullptr) *workers = pr.grid;
  if (failed != nullptr) *failed = static_cast<int>(h[1]);
  return CGQ_OK;
}

extern "C" int cgq_program_destroy(uint64_t handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_programs.find(handle);
  if (it == g_programs.end()) return CGQ_OK;
  cudaFree(it->second.d_ops);
  cudaFree(it->second.d_ctr);
  g_programs.erase(it);
  return CGQ_OK;
}
