// mkprobe — design probe for the one-launch decode step ("megakernel"): the SKELETON of a persistent kernel
// (one fat CTA per SM: 1 TMA producer warp + 4 consumer teams of 4 warps, a deep shared-memory ring, a flat grid
// barrier between dependent linears) streaming the 113 int4g32 linears of a ChatGLM2-6B decode token, with the
// consumer arithmetic replaced by a calibrated spin.  It answers, before the real kernel is written:
//   * how fast do [rows x BW]-byte TMA boxes stream for BW = 32 / 64 / 128 (narrow column slices need no
//     cross-CTA reduction, wide ones stream in fewer requests);
//   * what does a 148-CTA grid barrier cost while the producers keep HBM busy;
//   * how much consumer headroom (cycles per 9 KB stage) the chain tolerates before it stops being HBM-bound.
//
//   mkprobe BW Z STAGES SPIN BARRIER [reps]
//     BW      bytes (= columns) per slice: 32, 64 or 128; a ring stage is always 8 KB of packed weights + 1 KB scales
//     Z       k-parts per slice (work item = slice x k-part, dealt k-part-major round-robin to the CTAs)
//     STAGES  ring depth (<= 22)
//     SPIN    clocks a consumer team spends per stage (0 = no compute); 4 teams work on 4 stages concurrently
//     BARRIER 1 = grid barrier between consecutive linears, 0 = none
// Development tool (standalone, no torch); results are recorded in DESIGN.md §5.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));          \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

constexpr int kTeams = 4, kTeamWarps = 4;
constexpr int kConsumers = kTeams * kTeamWarps * 32;
constexpr int kThreads = kConsumers + 32;
constexpr int W_BYTES = 8192, S_BYTES = 1024;

struct alignas(64) Op {
  CUtensorMap tmW, tmS;
  int slices, spk;   // column slices, k-stages per slice (whole K)
  int pad[14];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::
          "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct Args {
  const Op* ops;
  int n_ops, BW, Z, S, spin, barrier;
  unsigned* ctr;
  unsigned long long* trace;   // [n_ops + 1] globaltimer of CTA 0 leaving each barrier
  unsigned* sink;
};

__global__ void __launch_bounds__(kThreads, 1) probe_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int S = a.S;
  uint8_t* Wsm = smem;
  uint8_t* Ssm = smem + S * W_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * (W_BYTES + S_BYTES));
  uint64_t* empty = full + S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = W_BYTES / a.BW;          // packed rows per stage
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTeamWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int W = gridDim.x, w = blockIdx.x;

  if (warp == kConsumers / 32) {
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      unsigned issued = 0;
      for (int op = 0; op < a.n_ops; ++op) {
        const Op* o = a.ops + op;
        const int items = o->slices * a.Z;
        for (int it = w; it < items; it += W) {
          const int z = it / o->slices, sl = it - z * o->slices;
          const int u0 = o->spk * z / a.Z, u1 = o->spk * (z + 1) / a.Z;
          for (int u = u0; u < u1; ++u) {
            const int slot = issued % S;
            if (issued >= (unsigned)S) mbar_wait(&empty[slot], ((issued / S) - 1) & 1);
            mbar_expect_tx(&full[slot], W_BYTES + S_BYTES);
            tma_load_2d(Wsm + slot * W_BYTES, &o->tmW, sl * a.BW, u * rows, &full[slot], pol);
            tma_load_2d(Ssm + slot * S_BYTES, &o->tmS, sl * a.BW, u * (rows / 16), &full[slot], pol);
            ++issued;
          }
        }
      }
    }
    return;
  }
  // consumers: team t takes the stages == t (mod 4) of this CTA's stream
  const int team = warp / kTeamWarps;
  unsigned seen = 0;   // stages of this CTA's stream so far (all teams count alike)
  unsigned acc = 0;
  if (threadIdx.x == 0 && a.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (w == 0) a.trace[0] = t;
  }
  for (int op = 0; op < a.n_ops; ++op) {
    const Op* o = a.ops + op;
    const int items = o->slices * a.Z;
    for (int it = w; it < items; it += W) {
      const int z = it / o->slices;
      const int n_u = o->spk * (z + 1) / a.Z - o->spk * z / a.Z;
      for (int u = 0; u < n_u; ++u, ++seen) {
        if ((int)(seen % kTeams) != team) continue;
        const int slot = seen % S;
        mbar_wait(&full[slot], (seen / S) & 1);
        acc += Wsm[slot * W_BYTES + threadIdx.x * 4];
        if (a.spin > 0) {
          const long long t0 = clock64();
          while (clock64() - t0 < a.spin) {
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[slot]);
      }
    }
    if (a.barrier) {
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(a.ctr), "r"(1u) : "memory");
        const unsigned target = static_cast<unsigned>(op + 1) * W;
        while (ld_acquire(a.ctr) < target) {
        }
        if (a.trace && w == 0) {
          unsigned long long t;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
          a.trace[op + 1] = t;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    }
  }
  if (acc == 0xdeadbeef) *a.sink = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int BW = argc > 1 ? atoi(argv[1]) : 32;
  const int Z = argc > 2 ? atoi(argv[2]) : 1;
  const int S = argc > 3 ? atoi(argv[3]) : 20;
  const int spin = argc > 4 ? atoi(argv[4]) : 0;
  const int barrier = argc > 5 ? atoi(argv[5]) : 1;
  const int reps = argc > 6 ? atoi(argv[6]) : 10;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  const int H = 4096, INNER = 13696, VOCAB = 65024, QKV = 4608, LAYERS = 28;
  struct Shape { int K, N; };
  std::vector<Shape> shapes;
  for (int l = 0; l < LAYERS; ++l) {
    shapes.push_back({H, QKV});
    shapes.push_back({H, H});
    shapes.push_back({H, 2 * INNER});
    shapes.push_back({INNER, H});
  }
  shapes.push_back({H, VOCAB});
  const int rows = W_BYTES / BW;
  std::vector<Op> ops(shapes.size());
  double bytes = 0;
  for (size_t i = 0; i < shapes.size(); ++i) {
    const int K = shapes[i].K, N = shapes[i].N;
    uint8_t* w;
    __half* s;
    CK(cudaMalloc(&w, (size_t)K / 2 * N));
    CK(cudaMalloc(&s, (size_t)(K / 32) * N * 2));
    CK(cudaMemset(w, 0x55, (size_t)K / 2 * N));
    CK(cudaMemset(s, 0, (size_t)(K / 32) * N * 2));
    bytes += (double)K / 2 * N + (double)(K / 32) * N * 2;
    cuuint64_t dW[2] = {(cuuint64_t)N, (cuuint64_t)K / 2}, sW[1] = {(cuuint64_t)N};
    cuuint32_t bW[2] = {(cuuint32_t)BW, (cuuint32_t)rows}, es[2] = {1, 1};
    CUresult r = enc(&ops[i].tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, w, dW, sW, bW, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dS[2] = {(cuuint64_t)N, (cuuint64_t)K / 32}, sS[1] = {(cuuint64_t)N * 2};
    cuuint32_t bS[2] = {(cuuint32_t)BW, (cuuint32_t)(rows / 16)};
    CUresult r2 = enc(&ops[i].tmS, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, s, dS, sS, bS, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      fprintf(stderr, "tensor map encode failed %d %d (K=%d N=%d BW=%d)\n", (int)r, (int)r2, K, N, BW);
      return 1;
    }
    ops[i].slices = N / BW;
    ops[i].spk = (K / 2 + rows - 1) / rows;
  }
  Op* d_ops;
  CK(cudaMalloc(&d_ops, sizeof(Op) * ops.size()));
  CK(cudaMemcpy(d_ops, ops.data(), sizeof(Op) * ops.size(), cudaMemcpyHostToDevice));
  unsigned* ctr;
  CK(cudaMalloc(&ctr, 8));
  unsigned long long* trace;
  CK(cudaMalloc(&trace, 8 * (ops.size() + 1)));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t smem = (size_t)S * (W_BYTES + S_BYTES) + 16 * S + 64;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe_kernel, kThreads, smem));
  if (occ < 1) {
    fprintf(stderr, "kernel does not fit (smem %zu)\n", smem);
    return 1;
  }
  Args a{d_ops, (int)ops.size(), BW, Z, S, spin, barrier, ctr, trace, ctr + 1};
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  auto run = [&]() {
    CK(cudaMemsetAsync(ctr, 0, 8, st));
    probe_kernel<<<sms, kThreads, smem, st>>>(a);
  };
  for (int i = 0; i < 2; ++i) run();
  CK(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; ++i) run();
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= reps;
  printf("mkprobe BW=%d Z=%d stages=%d spin=%d barrier=%d: %.1f us/token  %.2f TB/s  (%d CTAs, %.3f GB)", BW, Z, S, spin,
         barrier, ms * 1e3, bytes / (ms * 1e-3) / 1e12, sms, bytes / 1e9);
  if (barrier) {
    std::vector<unsigned long long> h(ops.size() + 1);
    CK(cudaMemcpy(h.data(), trace, 8 * h.size(), cudaMemcpyDeviceToHost));
    // second block: per-op time between barrier exits
    printf("  | layer 1 ops (us):");
    for (int i = 4; i < 8; ++i) printf(" %.2f", (h[i + 1] - h[i]) / 1e3);
    printf(" lm_head %.2f", (h[113] - h[112]) / 1e3);
  }
  printf("\n");
  return 0;
}
