// Glue kernels of the fused batch-1 decode step: everything of
// ChatGLM2Model.forward (chatglm_q/model.py:329-392) that sits BETWEEN the dequant-matmuls when one
// new token is decoded against a KV cache.  RMSNorm, SiLU*gate and the residual adds are fused into
// the M == 1 int4 kernel (gemv_w4.cu: cgq_w4a16_gemv_fused); what is left is
//   * decode_begin  : QEmbedding row gather of the new token (int4/qlinear.py:122-130) + position
//                     bookkeeping on the device (so the step is a static CUDA graph);
//   * decode_attn   : RoPE of q/k (model.py:47-59,148-149), KV-cache append (:151-155) and the
//                     multi-query attention of ONE query row (:157-174), one CTA per head.
// Every kernel takes part in the programmatic-dependent-launch chain of the step: it releases its
// dependents at once (the next dequant-matmul streams its weights while this kernel runs) and waits
// for its producers before touching activations.
//
// Numerics follow the reference's rounding points for T = fp16 / bf16:
//   rope      : complex product in fp32, rounded once to T           (torch complex-half mul = fp32 opmath)
//   q scaling : round_T(q * (1/sqrt(d)))                             (torch divides by a scalar via the reciprocal)
//   scores    : round_T(fp32 dot)                                    (matmul, :164)
//   softmax   : fp32 over the scores (online, per warp); the reference additionally rounds the
//               probabilities to T (:168) -- ours stay fp32, the more accurate side of the 1e-2 bar
//   output    : round_T(fp32 sum of p * v_T / fp32 sum of p)         (:171)
#include "common.cuh"
#include "ptx.cuh"

namespace cgq {
namespace {

// ------------------------------------------------------------------ decode_begin
// state[0] = tokens already in the KV cache BEFORE this step (host writes it after prefill, the kernel
//            increments it), state[1] = the value the kernels of THIS step use.
template <typename T>
__global__ void __launch_bounds__(256)
    decode_begin_w4_kernel(const int64_t* __restrict__ ids, const uint8_t* __restrict__ Wq,
                           const T* __restrict__ scale, T* __restrict__ x, int V, int D, int group,
                           int* __restrict__ state) {
  // First kernel of the step's graph: everything before it has completed.  The position is published
  // (written + fenced) BEFORE the dependents are released, so that the attention kernels further down the
  // programmatic-launch chain may read state[1] ahead of their own griddepcontrol.wait.
  ptx::pdl_wait_prior_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int cur = state[0];
    state[1] = cur;
    state[0] = cur + 1;
    state[2] = state[2] + 1;     // token counter (epoch source of the tensor-parallel exchanges)
    __threadfence();
  }
  __syncthreads();
  ptx::pdl_launch_dependents();
  int64_t t = ids[0];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);   // an id outside the vocabulary must never read out of bounds
  const uint8_t* wrow = Wq + (t >> 1) * D;
  const T* srow = scale + (t / group) * D;
  const int shift = static_cast<int>(t & 1) * 4;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) x[d] = dequant4<T>((wrow[d] >> shift) & 0xF, srow[d]);
}

// int8 QEmbedding row (int8/qlinear.py:118-120, scales are per embedding column): x[d] = round_T(q[t, d] * scale[d])
template <typename T>
__global__ void __launch_bounds__(256)
    decode_begin_w8_kernel(const int64_t* __restrict__ ids, const int8_t* __restrict__ Wq, const T* __restrict__ scale,
                           T* __restrict__ x, int V, int D, int* __restrict__ state) {
  ptx::pdl_wait_prior_grid();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int cur = state[0];
    state[1] = cur;
    state[0] = cur + 1;
    state[2] = state[2] + 1;
    __threadfence();
  }
  __syncthreads();
  ptx::pdl_launch_dependents();
  int64_t t = ids[0];
  t = t < 0 ? 0 : (t >= V ? V - 1 : t);
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < D) x[d] = dequant8<T>(Wq[t * D + d], scale[d]);
}

// ------------------------------------------------------------------ decode_attn
struct AttnParams {
  const void* qkv;     // [n_head*DH | n_groups*DH | n_groups*DH]
  const void* freqs;   // [max_pos, DH] (cos, sin) pairs, second half (1, 0)  (model.py:33-43)
  void* kcache;        // [max_len, n_groups, DH]
  void* vcache;
  void* out;           // [n_head*DH]
  const int* state;
  int n_head, n_groups, max_len;
  int splits;          // CTAs per head (cluster size): the context is dealt to them in blocks of 16 rows
  unsigned long long* trace;  // optional timeline (cgq_debug_trace), 8 words per CTA
  // K / V caches of the NEXT attention launch of the step (cgq_attention_next_kv) or null: their live rows are
  // requested into L2 one layer ahead.  A cache row is touched once per token and 3.4 GB of weights stream through
  // L2 in between, so every step finds it in HBM -- and the loads of a launch queue behind the ~20 MB of weight
  // requests the neighbouring linears keep in flight (2.9 us before the first score at a context of 96 rows).
  const void* next_k;
  const void* next_v;
};

__device__ __forceinline__ void attn_stamp(const AttnParams& p, int slot) {
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + slot] = t;
  }
}

constexpr int kAttnThreads = 512;
constexpr int kAttnWarps = kAttnThreads / 32;
constexpr int kPrefetchIters = 4;   // row batches (32 / LPR rows per warp each) a lane has in flight before the dependency wait

// activations written by the previous kernel of the chain are read past L1 (a stale line from an earlier
// layer's use of the same buffer must never be hit)
template <typename T>
__device__ __forceinline__ T ldcg_16(const T* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return *reinterpret_cast<T*>(&v);
}
__device__ __forceinline__ int ldcg_i32(const int* p) {
  int v;
  asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldcg_128(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// One query row, one head = one CLUSTER of S CTAs (S = p.splits in {1,2,4,8}).
// A warp works on RPW = 256 / DH cached rows at a time: lane = (row slot `sub` = lane / LPR, column chunk `c` =
// lane % LPR), a lane holds 8 consecutive elements of its row (one 16-byte load; a dot product is 8 FMAs and
// log2(LPR) shuffle steps for RPW rows, where one row per warp needed five steps per row -- the shuffle chain, not the
// cache, bounded the old kernel: 2.3 us for 96 rows).  Row batch i of warp w on cluster rank r: rows
// ((i S + r) 16 + w) RPW + sub, so every context length is balanced over warps and ranks.  Each row slot keeps an
// online softmax (running max, sum, un-normalised output) in registers; slots are merged by shuffles, warps through
// shared memory and cluster ranks through distributed shared memory, all in a fixed order (deterministic).
template <typename T, int DH>
__global__ void __launch_bounds__(kAttnThreads, 2) decode_attn_kernel(const AttnParams p) {
  constexpr int EPL = 8;               // elements of a row per lane: one 16-byte load
  constexpr int NV = 1;
  constexpr int LPR = DH / EPL;        // lanes per row
  constexpr int RPW = 32 / LPR;        // rows per warp and batch
  constexpr int kMaxSplit = 8;
  extern __shared__ float sm[];
  float* q_s = sm;                       // rotated, scaled query (T-rounded values)
  float* k_s = q_s + DH;                 // rotated new key
  float* v_s = k_s + DH;                 // new value
  float* red = v_s + DH;                 // [kAttnWarps][DH]
  float* wred = red + kAttnWarps * DH;   // [2][kAttnWarps]: running (max, sum) of every warp
  float* xstat = wred + 2 * kAttnWarps;  // [2][kMaxSplit]: (max, sum) of every cluster rank (used on rank 0)
  float* xacc = xstat + 2 * kMaxSplit;   // [kMaxSplit][DH]: partial outputs (used on rank 0)

  const int S = p.splits;
  const int h = blockIdx.x / S, rank = blockIdx.x - h * S;
  const int hpg = p.n_head / p.n_groups;
  const int g = h / hpg;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int sub = lane / LPR, c = lane % LPR;
  attn_stamp(p, 0);
  if (p.splits > 1) ptx::cluster_arrive_release();   // phase A: waited for before the first DSMEM store (racecheck)
  ptx::pdl_launch_dependents();
  // Everything that does not depend on this step's qkv is requested BEFORE the dependency wait, while the
  // qkv projection is still running: the position (published by decode_begin ahead of its own dependents),
  // the rotary row and the first cached K / V rows of this CTA (written by earlier steps).
  const int n_past = ldcg_i32(p.state + 1);
  const bool live = n_past < p.max_len;   // window exhausted: the host never launches in this state
  const T* qkv = static_cast<const T*>(p.qkv);
  const T* fr = static_cast<const T*>(p.freqs) + static_cast<size_t>(n_past + 1) * DH;  // position id = n_past + 1
  T* kc = static_cast<T*>(p.kcache);
  T* vc = static_cast<T*>(p.vcache);
  const size_t row_stride = static_cast<size_t>(p.n_groups) * DH;
  const bool writer = (h % hpg) == 0 && rank == 0;
  const T* kbase = kc + g * DH + c * EPL;
  const T* vbase = vc + g * DH + c * EPL;
  auto row_of = [&](int i) { return ((i * S + rank) * kAttnWarps + warp) * RPW + sub; };
  const int rows_per_iter = S * kAttnWarps * RPW;
  const int n_iter = live ? (n_past + rows_per_iter - 1) / rows_per_iter : 0;     // the same for every warp and rank
  uint4 kraw[kPrefetchIters][NV], vraw[kPrefetchIters][NV];
#pragma unroll
  for (int i = 0; i < kPrefetchIters; ++i) {
    const int l = row_of(i);
#pragma unroll
    for (int v = 0; v < NV; ++v) kraw[i][v] = vraw[i][v] = make_uint4(0u, 0u, 0u, 0u);   // (0 x garbage must stay 0)
    if (i < n_iter && l < n_past) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        kraw[i][v] = ldcg_128(kbase + l * row_stride + 8 * v);
        vraw[i][v] = ldcg_128(vbase + l * row_stride + 8 * v);
      }
    }
  }
  if (p.next_k != nullptr && live && t < 2) {
    // rows 0 .. n_past of the next layer's caches are one contiguous range each: CTA b takes the b-th slice
    const size_t bytes = (static_cast<size_t>(n_past) + 1) * row_stride * sizeof(T);
    const size_t per = ((bytes + gridDim.x - 1) / gridDim.x + 127) & ~static_cast<size_t>(127);
    const size_t off = per * blockIdx.x;
    if (off < bytes) {
      const size_t n = (bytes - off < per ? bytes - off : per) & ~static_cast<size_t>(15);
      const char* src = static_cast<const char*>(t == 0 ? p.next_k : p.next_v) + off;
      if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(static_cast<uint32_t>(n)) : "memory");
    }
  }
  float fc = 1.f, fs = 0.f;
  if (live && t < DH) {
    const int j = t < DH / 2 ? t : t - DH / 2;
    fc = DT<T>::to_f(fr[2 * j]);
    fs = DT<T>::to_f(fr[2 * j + 1]);
  }
  attn_stamp(p, 1);
  ptx::pdl_wait_prior_grid();
  attn_stamp(p, 2);

  if (live && t < DH) {
    // rope of pair j of q (t < DH/2) or k (t >= DH/2)
    const bool is_q = t < DH / 2;
    const int j = is_q ? t : t - DH / 2;
    const T* src = is_q ? qkv + h * DH : qkv + (p.n_head + g) * DH;
    const float a = DT<T>::to_f(ldcg_16(src + 2 * j)), b = DT<T>::to_f(ldcg_16(src + 2 * j + 1));
    const T re = DT<T>::from_f(a * fc - b * fs);
    const T im = DT<T>::from_f(a * fs + b * fc);
    if (is_q) {
      const float inv = 1.0f / sqrtf(static_cast<float>(DH));
      q_s[2 * j] = DT<T>::to_f(DT<T>::from_f(DT<T>::to_f(re) * inv));
      q_s[2 * j + 1] = DT<T>::to_f(DT<T>::from_f(DT<T>::to_f(im) * inv));
    } else {
      k_s[2 * j] = DT<T>::to_f(re);
      k_s[2 * j + 1] = DT<T>::to_f(im);
      if (writer) {
        kc[n_past * row_stride + g * DH + 2 * j] = re;
        kc[n_past * row_stride + g * DH + 2 * j + 1] = im;
      }
    }
  } else if (live && t < 2 * DH) {
    const int d = t - DH;
    const T v = ldcg_16(qkv + (p.n_head + p.n_groups + g) * DH + d);
    v_s[d] = DT<T>::to_f(v);
    if (writer) vc[n_past * row_stride + g * DH + d] = v;
  }
  __syncthreads();
  attn_stamp(p, 3);

  // ---- online softmax per row slot.  Scores are rounded to T like the reference's matmul output; the
  //      probabilities stay fp32 (the reference rounds them to T: a deviation of <= 2^-11 per weight).
  float qr[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) qr[e] = q_s[c * EPL + e];
  float m_w = -INFINITY, s_w = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  auto visit = [&](float score, const float (&vr)[EPL]) {   // one more (score, value row) for this slot
    const float m_new = fmaxf(m_w, score);
    const float rescale = __expf(m_w - m_new);              // exp(-inf) = 0 the first time
    const float pl = __expf(score - m_new);
    s_w = fmaf(s_w, rescale, pl);
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pl, vr[e], acc[e] * rescale);
    m_w = m_new;
  };
  auto group_sum = [](float d) {                            // over the LPR lanes of a row slot
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    return d;
  };
  for (int i0 = 0; i0 < n_iter; i0 += kPrefetchIters) {
    if (i0 != 0) {   // later batches: request their rows now (the first ones are already in flight)
#pragma unroll
      for (int i = 0; i < kPrefetchIters; ++i) {
        const int l = row_of(i0 + i);
#pragma unroll
        for (int v = 0; v < NV; ++v) kraw[i][v] = vraw[i][v] = make_uint4(0u, 0u, 0u, 0u);
        if (i0 + i < n_iter && l < n_past) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            kraw[i][v] = ldcg_128(kbase + l * row_stride + 8 * v);
            vraw[i][v] = ldcg_128(vbase + l * row_stride + 8 * v);
          }
        }
      }
    }
    // the kPrefetchIters dot products and their shuffle chains are independent of each other and of the softmax
    // state: no branches here, so that they overlap (a missing row scores -inf and weighs 0)
    float sc[kPrefetchIters];
    float m_new = m_w;
#pragma unroll
    for (int i = 0; i < kPrefetchIters; ++i) {
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const T* kh = reinterpret_cast<const T*>(&kraw[i][v]);
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          d0 = fmaf(qr[8 * v + e], DT<T>::to_f(kh[e]), d0);
          d1 = fmaf(qr[8 * v + e + 1], DT<T>::to_f(kh[e + 1]), d1);
        }
      }
      const float r = DT<T>::to_f(DT<T>::from_f(group_sum(d0 + d1)));
      sc[i] = (i0 + i < n_iter && row_of(i0 + i) < n_past) ? r : -INFINITY;
      m_new = fmaxf(m_new, sc[i]);
    }
    const float rescale = m_new == -INFINITY ? 1.f : __expf(m_w - m_new);   // exp(-inf) = 0 the first time
    s_w *= rescale;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] *= rescale;
#pragma unroll
    for (int i = 0; i < kPrefetchIters; ++i) {
      const float pl = sc[i] == -INFINITY ? 0.f : __expf(sc[i] - m_new);
      s_w += pl;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const T* vh = reinterpret_cast<const T*>(&vraw[i][v]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[8 * v + e] = fmaf(pl, DT<T>::to_f(vh[e]), acc[8 * v + e]);
      }
    }
    m_w = m_new;
  }
  if (live && rank == 0 && warp == 0) {   // the new token's own key / value (row slot 0 of warp 0; warp-uniform branch)
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) d = fmaf(qr[e], k_s[c * EPL + e], d);
    const float score = DT<T>::to_f(DT<T>::from_f(group_sum(d)));
    if (sub == 0) {
      float vr[EPL];
#pragma unroll
      for (int e = 0; e < EPL; ++e) vr[e] = v_s[c * EPL + e];
      visit(score, vr);
    }
  }
  // ---- merge the row slots of the warp (neighbouring slots first): all lanes end up with the warp's state
#pragma unroll
  for (int o = LPR; o <= 16; o <<= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m_w, o);
    const float s_o = __shfl_xor_sync(0xffffffffu, s_w, o);
    const float M2 = fmaxf(m_w, m_o);
    const float f_a = M2 == -INFINITY ? 0.f : __expf(m_w - M2), f_b = M2 == -INFINITY ? 0.f : __expf(m_o - M2);
    // (the lower slot's term first in both partners: identical bits on both sides)
    const bool low = (lane & o) == 0;
    s_w = low ? fmaf(s_o, f_b, s_w * f_a) : fmaf(s_w, f_a, s_o * f_b);
#pragma unroll
    for (int e = 0; e < EPL; ++e) {
      const float a_o = __shfl_xor_sync(0xffffffffu, acc[e], o);
      acc[e] = low ? fmaf(a_o, f_b, acc[e] * f_a) : fmaf(acc[e], f_a, a_o * f_b);
    }
    m_w = M2;
  }
  attn_stamp(p, 4);

  // ---- combine the warps (fixed order), then the cluster ranks (fixed order)
  if (lane == 0) {
    wred[warp] = m_w;
    wred[kAttnWarps + warp] = s_w;
  }
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < EPL; ++e) red[warp * DH + c * EPL + e] = acc[e];
  }
  __syncthreads();
  attn_stamp(p, 5);
  float M = -INFINITY, o = 0.f, ssum = 0.f;
  if (t < DH) {
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) M = fmaxf(M, wred[w]);
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float m_v = wred[w];
      const float f = m_v == -INFINITY ? 0.f : __expf(m_v - M);
      o = fmaf(red[w * DH + t], f, o);
      ssum = fmaf(wred[kAttnWarps + w], f, ssum);
    }
  }
  if (S > 1) {
    ptx::cluster_wait_acquire();                    // every CTA of the cluster is running (phase A)
    if (t < DH) ptx::st_cluster_f32(ptx::mapa_rank(ptx::smem_u32(xacc + rank * DH + t), 0), o);
    if (t == 0) {
      ptx::st_cluster_f32(ptx::mapa_rank(ptx::smem_u32(xstat + rank), 0), M);
      ptx::st_cluster_f32(ptx::mapa_rank(ptx::smem_u32(xstat + kMaxSplit + rank), 0), ssum);
    }
    ptx::cluster_arrive_release();
    ptx::cluster_wait_acquire();
    if (rank == 0 && t < DH) {
      float Mg = -INFINITY;
      for (int r = 0; r < S; ++r) Mg = fmaxf(Mg, xstat[r]);
      o = 0.f;
      ssum = 0.f;
      for (int r = 0; r < S; ++r) {
        const float m_v = xstat[r];
        const float f = m_v == -INFINITY ? 0.f : expf(m_v - Mg);
        o = fmaf(xacc[r * DH + t], f, o);
        ssum = fmaf(xstat[kMaxSplit + r], f, ssum);
      }
    }
  }
  if (live && rank == 0 && t < DH) static_cast<T*>(p.out)[h * DH + t] = DT<T>::from_f(o / ssum);
  attn_stamp(p, 6);
}

template <typename K, typename... Args>
int launch_pdl(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, args...));
  return CGQ_OK;
}

template <typename T, int DH>
int launch_attn(AttnParams p, cudaStream_t st) {
  // The cluster exchange costs ~1.5 us (DSMEM stores + barrier.cluster), one more 128-row pass of a single
  // CTA ~0.7 us: a KV window of up to 384 rows stays on one CTA per head, larger windows get one CTA per
  // 128 rows (16 warps x 8 rows in flight), powers of two up to the portable cluster size.
  int S = 1;
  if (p.max_len > 384)
    while (S < 8 && p.max_len > 128 * S) S *= 2;
  p.splits = S;
  const size_t smem = sizeof(float) * (3 * DH + kAttnWarps * DH + 2 * kAttnWarps + 2 * 8 + 8 * DH);
  auto kern = decode_attn_kernel<T, DH>;
  if (smem > 48 * 1024)
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.n_head * S);
  cfg.blockDim = dim3(kAttnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[na].val.programmaticStreamSerializationAllowed = 1;
  ++na;
  if (S > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = static_cast<unsigned>(S);
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  return CGQ_OK;
}

}  // namespace
}  // namespace cgq

using namespace cgq;

extern "C" int cgq_decode_begin_w4(const int64_t* ids, const uint8_t* Wq, const void* scale, void* x,
                                   int V, int D, int group, int dtype, int* state, void* stream) {
  if (V <= 0 || D <= 0 || group <= 0 || (group & 1) || V % group != 0 || ids == nullptr ||
      Wq == nullptr || scale == nullptr || x == nullptr || state == nullptr) {
    set_error("cgq_decode_begin_w4: bad arguments V=%d D=%d group=%d", V, D, group);
    return CGQ_ERR_BAD_SHAPE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((D + 255) / 256), block(256);
  if (dtype == CGQ_DTYPE_F16)
    return launch_pdl(decode_begin_w4_kernel<__half>, grid, block, 0, st, ids, Wq,
                      static_cast<const __half*>(scale), static_cast<__half*>(x), V, D, group, state);
  if (dtype == CGQ_DTYPE_BF16)
    return launch_pdl(decode_begin_w4_kernel<__nv_bfloat16>, grid, block, 0, st, ids, Wq,
                      static_cast<const __nv_bfloat16*>(scale), static_cast<__nv_bfloat16*>(x), V, D,
                      group, state);
  set_error("cgq_decode_begin_w4: bad dtype %d", dtype);
  return CGQ_ERR_BAD_DTYPE;
}

extern "C" int cgq_decode_begin_w8(const int64_t* ids, const int8_t* Wq, const void* scale, void* x, int V, int D,
                                   int dtype, int* state, void* stream) {
  if (V <= 0 || D <= 0 || ids == nullptr || Wq == nullptr || scale == nullptr || x == nullptr || state == nullptr) {
    set_error("cgq_decode_begin_w8: bad arguments V=%d D=%d", V, D);
    return CGQ_ERR_BAD_SHAPE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((D + 255) / 256), block(256);
  if (dtype == CGQ_DTYPE_F16)
    return launch_pdl(decode_begin_w8_kernel<__half>, grid, block, 0, st, ids, Wq, static_cast<const __half*>(scale),
                      static_cast<__half*>(x), V, D, state);
  if (dtype == CGQ_DTYPE_BF16)
    return launch_pdl(decode_begin_w8_kernel<__nv_bfloat16>, grid, block, 0, st, ids, Wq,
                      static_cast<const __nv_bfloat16*>(scale), static_cast<__nv_bfloat16*>(x), V, D, state);
  set_error("cgq_decode_begin_w8: bad dtype %d", dtype);
  return CGQ_ERR_BAD_DTYPE;
}

namespace {
thread_local const void* g_next_k = nullptr;
thread_local const void* g_next_v = nullptr;
}  // namespace
extern "C" void cgq_attention_next_kv(const void* kcache_next, const void* vcache_next) {
  g_next_k = kcache_next;
  g_next_v = vcache_next;
}

extern "C" int cgq_decode_attention(const void* qkv, const void* freqs, void* kcache, void* vcache,
                                    void* out, const int* state, int n_head, int n_groups,
                                    int d_head, int max_len, int dtype, void* stream) {
  if (n_head <= 0 || n_groups <= 0 || n_head % n_groups != 0 || max_len <= 0 ||
      (d_head != 64 && d_head != 128)) {
    set_error("cgq_decode_attention: bad shape n_head=%d n_groups=%d d_head=%d (64 or 128) max_len=%d",
              n_head, n_groups, d_head, max_len);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (qkv == nullptr || freqs == nullptr || kcache == nullptr || vcache == nullptr ||
      out == nullptr || state == nullptr ||
      ((reinterpret_cast<uintptr_t>(kcache) | reinterpret_cast<uintptr_t>(vcache)) & 15)) {
    set_error("cgq_decode_attention: null or misaligned pointer");
    return CGQ_ERR_MISALIGNED;
  }
  AttnParams p{qkv, freqs, kcache, vcache, out, state, n_head, n_groups, max_len, 1,
               static_cast<unsigned long long*>(take_trace_buffer()), g_next_k, g_next_v};
  g_next_k = g_next_v = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CGQ_DTYPE_F16)
    return d_head == 128 ? launch_attn<__half, 128>(p, st) : launch_attn<__half, 64>(p, st);
  if (dtype == CGQ_DTYPE_BF16)
    return d_head == 128 ? launch_attn<__nv_bfloat16, 128>(p, st)
                         : launch_attn<__nv_bfloat16, 64>(p, st);
  set_error("cgq_decode_attention: bad dtype %d", dtype);
  return CGQ_ERR_BAD_DTYPE;
}
