// Top-k / top-p sampling of ONE logits row in ONE launch.
// Replaces chatglm_q.decoder.top_p_sampling (chatglm_q/decoder.py:12-27), which is ~15 torch kernels per
// token (softmax over the vocabulary, a full 65 024-element sort, cumsum, masking, renormalisation, and
// torch.multinomial with its own host synchronisation) -- about a fifth of a fused decode step.
//
// What the reference computes, and what this kernel keeps of it:
//   probs = softmax(logits.float() / temperature)        fp32, over the whole vocabulary          (:14)
//   probs, indices = sort(probs, descending)[:top_k]      the sort is only needed for its head     (:15-17)
//   probs[(cumsum(probs) - probs) > top_p] = 0;  probs /= sum(probs)                                (:20-22)
//   token = indices[multinomial(probs, 1)]                torch: argmax(probs / q), q ~ Exp(1)      (:25-26)
// The head of the sort is found by an exact radix SELECT on the 16-bit logits (fp16 / bf16 order-preserving
// keys, two 8-bit digit passes), ties broken by the lower vocabulary index -- the order a stable descending
// sort gives.  The caller passes the Exp(1) variates `q` (drawn with the same generator call
// torch.multinomial makes), so with the same seed the kernel returns the reference's token.
//
// One CTA of 1024 threads (the row is 127 KB and comes from L2 where lm_head just wrote it; the work is
// ~200 K integer operations -- latency, not bandwidth):
//   P0  stage the row into shared memory as order-preserving keys (16-byte loads); every thread takes the
//       maximum of the 64 elements it owns;
//   F1  FAST PATH: the k-th largest of the 1024 thread maxima, tau, is a lower bound of the k-th largest
//       element (k threads hold an element >= tau), and for k << 1024 barely more than k elements reach it.
//       tau comes from an exact two-digit radix select over the 1024 maxima (per-warp private histograms,
//       lanes with the same digit merged with match.any before the shared-memory atomic);
//   F2  one scan: sum(exp(v - vmax)) and every element >= tau into the candidate list (ballot-aggregated);
//   S   SLOW PATH, only if more than 1024 elements reach tau (massive ties, top_k near 1024): the same radix
//       select over ALL elements (P1 high byte, P2 low byte inside the threshold bin -> threshold key t and
//       the per-warp tie counts, P3 collect key > t in any order and the first `need` ties in index order).
//       match.any costs ~100 cycles: 8 000 of them are 60 us, which is why this is not the common path;
//   P4  rank the <= 1024 candidates by counting (key desc, index asc), keep ranks < k, top-p mask on a
//       chunked warp scan, renormalise, argmax(p / q).
#include "common.cuh"

namespace cgq {
namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxTopK = 1024;
constexpr unsigned kFull = 0xffffffffu;

struct SampleParams {
  const void* logits;   // [V] T
  int V;
  int vpad;             // V rounded up to a multiple of 64 * kWarps
  int k;                // min(top_k, V)
  float top_p;
  float temperature;   // logits are DIVIDED by it, like the reference (decoder.py:14), not multiplied by 1/T
  const float* q;       // [k] Exp(1) variates or null
  long long* token;     // [1] or null
  float* probs;         // [k] or null
  long long* indices;   // [k] or null
};

template <typename T>
__device__ __forceinline__ float key_to_float(uint32_t key);
template <>
__device__ __forceinline__ float key_to_float<__half>(uint32_t key) {
  const uint16_t raw = static_cast<uint16_t>((key & 0x8000u) ? (key ^ 0x8000u) : ~key);
  return __half2float(__ushort_as_half(raw));
}
template <>
__device__ __forceinline__ float key_to_float<__nv_bfloat16>(uint32_t key) {
  const uint32_t raw = (key & 0x8000u) ? (key ^ 0x8000u) : (~key & 0xffffu);
  return __uint_as_float(raw << 16);
}

// two 16-bit floats in one word -> two keys that order like the numbers (negative: all bits flipped,
// non-negative: sign bit set)
__device__ __forceinline__ uint32_t keys2(uint32_t w) {
  const uint32_t neg = ((w & 0x80008000u) >> 15) * 0xffffu;
  return w ^ (neg | 0x80008000u);
}

// warp 0: the bin (from the top) in which the running count reaches k, and the count above that bin
__device__ __forceinline__ void select_bin(const uint32_t* tot, uint32_t k, int lane, uint32_t* out) {
  uint32_t c[8], sum = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i] = tot[255 - 8 * lane - i];
    sum += c[i];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  uint32_t a = inc - sum;
  if (a < k && k <= inc) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (a < k && a + c[i] >= k) {
        out[0] = static_cast<uint32_t>(255 - 8 * lane - i);
        out[1] = a;
      }
      a += c[i];
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1) top_p_sample_kernel(const SampleParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint32_t* keys32 = reinterpret_cast<uint32_t*>(smem);                       // vpad / 2 words
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + static_cast<size_t>(p.vpad) * 2);   // [kWarps][256]
  uint32_t* tot = hist + kWarps * 256;                                        // [256]
  unsigned long long* comp = reinterpret_cast<unsigned long long*>(tot + 256);   // [kMaxTopK] (key, ~index)
  float* sp = reinterpret_cast<float*>(comp + kMaxTopK);                      // [kMaxTopK] sorted probabilities
  int* sidx = reinterpret_cast<int*>(sp + kMaxTopK);                          // [kMaxTopK] sorted indices
  float* redf = reinterpret_cast<float*>(sidx + kMaxTopK);                    // [kWarps]
  uint32_t* redu = reinterpret_cast<uint32_t*>(redf + kWarps);                // [kWarps]
  uint32_t* sel = redu + kWarps;                                              // [16] bin / above per pass, counters

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int V = p.V, k = p.k;
  const uint32_t lt = (1u << lane) - 1u;

  // ---- P0: stage keys
  {
    const uint16_t* src = static_cast<const uint16_t*>(p.logits);
    const bool vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
    const int nvec = vec ? V >> 3 : 0;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    uint4* dst4 = reinterpret_cast<uint4*>(keys32);
    for (int i = tid; i < nvec; i += kThreads) {
      uint4 v = src4[i];
      v.x = keys2(v.x);
      v.y = keys2(v.y);
      v.z = keys2(v.z);
      v.w = keys2(v.w);
      dst4[i] = v;
    }
    uint16_t* k16 = reinterpret_cast<uint16_t*>(keys32);
    for (int i = nvec * 8 + tid; i < p.vpad; i += kThreads) {
      uint32_t kk = 0;
      if (i < V) kk = keys2(src[i]) & 0xffffu;
      k16[i] = static_cast<uint16_t>(kk);
    }
    for (int i = tid; i < kWarps * 256; i += kThreads) hist[i] = 0;
    if (tid < 16) sel[tid] = 0;
  }
  __syncthreads();

  const int seg = p.vpad / kWarps;            // elements owned by a warp, a multiple of 64
  const int steps = seg >> 6;
  const int e_base = warp * seg;
  const uint32_t* wkeys = keys32 + (e_base >> 1);
  uint32_t* myhist = hist + warp * 256;

  // ---- maximum of the elements this thread owns (invalid / padding elements count as key 0)
  uint32_t tmax = 0;
  for (int s = 0; s < steps; ++s) {
    const uint32_t kk = wkeys[s * 32 + lane];
    const int e0 = e_base + s * 64 + 2 * lane;
    if (e0 < V) tmax = max(tmax, kk & 0xffffu);
    if (e0 + 1 < V) tmax = max(tmax, kk >> 16);
  }
  {
    const uint32_t wmax = __reduce_max_sync(kFull, tmax);
    if (lane == 0) redu[warp] = wmax;
  }

  // one 8-bit digit pass of the radix select over the 1024 thread maxima
  auto tmax_pass = [&](bool act, uint32_t digit, uint32_t kk, uint32_t* out) {
    const uint32_t d = act ? digit : 256u;
    const uint32_t m = __match_any_sync(kFull, d);
    if (act && (m & lt) == 0) atomicAdd(&myhist[d], __popc(m));
    __syncthreads();
    if (tid < 256) {
      uint32_t t = 0;
#pragma unroll 8
      for (int w = 0; w < kWarps; ++w) t += hist[w * 256 + tid];
      tot[tid] = t;
    }
    __syncthreads();
    if (warp == 0) select_bin(tot, kk, lane, out);
    for (int i = tid; i < kWarps * 256; i += kThreads) hist[i] = 0;
    __syncthreads();
  };
  // ---- F1: tau = k-th largest thread maximum
  tmax_pass(true, tmax >> 8, static_cast<uint32_t>(k), sel + 8);
  const uint32_t ta = sel[8];
  tmax_pass((tmax >> 8) == ta, tmax & 0xffu, static_cast<uint32_t>(k) - sel[9], sel + 10);
  const uint32_t tau = (ta << 8) | sel[10];
  uint32_t gmax = redu[lane];
  gmax = __reduce_max_sync(kFull, gmax);
  const float vmax = key_to_float<T>(gmax) / p.temperature;

  // ---- F2: sum of exponentials, candidates = every element >= tau
  float se = 0.f;
  for (int s = 0; s < steps; ++s) {
    const uint32_t kk = wkeys[s * 32 + lane];
    const int e0 = e_base + s * 64 + 2 * lane;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t key = (kk >> (16 * h)) & 0xffffu;
      const bool valid = e0 + h < V;
      if (valid) se += expf(key_to_float<T>(key) / p.temperature - vmax);
      const bool cand = valid && key >= tau;
      const uint32_t bc = __ballot_sync(kFull, cand);
      if (bc != 0) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&sel[12], __popc(bc));
        base = __shfl_sync(kFull, base, 0);
        const uint32_t slot = base + __popc(bc & lt);
        if (cand && slot < static_cast<uint32_t>(kMaxTopK))
          comp[slot] = (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - static_cast<uint32_t>(e0 + h));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(kFull, se, o);
  if (lane == 0) redf[warp] = se;
  __syncthreads();
  int n_cand = static_cast<int>(sel[12]);       // >= k by construction of tau

  if (n_cand > kMaxTopK) {
    // ---- S: exact radix select over all elements (block-uniform branch)
    // P1: high-byte histogram
    for (int s = 0; s < steps; ++s) {
      const uint32_t kk = wkeys[s * 32 + lane];
      const int e0 = e_base + s * 64 + 2 * lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t key = (kk >> (16 * h)) & 0xffffu;
        const bool valid = e0 + h < V;
        const uint32_t d = valid ? (key >> 8) : 256u;
        const uint32_t m = __match_any_sync(kFull, d);
        if (valid && (m & lt) == 0) atomicAdd(&myhist[d], __popc(m));
      }
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t t = 0;
#pragma unroll 8
      for (int w = 0; w < kWarps; ++w) t += hist[w * 256 + tid];
      tot[tid] = t;
    }
    __syncthreads();
    if (warp == 0) select_bin(tot, static_cast<uint32_t>(k), lane, sel);
    for (int i = tid; i < kWarps * 256; i += kThreads) hist[i] = 0;
    __syncthreads();
    const uint32_t b1 = sel[0], above1 = sel[1];

    // P2: low-byte histogram inside bin b1
    for (int s = 0; s < steps; ++s) {
      const uint32_t kk = wkeys[s * 32 + lane];
      const int e0 = e_base + s * 64 + 2 * lane;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t key = (kk >> (16 * h)) & 0xffffu;
        const bool act = e0 + h < V && (key >> 8) == b1;
        const uint32_t d = act ? (key & 0xffu) : 256u;
        const uint32_t m = __match_any_sync(kFull, d);
        if (act && (m & lt) == 0) atomicAdd(&myhist[d], __popc(m));
      }
    }
    __syncthreads();
    if (tid < 256) {
      uint32_t t = 0;
#pragma unroll 8
      for (int w = 0; w < kWarps; ++w) t += hist[w * 256 + tid];
      tot[tid] = t;
    }
    __syncthreads();
    if (warp == 0) select_bin(tot, static_cast<uint32_t>(k) - above1, lane, sel + 2);
    __syncthreads();
    const uint32_t b2 = sel[2];
    const uint32_t tkey = (b1 << 8) | b2;
    const uint32_t c_gt = above1 + sel[3];                  // keys strictly above the threshold (< k)
    const uint32_t need = static_cast<uint32_t>(k) - c_gt;  // ties taken, in index order (>= 1)
    // ties held by the warps before this one
    uint32_t tie_rank = lane < warp ? hist[lane * 256 + b2] : 0u;
    tie_rank = __reduce_add_sync(kFull, tie_rank);

    // P3: collect key > t (any order) and the first `need` ties in index order
    for (int s = 0; s < steps; ++s) {
      const uint32_t kk = wkeys[s * 32 + lane];
      const int e0 = e_base + s * 64 + 2 * lane;
      bool tie[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t key = (kk >> (16 * h)) & 0xffffu;
        const bool valid = e0 + h < V;
        const bool gt = valid && key > tkey;
        tie[h] = valid && key == tkey;
        const uint32_t bg = __ballot_sync(kFull, gt);
        if (bg != 0) {
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(&sel[4], __popc(bg));
          base = __shfl_sync(kFull, base, 0);
          if (gt)
            comp[base + __popc(bg & lt)] =
                (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - static_cast<uint32_t>(e0 + h));
        }
      }
      const uint32_t t0 = __ballot_sync(kFull, tie[0]), t1 = __ballot_sync(kFull, tie[1]);
      if ((t0 | t1) != 0) {
        const uint32_t r0 = tie_rank + __popc(t0 & lt) + __popc(t1 & lt);
        const uint32_t r1 = r0 + (tie[0] ? 1u : 0u);
        if (tie[0] && r0 < need)
          comp[c_gt + r0] = (static_cast<unsigned long long>(tkey) << 32) | (0xffffffffu - static_cast<uint32_t>(e0));
        if (tie[1] && r1 < need)
          comp[c_gt + r1] = (static_cast<unsigned long long>(tkey) << 32) | (0xffffffffu - static_cast<uint32_t>(e0 + 1));
        tie_rank += __popc(t0) + __popc(t1);
      }
    }
    n_cand = k;
    __syncthreads();
  }

  // ---- P4: order the candidates (key descending, index ascending), keep the first k, probabilities
  float sumexp = redf[lane];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sumexp += __shfl_xor_sync(kFull, sumexp, o);
  if (tid < n_cand) {
    const unsigned long long mine = comp[tid];
    int r = 0;
    for (int j = 0; j < n_cand; ++j) r += comp[j] > mine ? 1 : 0;
    if (r < k) {
      const uint32_t key = static_cast<uint32_t>(mine >> 32);
      sp[r] = expf(key_to_float<T>(key) / p.temperature - vmax) / sumexp;
      sidx[r] = static_cast<int>(0xffffffffu - static_cast<uint32_t>(mine));
    }
  }
  __syncthreads();
  if (warp != 0) return;

  // top-p on a chunked scan: lane owns `per` consecutive ranks
  const int per = (k + 31) >> 5;
  const int r_lo = min(k, lane * per), r_hi = min(k, r_lo + per);
  float local = 0.f;
  for (int r = r_lo; r < r_hi; ++r) local += sp[r];
  float inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float v = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += v;
  }
  float cum = inc - local;       // probability mass of the ranks before r_lo
  float kept = 0.f;
  for (int r = r_lo; r < r_hi; ++r) {
    const float pr = sp[r];
    cum += pr;
    if ((cum - pr) > p.top_p) sp[r] = 0.f; else kept += pr;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(kFull, kept, o);
  float best = -1.f;
  int best_r = k;
  for (int r = r_lo; r < r_hi; ++r) {
    const float pr = sp[r] / kept;
    if (p.probs != nullptr) p.probs[r] = pr;
    if (p.indices != nullptr) p.indices[r] = sidx[r];
    if (p.q != nullptr) {
      const float sc = pr / p.q[r];
      if (sc > best) {
        best = sc;
        best_r = r;
      }
    }
  }
  if (p.q != nullptr && p.token != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(kFull, best, o);
      const int orr = __shfl_xor_sync(kFull, best_r, o);
      if (ob > best || (ob == best && orr < best_r)) {
        best = ob;
        best_r = orr;
      }
    }
    if (lane == 0) p.token[0] = sidx[min(best_r, k - 1)];
  }
}

size_t sample_smem_bytes(int vpad) {
  return static_cast<size_t>(vpad) * 2 + sizeof(uint32_t) * (kWarps * 256 + 256) +
         kMaxTopK * (sizeof(unsigned long long) + sizeof(float) + sizeof(int)) + kWarps * 8 + 16 * 4;
}

template <typename T>
int launch_sample(const SampleParams& p, cudaStream_t st) {
  const size_t smem = sample_smem_bytes(p.vpad);
  auto kern = top_p_sample_kernel<T>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || configured[dev] < smem) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (dev >= 0 && dev < 64) configured[dev] = smem;
  }
  kern<<<1, kThreads, smem, st>>>(p);
  CGQ_CUDA_TRY(cudaGetLastError());
  return CGQ_OK;
}

}  // namespace
}  // namespace cgq

extern "C" int cgq_top_p_sample(const void* logits, int V, int dtype, int top_k, float top_p,
                                float temperature, const float* q, int64_t* token, float* probs,
                                int64_t* indices, void* stream) {
  using namespace cgq;
  if (logits == nullptr || V <= 0 || top_k <= 0 || !(temperature > 0.f) || !(top_p >= 0.f) ||
      (reinterpret_cast<uintptr_t>(logits) & 1)) {
    set_error("cgq_top_p_sample: bad arguments V=%d top_k=%d top_p=%g temperature=%g", V, top_k,
              static_cast<double>(top_p), static_cast<double>(temperature));
    return CGQ_ERR_BAD_SHAPE;
  }
  const int k = top_k < V ? top_k : V;
  const int unit = 64 * kWarps;
  const int vpad = (V + unit - 1) / unit * unit;
  if (k > kMaxTopK || sample_smem_bytes(vpad) > 227u * 1024u) {
    set_error("cgq_top_p_sample: top_k=%d (max %d) or V=%d (row must fit in shared memory) not supported", top_k,
              kMaxTopK, V);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (token != nullptr && q == nullptr) {
    set_error("cgq_top_p_sample: a token is requested without the Exp(1) variates q");
    return CGQ_ERR_BAD_SHAPE;
  }
  SampleParams p{logits, V, vpad, k, top_p, temperature, q,
                 reinterpret_cast<long long*>(token), probs, reinterpret_cast<long long*>(indices)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CGQ_DTYPE_F16) return launch_sample<__half>(p, st);
  if (dtype == CGQ_DTYPE_BF16) return launch_sample<__nv_bfloat16>(p, st);
  set_error("cgq_top_p_sample: bad dtype %d (logits must be fp16 / bf16: the select works on 16-bit keys)", dtype);
  return CGQ_ERR_BAD_DTYPE;
}
