// One-launch decode step ("step program"): a whole batch-1 token of the int4g32 model -- embedding row, and per
// block RMSNorm+qkv_proj, RoPE/KV-append/attention, o_proj+residual, RMSNorm+w_in, SiLU*gate+w_out+residual, then
// final RMSNorm+lm_head (chatglm_q/model.py:230-246, 329-392) -- executed by ONE persistent kernel.
//
// Why: launched one kernel per linear, a token pays ~3.5 us of dependency bubble per launch
// (142 launches); the weights of the NEXT linear never depend on the previous one, only the 8 KB activation row
// does.  Here
//   * grid = one fat CTA per SM (cooperative launch): 1 TMA producer warp + 16 consumer warps (4 teams x 4 warps),
//     a shared-memory ring of ~21 stages x (8 KB packed weights + 1 KB scales) -- 190 KB per SM, 28 MB chip-wide;
//   * the producer lane walks the WHOLE step and keeps the ring full across phase boundaries: HBM streams the next
//     linears' weights while the consumers sit in a grid barrier / prologue / epilogue;
//   * a linear is cut into column SLICES of BW columns ([K/2 x BW] bytes, all of K): a slice belongs to ONE CTA, so
//     there is no cross-CTA reduction at all (no cluster, no workspace, no second pass): the 16 warps of the CTA
//     split the k-stages of the slice (team t takes stages t, t+4, ..), each warp one 1/4 of a stage's rows, and
//     meet once per slice in shared memory.  Slices are dealt round-robin, so at any moment the chip reads
//     full rows of the [K/2, N] byte matrix: sequential DRAM pages, every 128-byte line fetched once into L2
//     (TMA L2 promotion) and consumed by the 128/BW neighbouring CTAs;
//   * phases are separated by a flat grid barrier (one red.release + acquire-poll per CTA: 148 arrivals);
//   * the activation row of a linear is staged ONCE per CTA (not once per co-resident small CTA) straight into
//     MMA B-fragment order, with the fused prologue (RMSNorm / SiLU*gate) and the per-group activation sums
//     that carry the -8 offset of the nibbles (gemv_w4.cu's subnormal-operand arithmetic, fp16).
// Arithmetic per element is gemv_w4.cu's M == 1 fp16 path (nibbles fed to mma.sync.m16n8k16 as fp16 subnormals,
// group scale applied to the group's fp32 sum, reference roundings in prologue / epilogue); only the order in
// which the k-groups of a column are added differs (fixed, so results are bit-reproducible run to run).
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "w4_dev.cuh"

namespace cgq {
namespace {

using w4::ldcg128;
using w4::ldnc128;
using w4::PRO_NONE;
using w4::PRO_RMSNORM;
using w4::PRO_SILU_GATE;

constexpr int kTeams = 4;
constexpr int kTeamWarps = 4;
constexpr int kConsWarps = kTeams * kTeamWarps;
constexpr int kCons = kConsWarps * 32;            // 512 consumer threads
constexpr int kEpiWarp = kConsWarps;              // warp 16: slice epilogues (reduction, rounding, stores, TP exchange)
constexpr int kProdWarp = kConsWarps + 1;         // warp 17: TMA producer
constexpr int kSyncThreads = kCons + 32;          // consumers + epilogue warp meet in the grid barrier
constexpr int kMkThreads = kCons + 64;
constexpr int kWBytes = 8192, kSBytes = 1024;     // one ring stage: packed weights + scales
constexpr int kMaxBandK = 13824;                  // activation row capacity (k), multiple of every stage depth
constexpr int kBandBytes = kMaxBandK * 2;
constexpr int kGsumBytes = (kMaxBandK / 32) * 4;
constexpr int kMaxSmem = 232448;                  // 227 KB opt-in limit per CTA

enum { OP_LINEAR = CGQ_STEP_LINEAR, OP_ATTENTION = CGQ_STEP_ATTENTION, OP_EMBED = CGQ_STEP_EMBED };
enum { EPI_NONE = CGQ_EPI_NONE, EPI_SILU_PAIR = CGQ_EPI_SILU_PAIR };

template <int BW>
struct Geo {
  static constexpr int ROWS = kWBytes / BW;       // packed byte rows per stage
  static constexpr int KST = ROWS * 2;            // k per stage
  static constexpr int WR = ROWS / kTeamWarps;    // packed rows per warp
  static constexpr int RT = WR / 8;               // 8-row tiles per warp            (8 / 4 / 2)
  static constexpr int CC = BW / 16;              // 16-column chunks                (2 / 4 / 8)
  static constexpr int GP = RT / 2;               // quantisation groups per warp    (4 / 2 / 1)
  static constexpr int SWMASK = BW == 128 ? 7 : (BW == 64 ? 3 : 1);   // TMA swizzle: chunk ^= (row >> SWSH) & SWMASK
  static constexpr int SWSH = BW == 128 ? 0 : (BW == 64 ? 1 : 2);
  static constexpr int RED_BYTES = 2 * kConsWarps * BW * 4;
  static_assert(RT * CC == 16, "a warp-stage is 16 MMA tiles");
};

struct alignas(128) MkOp {
  CUtensorMap tmW, tmS;
  const void* A;
  const void* bias;
  const void* norm_w;
  const void* resid;
  void* C;
  const void* freqs;
  void* kcache;
  void* vcache;
  const int64_t* ids;
  int kind, N, K, slices, spk, prologue;
  float eps;
  int n_head, n_groups, max_len, V;
  int epi;       // EPI_SILU_PAIR: column n of the first half and of the second half are one CTA's consecutive slices
};

// slice sequence of CTA w for op o: (p, hh) -> slice index; EPI_SILU_PAIR walks (h slice p, gate slice p) pairs
struct SliceIter {
  int per, pairs;
  __device__ __forceinline__ explicit SliceIter(const MkOp& o)
      : per(o.epi == EPI_SILU_PAIR ? o.slices / 2 : o.slices), pairs(o.epi == EPI_SILU_PAIR ? 2 : 1) {}
};

// optional in-kernel timeline (cgq_debug_trace): 8 stamps per (op, CTA):
//   0 barrier passed   1 activation staged   2 first stage consumed (warp 0)   3 last stage consumed (warp 0)
//   4 last stage consumed (warp 15)   5 last slice stored (epilogue warp)   6 arrived at the next barrier
constexpr int kTraceSlots = 8;
__device__ __forceinline__ void stamp(unsigned long long* trace, int op, int slot) {
  if (trace != nullptr) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[(static_cast<size_t>(op) * gridDim.x + blockIdx.x) * kTraceSlots + slot] = t;
  }
}
__device__ __forceinline__ void cons_sync() { ptx::named_bar_sync(2, kCons); }          // the 16 consumer warps
__device__ __forceinline__ void step_sync() { ptx::named_bar_sync(1, kSyncThreads); }  // + the epilogue warp
// slice hand-off consumers -> epilogue warp, double-buffered, on hardware named barriers (3 + buf: partial columns
// written; 5 + buf: buffer free again): the writers `bar.arrive`, the reader `bar.sync`, and vice versa
__device__ __forceinline__ void bar_arrive(int id) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(kSyncThreads) : "memory");
}
__device__ __forceinline__ void bar_wait(int id) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kSyncThreads) : "memory");
}

// Consumer warps and the epilogue warp of all CTAs meet here between two dependent phases; the producer warps
// never do.  One release-add per CTA, relaxed polling (one L2 round trip per poll) and a single acquire fence.
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, unsigned long long* trace, int op) {
  if (threadIdx.x == 0 && op > 0) stamp(trace, op - 1, 6);
  step_sync();                                    // this CTA's global stores are issued
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned spins = 0, v;
    volatile unsigned* failed = ctr + 1;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      // a CTA is missing (seconds have passed): give up loudly instead of hanging the device; sticky flag,
      // later barriers fall through at once and cgq_step_status reports it
      if (++spins > (1u << 24) || *failed != 0) {
        *failed = 1;
        break;
      }
    }
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    stamp(trace, op, 0);
  }
  step_sync();
}

// 8 activations (k = 8c .. 8c+7) -> two B-fragment units [ (a0,a2), (a1,a3)*2^-4, (a4,a6), (a5,a7)*2^-4 ] and
// their sum as the MMA sees it (the odd-k values AFTER the 2^-4 scaling, times 16)
__device__ __forceinline__ uint4 frag8(const uint4& v, float& sum) {
  uint4 o;
  o.x = __byte_perm(v.x, v.y, 0x5410);
  o.y = w4::h2_mul(__byte_perm(v.x, v.y, 0x7632), 0x2C002C00u);
  o.z = __byte_perm(v.z, v.w, 0x5410);
  o.w = w4::h2_mul(__byte_perm(v.z, v.w, 0x7632), 0x2C002C00u);
  const __half2 e0 = *reinterpret_cast<const __half2*>(&o.x), d0 = *reinterpret_cast<const __half2*>(&o.y);
  const __half2 e1 = *reinterpret_cast<const __half2*>(&o.z), d1 = *reinterpret_cast<const __half2*>(&o.w);
  const float2 fe0 = __half22float2(e0), fd0 = __half22float2(d0), fe1 = __half22float2(e1), fd1 = __half22float2(d1);
  sum = ((fe0.x + fe0.y) + (fe1.x + fe1.y)) + 16.f * ((fd0.x + fd0.y) + (fd1.x + fd1.y));
  return o;
}

// ---- stage the activation row of a linear into shared memory (fragment order) with its fused prologue
template <int BW>
__device__ __forceinline__ void stage_activation(const MkOp& o, uint32_t Aband, float* gsum, float* sred) {
  using T = __half;
  using G = Geo<BW>;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = o.K, nchunk = K >> 3;
  const int total = o.spk * (G::KST / 8);         // chunks incl. the zero tail of a ragged last stage
  const T* A = static_cast<const T*>(o.A);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  // chunk c (8 k) of the prologue's output -> fragment order + the group's -8 * sum(a)
  auto emit = [&](int c, const uint4& v) {
    float s;
    const uint4 f = frag8(v, s);
    if (c < total) ptx::sts128(Aband + c * 16, f);
    s += __shfl_xor_sync(0xffffffffu, s, 1);      // a group of 32 k = 4 consecutive chunks = 4 consecutive lanes
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0 && c < total) gsum[c >> 2] = -8.f * s;
  };
  if (o.prologue == PRO_RMSNORM) {
    // round_T(round_T(x * rstd) * w), rstd over the whole row (model.py:68-73); the chunks a thread loads for
    // sum(x^2) stay in registers and are the ones it normalises (hidden size <= 2 * 512 * 8)
    const T* nw = static_cast<const T*>(o.norm_w);
    constexpr int U = 2;
    const bool in_regs = total <= U * kCons;
    uint4 keep[U], wr[U];
    float ss = 0.f;
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int c = tid + kCons * i;
        keep[i] = c < nchunk ? ldcg128(A + c * 8) : zero;
        wr[i] = c < nchunk ? ldnc128(nw + c * 8) : zero;
      }
#pragma unroll
      for (int i = 0; i < U; ++i) ss += w4::sumsq8<T>(keep[i]);
    } else {
      for (int c = tid; c < nchunk; c += kCons) ss += w4::sumsq8<T>(ldcg128(A + c * 8));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) sred[warp] = ss;
    cons_sync();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kConsWarps; ++w) tot += sred[w];
    const float rstd = rsqrtf(tot / static_cast<float>(K) + o.eps);
    if (in_regs) {
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const int c = tid + kCons * i;
        if (kCons * i < total) emit(c, c < nchunk ? w4::rmsnorm8<T>(keep[i], wr[i], rstd) : zero);
      }
    } else {
      for (int cb = 0; cb < total; cb += kCons) {
        const int c = cb + tid;
        emit(c, c < nchunk ? w4::rmsnorm8<T>(ldcg128(A + c * 8), ldnc128(nw + c * 8), rstd) : zero);
      }
    }
  } else if (o.prologue == PRO_SILU_GATE) {
    for (int cb = 0; cb < total; cb += kCons) {
      const int c = cb + tid;
      emit(c, c < nchunk ? w4::silu_gate8<T>(ldcg128(A + c * 8), ldcg128(A + K + c * 8)) : zero);
    }
  } else {
    for (int cb = 0; cb < total; cb += kCons) {
      const int c = cb + tid;
      emit(c, c < nchunk ? ldcg128(A + c * 8) : zero);
    }
  }
  cons_sync();
}

// ---- one ring stage of one warp: WR packed rows x BW columns
template <int BW>
__device__ __forceinline__ void consume_stage(uint32_t wst, uint32_t sst, uint32_t Aband, const float* gsum, int u,
                                              const uint32_t (&ld_off)[4], int wq, int g, int tig,
                                              uint64_t* empty_bar, uint32_t rt_zero, float (&tot)[Geo<BW>::CC][2],
                                              int dbg) {
  using G = Geo<BW>;
  using T = __half;
  const int lane = threadIdx.x & 31;
  uint32_t w[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t addr = wst + ld_off[i];
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(w[i][0]), "=r"(w[i][1]), "=r"(w[i][2]), "=r"(w[i][3])
                 : "r"(addr));
  }
  // B fragments of this warp's RT row tiles (token 0 only: lanes g == 0)
  const int unit0 = (u * G::KST + wq * G::WR * 2) / 4 + tig;       // + 4 per row tile (16 k)
  uint2 bf[G::RT];
#pragma unroll
  for (int rt = 0; rt < G::RT; ++rt) {
    bf[rt] = make_uint2(0u, 0u);
    if (g == 0) bf[rt] = ptx::lds64(Aband + (unit0 + 4 * rt) * 8);
  }
  const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
  float grp[G::GP][G::CC][4];
  // (the MMAs are volatile asm and issue in source order: the two dependent MMAs of a (group, chunk) tile pair are
  // kept 8 MMAs apart -- back to back they stall ~30 cycles each on the accumulator)
  if (dbg & 1) {      // diagnosis only (CGQ_STEP_DBG=1): no unpack / MMA work, everything else unchanged
#pragma unroll
    for (int gp = 0; gp < G::GP; ++gp)
#pragma unroll
      for (int cc = 0; cc < G::CC; ++cc) grp[gp][cc][0] = grp[gp][cc][2] = __uint_as_float(w[gp % 4][cc % 4]);
  } else
#pragma unroll
  for (int h = 0; h < 2; ++h) {                     // the two row tiles (16 k each) of a group
#pragma unroll
    for (int gp = 0; gp < G::GP; ++gp) {
#pragma unroll
      for (int cc = 0; cc < G::CC; ++cc) {
        const int i = gp * (G::CC / 2) + (cc >> 1);
        const uint32_t x = w[i][h + 2 * (cc & 1)];
        const uint32_t z = x >> 8;
        const uint32_t a[4] = {x & 0x000F000Fu, z & 0x000F000Fu, x & 0x00F000F0u, z & 0x00F000F0u};
        if (h == 0)
          ptx::mma_16816(grp[gp][cc], a, bf[2 * gp].x, bf[2 * gp].y, zero4, T());
        else
          ptx::mma_16816(grp[gp][cc], a, bf[2 * gp + 1].x, bf[2 * gp + 1].y, grp[gp][cc], T());
      }
    }
  }
  // group scales of columns 16 cc + 2 g, + 1 and the groups' -8 * sum(a)
  uint32_t sw[G::GP][G::CC];
  float c0[G::GP];
#pragma unroll
  for (int gp = 0; gp < G::GP; ++gp) {
#pragma unroll
    for (int cc = 0; cc < G::CC; ++cc) sw[gp][cc] = ptx::lds32(sst + ((wq * G::GP + gp) * BW + 16 * cc + 2 * g) * 2);
    c0[gp] = gsum[u * (G::KST / 32) + wq * G::GP + gp];
  }
  __syncwarp();
  if (lane == 0) {
    uint32_t dep = 0;     // every load from the slot: the four ldmatrix and all scale words
#pragma unroll
    for (int i = 0; i < 4; ++i) dep |= w[i][3];
#pragma unroll
    for (int gp = 0; gp < G::GP; ++gp)
#pragma unroll
      for (int cc = 0; cc < G::CC; ++cc) dep |= sw[gp][cc];
    ptx::mbar_arrive_after_loads(empty_bar, dep, rt_zero);
  }
#pragma unroll
  for (int gp = 0; gp < G::GP; ++gp) {
#pragma unroll
    for (int cc = 0; cc < G::CC; ++cc) {
      union {
        uint32_t u32;
        T h[2];
      } cv;
      cv.u32 = sw[gp][cc];
      const float t0 = fmaf(grp[gp][cc][0], 16777216.f, c0[gp]);
      const float t2 = fmaf(grp[gp][cc][2], 16777216.f, c0[gp]);
      tot[cc][0] = fmaf(DT<T>::to_f(cv.h[0]), t0, tot[cc][0]);
      tot[cc][1] = fmaf(DT<T>::to_f(cv.h[1]), t2, tot[cc][1]);
    }
  }
}

// ---- attention of one head for the new token (decode_step.cu's decode_attn_kernel, one CTA per head)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ __half ldcg_h(const __half* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return *reinterpret_cast<__half*>(&v);
}

template <int DH>
__device__ __forceinline__ void attention_head(const MkOp& o, int h, int n_past, float* sm) {
  using T = __half;
  constexpr int EPL = DH / 32;
  constexpr int kRows = 8;
  float* q_s = sm;
  float* k_s = q_s + DH;
  float* v_s = k_s + DH;
  float* red = v_s + DH;                    // [kConsWarps][DH]
  float* wred = red + kConsWarps * DH;      // [2][kConsWarps]
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int hpg = o.n_head / o.n_groups, g = h / hpg;
  const bool live = n_past < o.max_len;
  const T* qkv = static_cast<const T*>(o.A);
  const T* fr = static_cast<const T*>(o.freqs) + static_cast<size_t>(n_past + 1) * DH;   // position id = n_past + 1
  T* kc = static_cast<T*>(o.kcache);
  T* vc = static_cast<T*>(o.vcache);
  const size_t row_stride = static_cast<size_t>(o.n_groups) * DH;
  const bool writer = (h % hpg) == 0;
  const T* kbase = kc + g * DH;
  const T* vbase = vc + g * DH;
  if (live && t < DH) {
    const bool is_q = t < DH / 2;
    const int j = is_q ? t : t - DH / 2;
    const float fc = DT<T>::to_f(fr[2 * j]), fs = DT<T>::to_f(fr[2 * j + 1]);
    const T* src = is_q ? qkv + h * DH : qkv + (o.n_head + g) * DH;
    const float a = DT<T>::to_f(ldcg_h(src + 2 * j)), b = DT<T>::to_f(ldcg_h(src + 2 * j + 1));
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
    const T v = ldcg_h(qkv + (o.n_head + o.n_groups + g) * DH + d);
    v_s[d] = DT<T>::to_f(v);
    if (writer) vc[n_past * row_stride + g * DH + d] = v;
  }
  cons_sync();
  float qr[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) qr[e] = q_s[lane * EPL + e];
  float m_w = -INFINITY, s_w = 0.f, acc[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) acc[e] = 0.f;
  const int nblk = live ? (n_past + kConsWarps - 1) / kConsWarps : 0;
  for (int i0 = 0; i0 < nblk; i0 += kRows) {
    uint2 kraw[kRows], vraw[kRows];
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int l = (i0 + i) * kConsWarps + warp;
      kraw[i] = make_uint2(0u, 0u);
      vraw[i] = make_uint2(0u, 0u);
      if (i0 + i < nblk && l < n_past) {
        if (EPL == 4) {
          asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(kraw[i].x), "=r"(kraw[i].y) : "l"(kbase + l * row_stride + lane * 4));
          asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(vraw[i].x), "=r"(vraw[i].y) : "l"(vbase + l * row_stride + lane * 4));
        } else {
          asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(kraw[i].x) : "l"(kbase + l * row_stride + lane * 2));
          asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(vraw[i].x) : "l"(vbase + l * row_stride + lane * 2));
        }
      }
    }
    float sr[kRows];
    float mb = -INFINITY;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      sr[i] = -INFINITY;
      if (i0 + i < nblk && (i0 + i) * kConsWarps + warp < n_past) {
        const T* kh = reinterpret_cast<const T*>(&kraw[i]);
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < EPL; ++e) d = fmaf(qr[e], DT<T>::to_f(kh[e]), d);
        sr[i] = DT<T>::to_f(DT<T>::from_f(warp_sum(d)));
        mb = fmaxf(mb, sr[i]);
      }
    }
    if (mb != -INFINITY) {
      const float m_new = fmaxf(m_w, mb);
      const float rescale = expf(m_w - m_new);
      s_w *= rescale;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[e] *= rescale;
#pragma unroll
      for (int i = 0; i < kRows; ++i) {
        if (sr[i] != -INFINITY) {
          const float pl = expf(sr[i] - m_new);
          const T* vh = reinterpret_cast<const T*>(&vraw[i]);
          s_w += pl;
#pragma unroll
          for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pl, DT<T>::to_f(vh[e]), acc[e]);
        }
      }
      m_w = m_new;
    }
  }
  if (live && warp == 0) {   // the new token's own key / value
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) d = fmaf(qr[e], k_s[lane * EPL + e], d);
    const float sn = DT<T>::to_f(DT<T>::from_f(warp_sum(d)));
    const float m_new = fmaxf(m_w, sn);
    const float rescale = expf(m_w - m_new);
    const float pl = expf(sn - m_new);
    s_w = s_w * rescale + pl;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[e] = fmaf(pl, v_s[lane * EPL + e], acc[e] * rescale);
    m_w = m_new;
  }
  if (lane == 0) {
    wred[warp] = m_w;
    wred[kConsWarps + warp] = s_w;
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) red[warp * DH + lane * EPL + e] = acc[e];
  cons_sync();
  if (live && t < DH) {
    float M = -INFINITY, ov = 0.f, ssum = 0.f;
#pragma unroll
    for (int w = 0; w < kConsWarps; ++w) M = fmaxf(M, wred[w]);
#pragma unroll
    for (int w = 0; w < kConsWarps; ++w) {
      const float m_v = wred[w];
      const float f = m_v == -INFINITY ? 0.f : expf(m_v - M);
      ov = fmaf(red[w * DH + t], f, ov);
      ssum = fmaf(wred[kConsWarps + w], f, ssum);
    }
    static_cast<T*>(o.C)[h * DH + t] = DT<T>::from_f(ov / ssum);
  }
  cons_sync();     // the scratch aliases the activation band of the next linear
}

// ============================================================================================ the kernel
// SwiGLU of one element, the reference's roundings (model.py:200-201; same fast exp / divide as w4::silu_gate8)
__device__ __forceinline__ __half silu_mul(__half h, __half gate) {
  const float x = __half2float(h);
  const __half act = __float2half_rn(__fdividef(x, 1.f + __expf(-x)));
  return __float2half_rn(__half2float(act) * __half2float(gate));
}

template <int BW>
__global__ void __launch_bounds__(kMkThreads, 1)
    w4_step_kernel(const MkOp* __restrict__ ops, int n_ops, int S, unsigned* __restrict__ ctr, int* __restrict__ state,
                   unsigned long long* __restrict__ trace, int dbg) {
  using G = Geo<BW>;
  using T = __half;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t Wsm = base;
  const uint32_t Ssm = Wsm + S * kWBytes;
  const uint32_t off_band = S * (kWBytes + kSBytes);
  const uint32_t Aband = base + off_band;
  float* band_f = reinterpret_cast<float*>(gen + off_band);               // attention scratch aliases the band
  float* gsum = reinterpret_cast<float*>(gen + off_band + kBandBytes);
  float* red = reinterpret_cast<float*>(gen + off_band + kBandBytes + kGsumBytes);
  float* sred = reinterpret_cast<float*>(gen + off_band + kBandBytes + kGsumBytes + G::RED_BYTES);   // [16]
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_band + kBandBytes + kGsumBytes + G::RED_BYTES + 128);
  uint64_t* empty = full + S;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = gridDim.x, w = blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], kTeamWarps);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();
  __syncthreads();

  if (warp == kProdWarp) {
    // =========================== producer: one lane walks the whole step ===========================
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_first();
      unsigned issued = 0;
      for (int op = 0; op < n_ops; ++op) {
        const MkOp* o = ops + op;
        if (o->kind != OP_LINEAR) continue;
        const SliceIter it(*o);
        const int spk = o->spk;
        for (int p = w; p < it.per; p += W) {
          for (int hh = 0; hh < it.pairs; ++hh) {
            const int sl = p + hh * it.per;
            for (int u = 0; u < spk; ++u) {
              const int slot = issued % S;
              if (issued >= static_cast<unsigned>(S)) ptx::mbar_wait(&empty[slot], ((issued / S) - 1) & 1);
              ptx::mbar_expect_tx(&full[slot], kWBytes + kSBytes);
              ptx::tma_load_2d(gen + slot * kWBytes, &o->tmW, sl * BW, u * G::ROWS, &full[slot], pol);
              ptx::tma_load_2d(gen + S * kWBytes + slot * kSBytes, &o->tmS, sl * BW, u * (G::ROWS / 16), &full[slot], pol);
              ++issued;
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarp) {
    // =========================== epilogue warp ===========================
    // Sums the 16 consumer warps' partial columns of a finished slice (fixed order), rounds, adds bias / residual
    // and stores -- while the consumer warps are already in the next slice.  Joins every grid barrier.
    unsigned eseq = 0;
    for (int op = 0; op < n_ops; ++op) {
      if (op > 0) grid_barrier(ctr, static_cast<unsigned>(op) * W, trace, op);
      const MkOp& o = ops[op];
      if (o.kind != OP_LINEAR) continue;
      const SliceIter it(o);
      const T* bias = static_cast<const T*>(o.bias);
      const T* resid = static_cast<const T*>(o.resid);
      T* C = static_cast<T*>(o.C);
      for (int p = w; p < it.per; p += W) {
        T keep[BW / 32];
        for (int hh = 0; hh < it.pairs; ++hh, ++eseq) {
          const int sl = p + hh * it.per, buf = eseq & 1;
          bar_wait(3 + buf);
          const float* rb = red + buf * (kConsWarps * BW);
#pragma unroll
          for (int j = 0; j < BW / 32; ++j) {
            const int c = lane + 32 * j;
            float acc = 0.f;
#pragma unroll
            for (int ww = 0; ww < kConsWarps; ++ww) acc += rb[ww * BW + c];
            const int n = sl * BW + c;
            if (o.epi == EPI_SILU_PAIR) {
              const T v = epilogue<T>(acc, bias, n);
              if (hh == 0)
                keep[j] = v;                                             // h, rounded like w_in's output
              else
                C[p * BW + c] = silu_mul(keep[j], v);                    // u = silu(h) * gate
            } else if (n < o.N) {
              C[n] = w4::add_resid<T>(epilogue<T>(acc, bias, n), resid, n);
            }
          }
          bar_arrive(5 + buf);
        }
      }
      if (lane == 0) stamp(trace, op, 5);
    }
  } else {
  // =========================== consumers ===========================
  const int tid = threadIdx.x;
  const int team = warp / kTeamWarps, wq = warp % kTeamWarps;
  const int g = lane >> 2, tig = lane & 3;
  // tokens already in the KV cache (state[0]); CTA 0 advances it once every CTA has read it (after barrier 1)
  int n_past = 0;
  if (state != nullptr) asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(n_past) : "l"(state));
  // ldmatrix addresses of this lane inside a stage (loop-invariant): x4 number i covers the 2 x 2 block of
  // (row tile, column chunk) = (2 (i / (CC/2)) + (lm & 1), 2 (i % (CC/2)) + (lm >> 1)), lm = matrix of this lane
  uint32_t ld_off[4];
  {
    const int li = lane & 7, lm = lane >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rt = 2 * (i / (G::CC / 2)) + (lm & 1), cc = 2 * (i % (G::CC / 2)) + (lm >> 1);
      const int r = wq * G::WR + rt * 8 + li;
      ld_off[i] = static_cast<uint32_t>(r * BW + ((cc ^ ((r >> G::SWSH) & G::SWMASK)) << 4));
    }
  }
  unsigned seen = 0, sseq = 0;
  for (int op = 0; op < n_ops; ++op) {
    const MkOp& o = ops[op];
    if (op > 0) {
      grid_barrier(ctr, static_cast<unsigned>(op) * W, trace, op);
      if (op == 1 && w == 0 && tid == 0 && state != nullptr) state[0] = n_past + 1;
    }
    if (o.kind == OP_EMBED) {
      // QEmbedding row of the token (int4/qlinear.py:122-130): x[d] = round((nib - 8) * scale)
      int64_t t = o.ids[0];
      t = t < 0 ? 0 : (t >= o.V ? o.V - 1 : t);
      const int d = w * kCons + tid;
      if (d < o.N) {
        const uint8_t* wrow = static_cast<const uint8_t*>(o.A) + (t >> 1) * o.N;
        const T* srow = static_cast<const T*>(o.norm_w) + (t / 32) * o.N;
        static_cast<T*>(o.C)[d] = dequant4<T>((wrow[d] >> (static_cast<int>(t & 1) * 4)) & 0xF, srow[d]);
      }
      continue;
    }
    if (o.kind == OP_ATTENTION) {
      for (int h = w; h < o.n_head; h += W) {
        if (o.K == 128)
          attention_head<128>(o, h, n_past, band_f);
        else
          attention_head<64>(o, h, n_past, band_f);
      }
      continue;
    }
    // ---- linear
    const SliceIter it(o);
    const int spk = o.spk;
    if (w >= it.per) continue;
    if (!(dbg & 2)) stage_activation<BW>(o, Aband, gsum, sred);     // (CGQ_STEP_DBG=2: diagnosis, no prologue)
    if (tid == 0) stamp(trace, op, 1);
    bool first_done = false;
    const uint32_t rt_zero = static_cast<uint32_t>(o.K) >> 31;
    for (int p = w; p < it.per; p += W) {
      for (int hh = 0; hh < it.pairs; ++hh, ++sseq) {
        float tot[G::CC][2];
#pragma unroll
        for (int cc = 0; cc < G::CC; ++cc) tot[cc][0] = tot[cc][1] = 0.f;
        for (int u = 0; u < spk; ++u, ++seen) {
          if (static_cast<int>(seen % kTeams) != team) continue;
          const int slot = seen % S;
          ptx::mbar_wait(&full[slot], (seen / S) & 1);
          consume_stage<BW>(Wsm + slot * kWBytes, Ssm + slot * kSBytes, Aband, gsum, u, ld_off, wq, g, tig, &empty[slot],
                            rt_zero, tot, dbg);
          if (trace != nullptr && !first_done && tid == 0) stamp(trace, op, 2);
          first_done = true;
        }
        // ---- hand the warp's partial columns of this slice to the epilogue warp (double-buffered)
        const int buf = sseq & 1;
        if (sseq >= 2) bar_wait(5 + buf);
        float* rb = red + buf * (kConsWarps * BW);
        if (tig == 0) {
#pragma unroll
          for (int cc = 0; cc < G::CC; ++cc) {
            rb[warp * BW + 16 * cc + 2 * g] = tot[cc][0];
            rb[warp * BW + 16 * cc + 2 * g + 1] = tot[cc][1];
          }
        }
        bar_arrive(3 + buf);
      }
    }
    if (lane == 0 && warp == 0) stamp(trace, op, 3);
    if (lane == 0 && warp == kConsWarps - 1) stamp(trace, op, 4);
  }
  }  // consumers
  // (no warp leaves early: every role falls through to here, which keeps the named barriers' thread sets intact
  // for compute-sanitizer's synccheck)
}

// ------------------------------------------------------------------ host: step objects
struct Step {
  MkOp* d_ops = nullptr;
  unsigned* d_ctr = nullptr;   // [0] grid-barrier counter, [1] failure flag
  int* state = nullptr;
  int n_ops = 0, grid = 0, stages = 0, bw = 32, device = 0;
  size_t smem = 0;
};
std::mutex g_mu;
std::unordered_map<uint64_t, Step> g_steps;
uint64_t g_next_handle = 1;

template <int BW>
size_t fixed_smem() {
  return 1024 + kBandBytes + kGsumBytes + Geo<BW>::RED_BYTES + 128 + 64;
}
template <int BW>
int configure(Step& st) {
  auto kern = w4_step_kernel<BW>;
  int stages = static_cast<int>((kMaxSmem - fixed_smem<BW>()) / (kWBytes + kSBytes + 16));
  static const int env = [] {
    const char* s = getenv("CGQ_STEP_STAGES");
    return s != nullptr ? atoi(s) : 0;
  }();
  if (env >= 2 && env < stages) stages = env;
  st.stages = stages;
  st.smem = fixed_smem<BW>() + static_cast<size_t>(stages) * (kWBytes + kSBytes + 16);
  CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(st.smem)));
  int occ = 0;
  CGQ_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kMkThreads, st.smem));
  if (occ < 1) {
    set_error("cgq_step_create: the step kernel does not fit on an SM (%zu bytes of shared memory)", st.smem);
    return CGQ_ERR_UNSUPPORTED;
  }
  st.grid = sm_count();
  return CGQ_OK;
}
template <int BW>
int launch_step(const Step& st, cudaStream_t stream) {
  auto kern = w4_step_kernel<BW>;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(st.grid);
  cfg.blockDim = dim3(kMkThreads);
  cfg.dynamicSmemBytes = st.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident or the launch fails: no barrier deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CGQ_CUDA_TRY(cudaMemsetAsync(st.d_ctr, 0, 2 * sizeof(unsigned), stream));
  static const int dbg = [] {
    const char* s = getenv("CGQ_STEP_DBG");       // diagnosis switches (wrong results): 1 no MMA work, 2 no prologue
    return s != nullptr ? atoi(s) : 0;
  }();
  CGQ_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, static_cast<const MkOp*>(st.d_ops), st.n_ops, st.stages, st.d_ctr,
                                  st.state, static_cast<unsigned long long*>(take_trace_buffer()), dbg));
  return CGQ_OK;
}

int find(uint64_t handle, Step* out, const char* fn) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_steps.find(handle);
  if (it == g_steps.end()) {
    set_error("%s: unknown step handle", fn);
    return CGQ_ERR_BAD_SHAPE;
  }
  *out = it->second;
  return CGQ_OK;
}

int slice_width() {
  static const int bw = [] {
    const char* s = getenv("CGQ_STEP_BW");
    const int v = s != nullptr ? atoi(s) : 32;
    return (v == 32 || v == 64 || v == 128) ? v : 32;
  }();
  return bw;
}

}  // namespace
}  // namespace cgq

using namespace cgq;

extern "C" int cgq_step_create(const cgq_step_op* ops, int n_ops, int dtype, int* state, uint64_t* handle) {
  const char* fn = "cgq_step_create";
  if (ops == nullptr || n_ops <= 0 || handle == nullptr) {
    set_error("%s: null / empty step", fn);
    return CGQ_ERR_BAD_SHAPE;
  }
  if (dtype != CGQ_DTYPE_F16) {
    set_error("%s: the one-launch step is built for float16 (the reference checkpoints' dtype), got dtype code %d", fn, dtype);
    return CGQ_ERR_BAD_DTYPE;
  }
  Step st;
  st.bw = slice_width();
  int rc = st.bw == 32 ? configure<32>(st) : (st.bw == 64 ? configure<64>(st) : configure<128>(st));
  if (rc != CGQ_OK) return rc;
  const int BW = st.bw, rows = kWBytes / BW;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  std::vector<MkOp> host(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    const cgq_step_op& s = ops[i];
    MkOp& d = host[i];
    memset(&d, 0, sizeof(d));
    d.kind = s.kind;
    d.A = s.A;
    d.C = s.C;
    if (s.kind == CGQ_STEP_LINEAR) {
      if (s.N <= 0 || s.K <= 0 || s.K % 32 != 0 || s.N % BW != 0 || s.K > kMaxBandK || s.Wq == nullptr ||
          s.scale == nullptr || s.A == nullptr || s.C == nullptr || !al16(s.Wq) || !al16(s.scale) || !al16(s.A) ||
          (s.prologue == CGQ_PRO_RMSNORM && (s.norm_w == nullptr || !al16(s.norm_w))) ||
          (s.prologue != CGQ_PRO_NONE && s.prologue != CGQ_PRO_RMSNORM && s.prologue != CGQ_PRO_SILU_GATE)) {
        set_error("%s: op %d: bad linear (N=%d must be a multiple of %d, K=%d a multiple of 32 and <= %d, prologue=%d)",
                  fn, i, s.N, BW, s.K, kMaxBandK, s.prologue);
        return CGQ_ERR_BAD_SHAPE;
      }
      d.bias = s.bias;
      d.norm_w = s.norm_w;
      d.resid = s.resid;
      d.N = s.N;
      d.K = s.K;
      d.prologue = s.prologue;
      d.eps = s.eps;
      d.epi = s.epilogue;
      if (s.epilogue != CGQ_EPI_NONE && (s.epilogue != CGQ_EPI_SILU_PAIR || s.N % (2 * BW) != 0 || s.resid != nullptr)) {
        set_error("%s: op %d: bad epilogue %d (CGQ_EPI_SILU_PAIR needs N a multiple of %d and no residual)", fn, i,
                  s.epilogue, 2 * BW);
        return CGQ_ERR_BAD_SHAPE;
      }
      d.slices = s.N / BW;
      d.spk = (s.K / 2 + rows - 1) / rows;
      const int swz = BW == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (BW == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      TmapKey kw{s.Wq, static_cast<uint64_t>(s.N), static_cast<uint64_t>(s.K / 2), static_cast<uint64_t>(s.N),
                 static_cast<uint32_t>(BW), static_cast<uint32_t>(rows), CU_TENSOR_MAP_DATA_TYPE_UINT8, swz};
      rc = get_tmap_2d(kw, &d.tmW);
      if (rc != CGQ_OK) return rc;
      TmapKey ks{s.scale, static_cast<uint64_t>(s.N), static_cast<uint64_t>(s.K / 32), static_cast<uint64_t>(s.N) * 2,
                 static_cast<uint32_t>(BW), static_cast<uint32_t>(rows / 16), CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                 CU_TENSOR_MAP_SWIZZLE_NONE};
      rc = get_tmap_2d(ks, &d.tmS);
      if (rc != CGQ_OK) return rc;
    } else if (s.kind == CGQ_STEP_ATTENTION) {
      if (s.n_head <= 0 || s.n_groups <= 0 || s.n_head % s.n_groups != 0 || s.max_len <= 0 ||
          (s.d_head != 64 && s.d_head != 128) || s.A == nullptr || s.C == nullptr || s.freqs == nullptr ||
          s.kcache == nullptr || s.vcache == nullptr || state == nullptr ||
          ((reinterpret_cast<uintptr_t>(s.kcache) | reinterpret_cast<uintptr_t>(s.vcache)) & 7)) {
        set_error("%s: op %d: bad attention (n_head=%d n_groups=%d d_head=%d max_len=%d, or a null pointer)", fn, i,
                  s.n_head, s.n_groups, s.d_head, s.max_len);
        return CGQ_ERR_BAD_SHAPE;
      }
      d.freqs = s.freqs;
      d.kcache = s.kcache;
      d.vcache = s.vcache;
      d.n_head = s.n_head;
      d.n_groups = s.n_groups;
      d.max_len = s.max_len;
      d.K = s.d_head;
    } else if (s.kind == CGQ_STEP_EMBED) {
      if (s.ids == nullptr || s.Wq == nullptr || s.scale == nullptr || s.C == nullptr || s.N <= 0 || s.V <= 0 ||
          s.V % 32 != 0) {
        set_error("%s: op %d: bad embedding (V=%d D=%d)", fn, i, s.V, s.N);
        return CGQ_ERR_BAD_SHAPE;
      }
      d.A = s.Wq;          // [V/2, D] packed along the vocabulary axis
      d.norm_w = s.scale;  // [V/32, D]
      d.ids = s.ids;
      d.N = s.N;
      d.V = s.V;
    } else {
      set_error("%s: op %d: unknown kind %d", fn, i, s.kind);
      return CGQ_ERR_BAD_SHAPE;
    }
  }
  st.n_ops = n_ops;
  st.state = state;
  CGQ_CUDA_TRY(cudaGetDevice(&st.device));
  CGQ_CUDA_TRY(cudaMalloc(&st.d_ops, sizeof(MkOp) * n_ops));
  CGQ_CUDA_TRY(cudaMalloc(&st.d_ctr, 2 * sizeof(unsigned)));
  CGQ_CUDA_TRY(cudaMemcpy(st.d_ops, host.data(), sizeof(MkOp) * n_ops, cudaMemcpyHostToDevice));
  std::lock_guard<std::mutex> lk(g_mu);
  *handle = g_next_handle++;
  g_steps[*handle] = st;
  return CGQ_OK;
}

extern "C" int cgq_step_run(uint64_t handle, void* stream) {
  Step st;
  int rc = find(handle, &st, "cgq_step_run");
  if (rc != CGQ_OK) return rc;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return st.bw == 32 ? launch_step<32>(st, s) : (st.bw == 64 ? launch_step<64>(st, s) : launch_step<128>(st, s));
}

extern "C" int cgq_step_status(uint64_t handle, int* ctas, int* stages, int* failed) {
  Step st;
  int rc = find(handle, &st, "cgq_step_status");
  if (rc != CGQ_OK) return rc;
  unsigned h[2] = {0, 0};
  CGQ_CUDA_TRY(cudaDeviceSynchronize());
  CGQ_CUDA_TRY(cudaMemcpy(h, st.d_ctr, sizeof(h), cudaMemcpyDeviceToHost));
  if (ctas != nullptr) *ctas = st.grid;
  if (stages != nullptr) *stages = st.stages;
  if (failed != nullptr) *failed = static_cast<int>(h[1]);
  return CGQ_OK;
}

extern "C" int cgq_step_destroy(uint64_t handle) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_steps.find(handle);
  if (it == g_steps.end()) return CGQ_OK;
  cudaFree(it->second.d_ops);
  cudaFree(it->second.d_ctr);
  g_steps.erase(it);
  return CGQ_OK;
}
