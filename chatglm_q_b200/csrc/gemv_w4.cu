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
//     tile of packed weights (128-byte swizzle) and the [4 x 128] scale tile on one mbarrier.  The ring
//     carries weights only.  Activations: M == 1 -- the consumers stage the CTA's k-band of the row ONCE
//     with plain L2 loads right after the dependency wait, through an optional fused prologue (RMSNorm
//     of the whole row, SiLU*gate), and the residual is added in the epilogue (cgq_w4a16_gemv_fused);
//     M > 1 -- every lane reads its own B-fragment words from L2 one stage ahead of use;
//   * 4 consumer warps each own one 32-k quantisation group of the stage.  A lane reads two
//     16-byte runs (16 columns, packed rows 2t and 2t+1), and one PRMT per column makes the
//     32-bit word [byte(r), -, byte(r+1), -], whose masked halves are exactly the (k, k') pairs
//     of an m16n8k16 A fragment with the weight COLUMN as the MMA row.  Activations are the
//     B fragment (<= 8 tokens), accumulation is fp32 in the tensor core;
//   * the group scale is applied to the group's fp32 partial sum:  out = Σ_g s_g · Σ_{k∈g} a_k (q_k-8)
//     (more accurate than the reference's per-element fp16 rounding; within the 1e-2 parity bar);
//   * M == 1, fp16 ("trick"): the masked nibble IS an fp16 subnormal q·2^-24 (q·2^-20 for the high
//     nibble, compensated by scaling the odd-k activations by 2^-4), so no int->fp conversion is
//     executed at all; the -8 offset becomes -8·Σ_{k∈g} a_k, obtained from one extra MMA against a
//     constant fragment.  bf16 and every M > 1 convert exactly ((1024+q) - 1032);
//   * a ring slot is released with ptx::mbar_arrive_after_loads: the mbarrier address depends on the
//     registers loaded from the slot, so the arrive cannot overtake an ld.shared that is still queued
//     (the producer answers a release with a TMA write into the same slot; a refill that hit L2 used to
//     land before the stage's scale loads in ~10 % of the M = 8 launches).
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "tmap.cuh"
#include "w4_dev.cuh"

namespace cgq {
int default_w4_arith(int set);
namespace {

using namespace w4;

template <bool kM1>
struct Cfg {
  static constexpr int MR = kM1 ? 1 : MMAX;           // token rows staged / reduced
  // Activations never go through the TMA ring.  M == 1: the activation band of the CTA is staged ONCE
  // by the consumer warps (plain L2 loads right after the dependency wait, optionally through a fused
  // prologue).  M > 1: every lane reads its own B-fragment words straight from L2, one stage ahead.
  static constexpr int A_BYTES = 0;
  static constexpr int RED_BYTES = CW * MR * BN * 4;
  // band sums of the Z cluster ranks, received by rank 0 (same layout in every CTA of the cluster)
  __host__ __device__ static constexpr int xred_bytes(int Z) { return Z > 1 ? Z * MR * BN * 4 : 0; }
  static constexpr int STAGE_BYTES = W_BYTES + S_BYTES + A_BYTES;
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
  // fused decode-step pieces (M == 1 only; cgq_w4a16_gemv_fused)
  const void* resid;   // [N] or null: C = resid + round(acc) (+bias)   (model.py:243,246 `x = x + h`)
  const void* norm_w;  // [K] RMSNorm weight (PRO_RMSNORM)
  float eps;
  int band_units;      // k-stages of activation band staged per CTA (max over the cluster ranks)
  // L2 prefetch of the NEXT decode launch's weights (cgq_prefetch_next_w4): the first pf_depth k-stages
  // of each of its pf_Z bands, in the lockstep order that launch will read them
  const uint8_t* pf_w;
  const uint8_t* pf_s;   // scales, 2 bytes per element
  int pf_N, pf_rows, pf_groups, pf_SPT, pf_Z, pf_depth, pf_ppc_w, pf_ppc, pf_pieces;
  // tile-granular hand-over (cgq_handover_next, kHand kernels only): the consumers acquire-poll `wait_ctr` until it
  // reaches `wait_count` INSTEAD of griddepcontrol.wait; every stored output tile release-increments `signal_ctr`
  const unsigned* wait_ctr;
  unsigned wait_count;
  unsigned* signal_ctr;
  unsigned hand_mode, hand_sleep;   // polling variant (CGQ_HAND_MODE bits, CGQ_HAND_SLEEP ns)
  // root-cause experiment only (CGQ_HACK_PLAIN_RELEASE=1, M > 1 kernels): release ring slots with a plain
  // mbarrier.arrive, i.e. the pre-fix protocol that lets the arrive overtake outstanding ld.shared
  int plain_release;
  // tensor parallelism (cgq_tp_next, M == 1 fused launches only; no reference counterpart -- SURVEY §8e):
  //  * tp.recv[0] != null: ROW-parallel linear.  The fp32 partial sum of every output column is exchanged with the
  //    peer ranks in the epilogue (8-byte {value, epoch} words stored straight into every rank's receive buffer over
  //    NVLink, NCCL-LL style: no fence, no flag, no collective launch) and summed in rank order -- every rank ends
  //    up with the bit-identical row;
  //  * tp.out[0] != null: COLUMN-parallel linear whose output every rank needs whole (lm_head): the rounded values
  //    are stored into every rank's output buffer at this rank's column offset.
  cgq_tp_ctx tp;
  unsigned tp_idx;
};
constexpr int kPfPiece = 16384;

__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// piece `idx` of the next launch's weight stream -> one bulk L2 prefetch
__device__ __forceinline__ void prefetch_piece(const Params& p, int idx) {
  const int chunk = idx / p.pf_ppc, j = idx - chunk * p.pf_ppc;
  const int i = chunk / p.pf_Z, z = chunk - i * p.pf_Z;
  const int u = p.pf_SPT * z / p.pf_Z + i;
  if (u >= p.pf_SPT * (z + 1) / p.pf_Z) return;
  if (j < p.pf_ppc_w) {
    const int row0 = u * ROWS;
    const int64_t bytes = static_cast<int64_t>(min(ROWS, p.pf_rows - row0)) * p.pf_N;
    const int64_t off = static_cast<int64_t>(j) * kPfPiece;
    if (off < bytes)
      prefetch_l2_bulk(p.pf_w + static_cast<int64_t>(row0) * p.pf_N + off,
                       static_cast<uint32_t>(bytes - off < kPfPiece ? bytes - off : kPfPiece));
  } else {
    const int g0 = u * CW;
    const int64_t bytes = static_cast<int64_t>(min(CW, p.pf_groups - g0)) * p.pf_N * 2;
    const int64_t off = static_cast<int64_t>(j - p.pf_ppc_w) * kPfPiece;
    if (off < bytes)
      prefetch_l2_bulk(p.pf_s + static_cast<int64_t>(g0) * p.pf_N * 2 + off,
                       static_cast<uint32_t>(bytes - off < kPfPiece ? bytes - off : kPfPiece));
  }
}

__device__ __forceinline__ void stamp(const Params& p, int slot) {
  if (p.trace != nullptr && blockIdx.x < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[blockIdx.x * 8 + slot] = t;
  }
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Activations of this launch are complete.  Default: the whole previous grid has drained and flushed
// (griddepcontrol.wait).  Hand-over variant: every output tile of the producing launch has been stored and
// announced -- no wait for the grid to drain (the ~1.2 us between its last CTA's exit and the wait returning).
// One lane per warp polls; the spin is bounded (a lost producer gives wrong numbers, never a hung device).
template <bool kHand>
__device__ __forceinline__ void wait_inputs(const Params& p) {
  if constexpr (kHand) {
    if (p.wait_ctr != nullptr) {
      const bool one_poller = (p.hand_mode & 1) != 0;     // thread 0 polls for the CTA, else lane 0 of every warp
      const bool relaxed = (p.hand_mode & 2) != 0;        // relaxed polls + one acquire at the end
      if (one_poller ? threadIdx.x == 0 : (threadIdx.x & 31) == 0) {
        unsigned spins = 0;
        if (relaxed) {
          unsigned v;
          do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_ctr) : "memory");
            if (v >= p.wait_count) break;
            __nanosleep(p.hand_sleep);
          } while (++spins < (1u << 15));
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
        } else {
          while (ld_acquire_gpu(p.wait_ctr) < p.wait_count && ++spins < (1u << 15)) __nanosleep(p.hand_sleep);   // <= ~25 ms
        }
      }
      if (one_poller)
        ptx::named_bar_sync(1, CW * 32);
      else
        __syncwarp();
      return;
    }
  }
  ptx::pdl_wait_prior_grid();
}
// the calling (consumer) threads have stored one output tile: announce it
template <bool kHand>
__device__ __forceinline__ void signal_tile(const Params& p) {
  if constexpr (kHand) {
    if (p.signal_ctr != nullptr) {
      ptx::named_bar_sync(1, CW * 32);
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p.signal_ctr), "r"(1u) : "memory");
      }
    }
  }
}

// One-shot all-reduce of a row-parallel linear's partial sums.  Word (slot, src rank, column) of a
// rank's receive buffer holds {fp32 partial, epoch}; epoch = step_no * 128 + exchange index + 1 is unique per
// (token, linear), the two slots alternate between consecutive exchanges.  A rank can only reach exchange i + 2
// after every peer has published exchange i + 1, i.e. after every peer has finished reading exchange i: two slots
// are enough.  The spin is bounded: a lost peer sets a sticky error word instead of hanging the device.
__device__ __forceinline__ float tp_allreduce(const cgq_tp_ctx& tp, unsigned idx, int step_no, float acc, int n) {
  const unsigned epoch = (static_cast<unsigned>(step_no) << 7) + idx + 1u;
  const size_t slot_words = static_cast<size_t>(tp.world) * tp.max_n;
  const size_t base = (idx & 1u) * slot_words;
  const size_t mine = base + static_cast<size_t>(tp.rank) * tp.max_n + n;
  for (int r = 0; r < tp.world; ++r) {
    uint2* dst = static_cast<uint2*>(tp.recv[r]) + mine;
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(acc)), "r"(epoch) : "memory");
  }
  float sum = 0.f;
  const uint2* own = static_cast<const uint2*>(tp.recv[tp.rank]) + base + n;
  for (int r = 0; r < tp.world; ++r) {
    const uint2* src = own + static_cast<size_t>(r) * tp.max_n;
    unsigned v, e, spins = 0;
    for (;;) {
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(e) : "l"(src) : "memory");
      if (e == epoch) break;
      if (++spins > (1u << 26)) {
        if (tp.err != nullptr) *tp.err = epoch;
        break;
      }
    }
    sum += __uint_as_float(v);
  }
  return sum;
}
// final value of output column n: (all-reduce) -> round -> bias -> residual -> store (to every rank for tp.out)
// bias[n] / resid[n] of the column a consumer thread will store, requested right after the activation band is staged:
// the tail of a 4 us launch should not wait for one more L2 round trip
template <typename T>
struct ColPre {
  T bias, resid;
};
template <typename T>
__device__ __forceinline__ ColPre<T> preload_column(const Params& p, int n) {
  ColPre<T> c;
  c.bias = c.resid = DT<T>::from_f(0.f);
  uint16_t b = 0, r = 0;
  if (n < p.N) {
    if (p.bias != nullptr) asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(b) : "l"(static_cast<const T*>(p.bias) + n));
    if (p.resid != nullptr) asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(static_cast<const T*>(p.resid) + n));
  }
  c.bias = reinterpret_cast<const T&>(b);
  c.resid = reinterpret_cast<const T&>(r);
  return c;
}
template <typename T>
__device__ __forceinline__ void store_column(const Params& p, float acc, int n, int step_no, const ColPre<T>& pre) {
  if (p.tp.world > 1 && p.tp.recv[0] != nullptr) acc = tp_allreduce(p.tp, p.tp_idx, step_no, acc, n);
  // round -> (+ bias, rounded) -> (+ residual, rounded): epilogue<T> / add_resid<T> on the preloaded values
  T val = DT<T>::from_f(acc);
  if (p.bias != nullptr) val = DT<T>::from_f(DT<T>::to_f(val) + DT<T>::to_f(pre.bias));
  if (p.resid != nullptr) val = DT<T>::from_f(DT<T>::to_f(pre.resid) + DT<T>::to_f(val));
  if (p.tp.world > 1 && p.tp.out[0] != nullptr) {
    for (int r = 0; r < p.tp.world; ++r) static_cast<T*>(p.tp.out[r])[p.tp.out_offset + n] = val;
  } else {
    static_cast<T*>(p.C)[n] = val;
  }
}

template <typename T, bool kTrick, bool kM1, int kPro, bool kHand = false, bool kImma = false>
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
  const uint32_t off_red = S * C::STAGE_BYTES;
  float* red = reinterpret_cast<float*>(gen + off_red);
  float* xred = reinterpret_cast<float*>(gen + off_red + C::RED_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(gen + off_red + C::RED_BYTES + C::xred_bytes(p.Z));
  uint64_t* empty = full + S;
  uint64_t* xbar = empty + S;    // rank 0: completes when the band sums of ranks 1 .. Z-1 have landed in xred
  // M == 1: activation band of this CTA (band_units x 128 k, zero beyond K), staged by the consumers
  const uint32_t off_band = (off_red + C::RED_BYTES + C::xred_bytes(p.Z) + 16u * S + 16u + 15u) & ~15u;
  const uint32_t Aband = base + off_band;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Z = p.Z;
  const int tile = blockIdx.x / Z, z = blockIdx.x - tile * Z;   // z == rank in the cluster
  const int u0 = p.SPT * z / Z, u1 = p.SPT * (z + 1) / Z;      // this CTA's k-stages of the tile
  const int n_units = u1 - u0;
  const T* A = static_cast<const T*>(p.A);
  // token counter of the step (published by cgq_decode_begin_w4 before it releases its dependents, like state[1])
  int step_no = 0;
  if (kM1 && p.tp.world > 1 && p.tp.step != nullptr)
    asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(step_no) : "l"(p.tp.step));

  ColPre<T> pre;
  pre.bias = pre.resid = DT<T>::from_f(0.f);

  // ---- prologue: barriers
  if (threadIdx.x == CW * 32) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmS);
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], CW);
    }
    if (p.Z > 1 && z == 0) {
      // the remote stores carry the arrival (st.async ... complete_tx): expected bytes = every other rank's band sums
      ptx::mbar_init(xbar, 1);
      ptx::mbar_expect_tx(xbar, static_cast<uint32_t>(p.Z - 1) * static_cast<uint32_t>(min(p.M, MR)) * BN * 4u);
    }
    ptx::fence_mbar_init();
  }
  __syncwarp();      // (the single-thread branch above must have reconverged before the aligned block barrier)
  __syncthreads();
  if (threadIdx.x == 0) stamp(p, 0);
  // Distributed shared memory may only be written once the target CTA is known to run: every CTA of the cluster
  // arrives here (non-blocking) and waits right before its first remote store (compute-sanitizer racecheck:
  // "block that might not have entered yet").  Free in practice -- the cluster is co-scheduled.
  if (p.Z > 1) ptx::cluster_arrive_release();
  // Let the next kernel in the stream start its own prologue / weight prefetch (PDL).
  ptx::pdl_launch_dependents();

  if (warp == CW) {
    // =========================== producer: lane 0 drives TMA ===========================
    // (every lane runs the loop, the issue is predicated on lane 0 inside the asm: a divergent single-lane region
    //  costs an ELECT / BRA.U.ANY waterfall per TMA / mbarrier instruction -- see ptx::tma_load_2d_warp)
    {
      const uint32_t issue = lane == 0 ? 1u : 0u;
      if (lane == 0) stamp(p, 1);
      const uint64_t pol = ptx::policy_evict_first();
      auto issue_w = [&](int i, int slot) {
        const int ks = u0 + i;
        ptx::mbar_expect_tx_warp(&full[slot], W_BYTES + S_BYTES, issue);
        ptx::tma_load_2d_warp(Wsm + slot * W_BYTES, &tmW, tile * BN, ks * ROWS, &full[slot], pol, issue);
        ptx::tma_load_2d_warp(Ssm + slot * S_BYTES, &tmS, tile * BN, ks * CW, &full[slot], pol, issue);
      };
      const int prefill = min(n_units, S);
      // weights do not depend on the previous kernel: start streaming them before the PDL wait
      for (int i = 0; i < prefill; ++i) issue_w(i, i);
      // next launch's weights -> L2, spread over this CTA's refill iterations (own loads first; off unless CGQ_PF_MB)
      int pf_next = blockIdx.x;
      const int pf_mine = p.pf_pieces > 0 ? (p.pf_pieces - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
      const int pf_iters = n_units - prefill;
      const int pf_per = pf_iters > 0 ? (pf_mine + pf_iters - 1) / pf_iters : pf_mine;
      auto pf_issue = [&](int count) {
        if (lane != 0) return;
        for (int c = 0; c < count && pf_next < p.pf_pieces; ++c, pf_next += gridDim.x)
          prefetch_piece(p, pf_next);
      };
      if (pf_iters <= 0) pf_issue(pf_mine);
      int slot = 0, phase = 1;
      for (int i = prefill; i < n_units; ++i) {
        ptx::mbar_wait(&empty[slot], phase ^ 1);
        issue_w(i, slot);
        pf_issue(pf_per);
        if (++slot == S) {
          slot = 0;
          phase ^= 1;
        }
      }
      pf_issue(pf_mine);
    }
    __syncwarp();
    if (p.Z > 1) ptx::cluster_wait_acquire();    // phase A of the cluster barrier (see the prologue)
  } else {
  // =========================== consumers ===========================
  if (kM1) {
    // ---- stage the activation band (16-byte chunks of 8 k, zero beyond K) with the fused prologue
    const int tid = threadIdx.x;                 // 0 .. CW*32-1
    const int nchunk = p.K >> 3;                 // valid chunks of the activation row
    const int c_lo = u0 * (KSTAGE / 8), c_hi = u1 * (KSTAGE / 8);
    const uint32_t dig_info = Aband + p.band_units * DIG_STAGE;   // kImma: per-group 2^-shift behind the digits
    auto put = [&](int c, bool inband, const uint4& v) {
      if constexpr (kImma) {
        put_digits<T>(Aband, dig_info, c - c_lo, inband, v);
      } else {
        if (inband) ptx::sts128(Aband + (c - c_lo) * 16, v);
      }
    };
    const uint4 zero = make_uint4(0, 0, 0, 0);
    if (kPro == PRO_RMSNORM) {
      // every CTA needs sum(x^2) over the WHOLE row (8 KB from L2 at K=4096) but normalises only its
      // band; the chunks a thread loads for the sum stay in registers and are the ones it normalises
      constexpr int U = 4;
      const bool single = nchunk <= U * CW * 32;
      // kImma, several bands per tile: the digit conversion costs ~150 instructions per chunk, so the band's chunks
      // are dealt evenly to the 128 threads in a second pass (the raw chunks cross through shared memory) instead of
      // staying with whichever thread loaded them for the sum (at Z = 8 that is 1 thread in 8)
      const bool deal = kImma && single && Z > 1;
      const uint32_t raw = Aband + p.band_units * (DIG_STAGE + DIG_INFO);
      const T* nw = static_cast<const T*>(p.norm_w);
      uint4 xr[U], wr[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {  // the norm weight does not depend on the previous kernel
        const int c = deal ? c_lo + tid + CW * 32 * i : tid + CW * 32 * i;
        wr[i] = (single && c >= c_lo && c < c_hi && c < nchunk) ? ldnc128(nw + c * 8) : zero;
      }
      wait_inputs<kHand>(p);
      if (threadIdx.x == 0) stamp(p, 2);
      float ss = 0.f;
      for (int cb = 0; cb < nchunk; cb += U * CW * 32) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int c = cb + tid + CW * 32 * i;
          xr[i] = c < nchunk ? ldcg128(A + c * 8) : zero;
        }
#pragma unroll
        for (int i = 0; i < U; ++i) ss += sumsq8<T>(xr[i]);
      }
      if (threadIdx.x == 0 && ss >= 0.f) stamp(p, 6);     // the row has arrived from L2 (the stamp waits for the sum)
      if (deal) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int c = tid + CW * 32 * i;
          if (c >= c_lo && c < c_hi) ptx::sts128(raw + (c - c_lo) * 16, xr[i]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if (lane == 0) red[warp] = ss;
      ptx::named_bar_sync(1, CW * 32);
      float tot_ss = 0.f;
#pragma unroll
      for (int w = 0; w < CW; ++w) tot_ss += red[w];
      const float rstd = rsqrtf(tot_ss / static_cast<float>(p.K) + p.eps);
      if (deal) {
        // (a rolled loop on purpose: this code runs once per launch from a cold instruction cache, the second pass
        //  should find the first one's instructions there)
#pragma unroll 1
        for (int i = 0; c_lo + CW * 32 * i < c_hi; ++i) {   // the band is at most the row: <= U passes
          const int c = c_lo + tid + CW * 32 * i;
          const bool in = c < c_hi, ld = in && c < nchunk;
          const uint4 x = ld ? ptx::lds128(raw + (c - c_lo) * 16) : zero;
          uint4 w = wr[0];
#pragma unroll
          for (int j = 1; j < U; ++j)
            if (i == j) w = wr[j];
          put(c, in, ld ? rmsnorm8<T>(x, w, rstd) : zero);
        }
      } else if (single) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
          const int c = tid + CW * 32 * i;
          put(c, c >= c_lo && c < c_hi, c < nchunk ? rmsnorm8<T>(xr[i], wr[i], rstd) : zero);
        }
      } else {
        for (int cb = c_lo; cb < c_hi; cb += CW * 32) {
          const int c = cb + tid;
          const bool ld = c < c_hi && c < nchunk;
          put(c, c < c_hi, ld ? rmsnorm8<T>(ldcg128(A + c * 8), ldnc128(nw + c * 8), rstd) : zero);
        }
      }
    } else {
      wait_inputs<kHand>(p);
      if (threadIdx.x == 0) stamp(p, 2);
      for (int cb = c_lo; cb < c_hi; cb += CW * 32) {
        const int c = cb + tid;
        const bool ld = c < c_hi && c < nchunk;
        uint4 v = ld ? ldcg128(A + c * 8) : zero;
        if (kPro == PRO_SILU_GATE && ld) v = silu_gate8<T>(v, ldcg128(A + p.K + c * 8));
        put(c, c < c_hi, v);
      }
    }
    if (Z == 1 || z == 0) pre = preload_column<T>(p, tile * BN + tid);
    ptx::named_bar_sync(1, CW * 32);
    if (threadIdx.x == 0) stamp(p, 7);       // activation band staged
  } else {
    ptx::pdl_wait_prior_grid();
    if (threadIdx.x == 0) stamp(p, 2);
  }
  const int g = lane >> 2, tig = lane & 3;
  const uint32_t rt_zero = static_cast<uint32_t>(p.K) >> 31;   // 0 at run time, opaque to the compiler
  constexpr int NT = kM1 ? 2 : 4;  // accumulator registers kept per MMA tile
  float tot[8][NT];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < NT; ++i) tot[j][i] = 0.f;

  int slot = 0, phase = 0;
  constexpr bool kLdsm = kM1 && kTrick && !kImma;   // fp16, one token: unpack through ldmatrix, see below
  if constexpr (kImma) {
    // One token on the INTEGER tensor pipe.  The packed bytes are the A operand almost as they
    // are: ldmatrix.m16n16.trans.b8 hands lane (g, tig) the bytes of packed rows 4 tig .. 4 tig + 3 of columns
    // 16 jj + g and 16 jj + g + 8, i.e. one 32-bit A-fragment register of IMMA.16832 per output column, and
    // `word & 0x0F0F0F0F` / `word & 0xF0F0F0F0` are its even-k nibbles q and its odd-k nibbles as 16 q: TWO
    // integer-pipe instructions per 8 weights (the f16 path needs five) and one MMA per 512 (f16: 256).
    // The activation enters as four signed base-256 digits per element (put_digits) in MMA columns 0-3 / 4-7:
    // the MMA of an even jj carries them in columns 0-3, the MMA of jj + 1 in columns 4-7 and accumulates onto the
    // first one's result, so all 32 lanes own live accumulators (lane tig: digits 2 (tig & 1), + 1 of column block
    // jj = 2 c + (tig >> 1)).  The accumulator starts from kMagicI - 8 (sum_even d + 16 sum_odd d), the -8 offset
    // of the nibbles, produced by one extra MMA against constant (-8, -128) rows: the s32 result IS the fp32
    // number 1.5 2^23 + sum, one FADD away from the exact integer group sum -- no I2F on the quarter-rate pipe.
    const int P = tig >> 1;
    const float wa = (tig & 1) ? 65536.f : 1.f, wb = (tig & 1) ? 16777216.f : 256.f;   // weights of this lane's digits
    const uint32_t m0 = g < 4 ? 0xFFFFFFFFu : 0u;                  // this lane's B column is a digit of block P = 0 / 1
    const int lrow = 16 * warp + (lane & 15);
    uint32_t ld_off[4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      ld_off[c] = static_cast<uint32_t>(lrow * BN + (((2 * c + (lane >> 4)) ^ (lrow & 7)) << 4));
    // scale rows through one ldmatrix.x4.trans.b16: matrix c, row r <-> 16-byte chunk 4 c + 2 (r >> 2) + (r & 1)
    const uint32_t sc_off = static_cast<uint32_t>(warp * (BN * 2) + (4 * (lane >> 3) + 2 * ((lane & 7) >> 2) + (lane & 1)) * 16);
    const uint32_t dig_off = Aband + warp * 128 + (g & 3) * 32 + tig * 8;
    const uint32_t info_off = Aband + p.band_units * DIG_STAGE + warp * 4;
    float ti[4][2];
#pragma unroll
    for (int c = 0; c < 4; ++c) ti[c][0] = ti[c][1] = 0.f;
    for (int it = 0; it < n_units; ++it) {
      const uint2 bv = ptx::lds64(dig_off + it * DIG_STAGE);
      const float inv = __uint_as_float(ptx::lds32(info_off + it * DIG_INFO));
      const uint32_t neg[4] = {0xF8F8F8F8u, 0xF8F8F8F8u, 0x80808080u, 0x80808080u};   // -8 (even k), -128 (odd k / 16)
      const int magic[4] = {kMagicI, kMagicI, kMagicI, kMagicI};
      int off[4];
      ptx::imma_s8s8(off, neg, bv.x, bv.y, magic);
      const uint32_t b0a = bv.x & m0, b1a = bv.y & m0, b0b = bv.x & ~m0, b1b = bv.y & ~m0;
      ptx::mbar_wait(&full[slot], phase);
      if (it == 0 && threadIdx.x == 0) stamp(p, 3);
      const uint32_t wrow = Wsm + slot * W_BYTES;
      int d[4][4];
      uint32_t dep = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[4];
        ptx::ldsm_x2_trans_b8(wrow + ld_off[c], r);
        dep |= r[0];
        const uint32_t a0[4] = {r[0] & 0x0F0F0F0Fu, r[1] & 0x0F0F0F0Fu, r[0] & 0xF0F0F0F0u, r[1] & 0xF0F0F0F0u};
        const uint32_t a1[4] = {r[2] & 0x0F0F0F0Fu, r[3] & 0x0F0F0F0Fu, r[2] & 0xF0F0F0F0u, r[3] & 0xF0F0F0F0u};
        ptx::imma_u8s8(d[c], a0, b0a, b1a, off);
        ptx::imma_u8s8(d[c], a1, b0b, b1b, d[c]);
      }
      uint32_t sw[4];
      ptx::ldsm_x4_trans_b16(Ssm + slot * S_BYTES + sc_off, sw);
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_after_loads(&empty[slot], dep | sw[3], rt_zero);
      const float fa = inv * wa, fb = inv * wb;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        union {
          uint32_t u;
          T h[2];
        } cv;
        cv.u = sw[c];
        const float e0 = __int_as_float(d[c][0]) - kMagicF, e1 = __int_as_float(d[c][1]) - kMagicF;
        const float e2 = __int_as_float(d[c][2]) - kMagicF, e3 = __int_as_float(d[c][3]) - kMagicF;
        ti[c][0] = fmaf(DT<T>::to_f(cv.h[0]), fmaf(fa, e0, fb * e1), ti[c][0]);   // column 16 (2 c + P) + g
        ti[c][1] = fmaf(DT<T>::to_f(cv.h[1]), fmaf(fa, e2, fb * e3), ti[c][1]);   // column 16 (2 c + P) + g + 8
      }
      if (++slot == S) {
        slot = 0;
        phase ^= 1;
      }
    }
    // the two digit pairs of a column sit in lanes tig and tig ^ 1
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float v = ti[c][h] + __shfl_xor_sync(0xffffffffu, ti[c][h], 1);
        if ((tig & 1) == 0) red[warp * BN + 16 * (2 * c + P) + g + 8 * h] = v;
      }
  } else if constexpr (kLdsm) {
    // The integer-ALU pipe (16 lanes per sub-partition) is this kernel's scarcest resource, so
    // the byte transposition is left to the load unit: ldmatrix.trans on 8x8 tiles of 16-bit elements gives
    // lane (g, tig) the word [B(2t,2g), B(2t,2g+1), B(2t+1,2g), B(2t+1,2g+1)] (B = packed byte, rows 2t, 2t+1
    // of the tile, columns 2g, 2g+1 of a 16-column block).  With z = word >> 8 the four A-fragment registers
    // of an m16n8k16 are word & 0x000F000F, z & 0x000F000F (low nibbles of the two columns, q * 2^-24 as fp16
    // subnormals) and word & 0x00F000F0, z & 0x00F000F0 (high nibbles, q * 2^-20, the odd-k activations carry
    // the 2^-4): one shift and four masks per 8 nibbles where the PRMT path needs two permutes and four masks.
    // (A nibble may not sit above bit 7 of its 16-bit lane: the bit pattern n is only the number n * 2^-24
    // while n < 2^11.)  MMA row g <-> column 16 j + 2 g, row g + 8 <-> column 16 j + 2 g + 1.
    const bool has_tok = g == 0;
    const int li = lane & 7, lm = lane >> 3;           // ldmatrix: this lane addresses row li of matrix lm
    // matrices of one ldmatrix.x4: (row tile lm & 1, 16-byte column chunk c0 + (lm >> 1))
    uint32_t ld_off[4];   // loop-invariant: row address + swizzled 16-byte chunk of this lane's four loads
#pragma unroll
    for (int c = 0; c < 4; ++c)
      ld_off[c] = static_cast<uint32_t>((16 * warp + 8 * (lm & 1) + li) * BN + (((2 * c + (lm >> 1)) ^ li) << 4));
    for (int it = 0; it < n_units; ++it) {
      ptx::mbar_wait(&full[slot], phase);
      if (it == 0 && threadIdx.x == 0) stamp(p, 3);
      const uint32_t wrow = Wsm + slot * W_BYTES;
      const uint32_t srow = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 4;
      uint32_t w[4][4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t addr = wrow + ld_off[c];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(w[c][0]), "=r"(w[c][1]), "=r"(w[c][2]), "=r"(w[c][3])
                     : "r"(addr));
      }
      float grp[8][4];
      float ag[4];
      const float zero4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        uint2 av = make_uint2(0u, 0u);
        if (has_tok) av = ptx::lds64(Aband + (it * KSTAGE + 32 * warp + 4 * tig) * 2 + 32 * b);
        const uint32_t b0 = __byte_perm(av.x, av.y, 0x5410);   // (a[+0], a[+2]) <-> low nibbles of rows 2t, 2t+1
        const uint32_t b1 = h2_mul(__byte_perm(av.x, av.y, 0x7632), 0x2C002C00u);   // (a[+1], a[+3]) * 2^-4 <-> high nibbles
        const uint32_t ones[4] = {0x3C003C00u, 0x3C003C00u, 0x4C004C00u, 0x4C004C00u};  // 1,1 | 16,16
        if (b == 0)
          ptx::mma_16816(ag, ones, b0, b1, zero4, T());
        else
          ptx::mma_16816(ag, ones, b0, b1, ag, T());
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t x = w[j >> 1][(j & 1) * 2 + b];
          const uint32_t z = x >> 8;
          const uint32_t a[4] = {x & 0x000F000Fu, z & 0x000F000Fu, x & 0x00F000F0u, z & 0x00F000F0u};
          if (b == 0)
            ptx::mma_16816(grp[j], a, b0, b1, zero4, T());
          else
            ptx::mma_16816(grp[j], a, b0, b1, grp[j], T());
        }
      }
      uint32_t sw[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sw[j] = ptx::lds32(srow + j * 32);   // scales of columns 16 j + 2 g, + 1
      __syncwarp();
      if (lane == 0)
        ptx::mbar_arrive_after_loads(&empty[slot], w[0][0] | w[1][0] | w[2][0] | w[3][0] | sw[7], rt_zero);
      const float c0 = -8.f * ag[0];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        union {
          uint32_t u;
          T h[2];
        } cv;
        cv.u = sw[j];
        const float t0 = fmaf(grp[j][0], 16777216.f, c0);   // column 16 j + 2 g
        const float t2 = fmaf(grp[j][2], 16777216.f, c0);   // column 16 j + 2 g + 1
        tot[j][0] = fmaf(DT<T>::to_f(cv.h[0]), t0, tot[j][0]);
        tot[j][1] = fmaf(DT<T>::to_f(cv.h[1]), t2, tot[j][1]);
      }
      if (++slot == S) {
        slot = 0;
        phase ^= 1;
      }
    }
  } else {
  // B fragment: lane (g, tig) supplies token g's activations a[k0+4t .. k0+4t+3] of both halves of the
  // warp's 32-k group, read from L2 one stage ahead of use (zeros for g >= M and for k >= K)
  const bool has_tok = kM1 ? (g == 0) : (g < p.M);
  const T* arow = A + static_cast<int64_t>(has_tok ? g : 0) * p.lda + 32 * warp + 4 * tig;
  auto load_a = [&](int it, uint2 (&dst)[2]) {
    const int k0 = (u0 + it) * KSTAGE;
    const bool ok = has_tok && it < n_units && k0 + 32 * warp < p.K;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      dst[b] = make_uint2(0u, 0u);
      if (ok)
        asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];"
                     : "=r"(dst[b].x), "=r"(dst[b].y)
                     : "l"(arow + k0 + 16 * b));
    }
  };
  uint2 a_cur[2], a_nxt[2];
  if (!kM1) load_a(0, a_cur);
  for (int it = 0; it < n_units; ++it) {
    if (!kM1) load_a(it + 1, a_nxt);
    ptx::mbar_wait(&full[slot], phase);
    if (it == 0 && threadIdx.x == 0) stamp(p, 3);
    const uint32_t wrow = Wsm + slot * W_BYTES + (16 * warp) * BN;
    const uint32_t srow = Ssm + slot * S_BYTES + warp * (BN * 2) + g * 32;

    float grp[8][4];
    float ag[4];
    uint32_t w_dep = 0;
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
      w_dep = q.x | pp.x;
      uint2 av = make_uint2(0u, 0u);                  // a[k0+4t .. k0+4t+3]
      if (kM1) {
        if (has_tok) av = ptx::lds64(Aband + (it * KSTAGE + 32 * warp + 4 * tig) * 2 + 32 * b);
      } else {
        av = a_cur[b];
      }
      uint32_t b0 = __byte_perm(av.x, av.y, 0x5410);   // (a[+0], a[+2]) <-> low nibbles of rows r, r+1
      uint32_t b1 = __byte_perm(av.x, av.y, 0x7632);   // (a[+1], a[+3]) <-> high nibbles
      if (kTrick) b1 = h2_mul(b1, 0x2C002C00u);        // * 2^-4: high nibble enters as q * 2^-20
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
    // release the slot only once every load from it has returned (ptx::mbar_arrive_after_loads)
    if (lane == 0) {
      if (p.plain_release)
        ptx::mbar_arrive(&empty[slot]);
      else
        ptx::mbar_arrive_after_loads(&empty[slot], sv0.x | sv1.x | w_dep, rt_zero);
    }
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
      if (kM1) {   // one token: only MMA column 0 carries data
        const float t0 = kTrick ? fmaf(grp[j][0], 16777216.f, c0) : grp[j][0];
        const float t2 = kTrick ? fmaf(grp[j][2], 16777216.f, c0) : grp[j][2];
        tot[j][0] = fmaf(sa, t0, tot[j][0]);
        tot[j][1] = fmaf(sb, t2, tot[j][1]);
      } else {
        const float t0 = kTrick ? fmaf(grp[j][0], 16777216.f, c0) : grp[j][0];
        const float t1 = kTrick ? fmaf(grp[j][1], 16777216.f, c1) : grp[j][1];
        const float t2 = kTrick ? fmaf(grp[j][2], 16777216.f, c0) : grp[j][2];
        const float t3 = kTrick ? fmaf(grp[j][3], 16777216.f, c1) : grp[j][3];
        constexpr int I2 = kM1 ? 0 : 2, I3 = kM1 ? 0 : 3;
        tot[j][0] = fmaf(sa, t0, tot[j][0]);
        tot[j][1] = fmaf(sa, t1, tot[j][1]);
        tot[j][I2] = fmaf(sb, t2, tot[j][I2]);
        tot[j][I3] = fmaf(sb, t3, tot[j][I3]);
      }
    }
    if (++slot == S) {
      slot = 0;
      phase ^= 1;
    }
    if (!kM1) {
      a_cur[0] = a_nxt[0];
      a_cur[1] = a_nxt[1];
    }
  }
  }
  if (threadIdx.x == 0) stamp(p, 4);

  // ---------------- band sum of this CTA: cross-warp (k-group) reduction through shared memory
#pragma unroll
  for (int j = 0; j < (kImma ? 0 : 8); ++j) {
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      const int tok = kM1 ? 0 : 2 * tig + (i & 1);
      const int col = kLdsm ? 16 * j + 2 * g + i : 16 * g + 2 * j + (kM1 ? i : (i >> 1));
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
        if constexpr (kM1) {
          store_column<T>(p, v[0], n, step_no, pre);
        } else {
          T* Cp = static_cast<T*>(p.C);
          const T* bias = static_cast<const T*>(p.bias);
#pragma unroll
          for (int m = 0; m < MR; ++m)
            if (m < p.M) Cp[m * p.ldc + n] = epilogue<T>(v[m], bias, n);
        }
      }
      signal_tile<kHand>(p);
    } else {
      // The band sums meet in rank 0's shared memory (DSMEM).  Ranks 1 .. Z-1 push theirs with st.async, which
      // completes the transaction count of rank 0's mbarrier: the data is its own arrival signal, no cluster barrier
      // at the end of a 4 us launch, and the pushing CTA exits right away.  Rank 0 adds in rank order.
      ptx::cluster_wait_acquire();               // every CTA of the cluster has started (phase A)
      if (z != 0) {
        const uint32_t local = ptx::smem_u32(xred) + static_cast<uint32_t>((z * MR) * BN + t) * 4u;
        const uint32_t remote = ptx::mapa_rank(local, 0), rbar = ptx::mapa_rank(ptx::smem_u32(xbar), 0);
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < p.M) ptx::st_async_cluster_f32(remote + m * BN * 4, v[m], rbar);
      } else {
        ptx::mbar_wait(xbar, 0);
        const int n = tile * BN + t;
        if (n < p.N) {
          T* Cp = static_cast<T*>(p.C);
          const T* bias = static_cast<const T*>(p.bias);
#pragma unroll
          for (int m = 0; m < MR; ++m) {
            if (m < p.M) {
              float acc = v[m];
              for (int zz = 1; zz < Z; ++zz) acc += xred[(zz * MR + m) * BN + t];
              if constexpr (kM1)
                store_column<T>(p, acc, n, step_no, pre);
              else
                Cp[m * p.ldc + n] = epilogue<T>(acc, bias, n);
            }
          }
        }
        signal_tile<kHand>(p);
      }
    }
  }
  }  // consumers
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

// (k-stages per tile, k-bands per tile == cluster size) of a decode launch
void plan_bands(int N, int K, int* SPT_out, int* Z_out) {
  static const int z_env = env_int("CGQ_GEMV_Z", 0, 0, 8);
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 4, 1, 4);
  const int SPT = (K / 32 + CW - 1) / CW;
  const int tiles = (N + BN - 1) / BN;
  const int slots = cps * sm_count();
  // fill the CTA slots of the SMs in one wave, powers of two up to the portable cluster size, never more bands than
  // k-stages -- but once the launch has ~1.4 CTAs per SM, do not cut the bands below 16 k-stages: every CTA pays its
  // own prologue and cluster reduction (K = 4096, N = 13696: 214 CTAs x 16 stages 7.7 us, 428 x 8 stages 10.0 us)
  int Z = 1;
  while (Z < 8 && tiles * (Z * 2) <= slots && SPT >= Z * 2 && (SPT / (Z * 2) >= 16 || tiles * Z * 5 < sm_count() * 7))
    Z *= 2;
  if (z_env > 0) Z = z_env;
  if (Z > SPT) Z = 1;
  *SPT_out = SPT;
  *Z_out = Z;
}

struct NextHint {
  const void* w;
  const void* s;
  int N, K;
};
thread_local NextHint g_next = {nullptr, nullptr, 0, 0};
struct Handover {
  const unsigned* wait_ctr;
  unsigned wait_count;
  unsigned* signal_ctr;
};
thread_local Handover g_hand = {nullptr, 0, nullptr};
struct TpHint {
  cgq_tp_ctx ctx;
  unsigned idx;
  bool set;
};
thread_local TpHint g_tp = {{}, 0, false};

template <typename T, bool kTrick, bool kM1, int kPro, bool kHand = false, bool kImma = false>
int launch_inst(const GemmArgs& a, const CUtensorMap& tmW, const CUtensorMap& tmS, Params prm,
                int grid, int stages, bool pdl) {
  using C = Cfg<kM1>;
  static_assert(!kImma || kM1, "the integer-MMA arithmetic is the one-token kernel's");
  // (+ the raw chunks of the band while the RMSNorm prologue deals them to the threads, Z > 1 only)
  const size_t band = kImma ? static_cast<size_t>(prm.band_units) *
                                  (DIG_STAGE + DIG_INFO + (kPro == PRO_RMSNORM && prm.Z > 1 ? KSTAGE * 2 : 0))
                            : (kM1 ? static_cast<size_t>(prm.band_units) * KSTAGE * 2 : 0);
  const size_t smem = 1024 + static_cast<size_t>(stages) * C::STAGE_BYTES + C::RED_BYTES +
                      C::xred_bytes(prm.Z) + 16 * stages + 48 + band;
  auto kern = w4_gemv_kernel<T, kTrick, kM1, kPro, kHand, kImma>;
  static size_t configured[64] = {0};
  int dev = 0;
  CGQ_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && smem > configured[dev]) {
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(smem)));
    CGQ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
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

template <typename T, bool kTrick>
int launch_m1(const GemmArgs& a, const CUtensorMap& tmW, const CUtensorMap& tmS, const Params& prm,
              int grid, int stages, bool pdl, int pro, bool imma) {
  if (imma && prm.wait_ctr == nullptr && prm.signal_ctr == nullptr) {   // integer-MMA arithmetic (the default)
    switch (pro) {
      case PRO_RMSNORM:
        return launch_inst<T, false, true, PRO_RMSNORM, false, true>(a, tmW, tmS, prm, grid, stages, pdl);
      case PRO_SILU_GATE:
        return launch_inst<T, false, true, PRO_SILU_GATE, false, true>(a, tmW, tmS, prm, grid, stages, pdl);
      default:
        return launch_inst<T, false, true, PRO_NONE, false, true>(a, tmW, tmS, prm, grid, stages, pdl);
    }
  }
  if (prm.wait_ctr != nullptr || prm.signal_ctr != nullptr) {   // tile-granular hand-over variant
    switch (pro) {
      case PRO_RMSNORM:
        return launch_inst<T, kTrick, true, PRO_RMSNORM, true>(a, tmW, tmS, prm, grid, stages, pdl);
      case PRO_SILU_GATE:
        return launch_inst<T, kTrick, true, PRO_SILU_GATE, true>(a, tmW, tmS, prm, grid, stages, pdl);
      default:
        return launch_inst<T, kTrick, true, PRO_NONE, true>(a, tmW, tmS, prm, grid, stages, pdl);
    }
  }
  switch (pro) {
    case PRO_RMSNORM:
      return launch_inst<T, kTrick, true, PRO_RMSNORM>(a, tmW, tmS, prm, grid, stages, pdl);
    case PRO_SILU_GATE:
      return launch_inst<T, kTrick, true, PRO_SILU_GATE>(a, tmW, tmS, prm, grid, stages, pdl);
    default:
      return launch_inst<T, kTrick, true, PRO_NONE>(a, tmW, tmS, prm, grid, stages, pdl);
  }
}

// arith: W4_ARITH_* (common.cuh)
template <typename T>
int launch_t(const GemmArgs& a, int arith, const GemvFused* fu) {
  if (arith == W4_ARITH_DEFAULT) arith = default_w4_arith(W4_ARITH_DEFAULT);
  const bool exact = arith == W4_ARITH_EXACT;
  const int G = a.K / 32;
  const int tiles = (a.N + BN - 1) / BN;
  static const int stages_env = env_int("CGQ_GEMV_STAGES", 0, 0, 16);
  static const bool pdl = env_int("CGQ_PDL", 1, 0, 1) != 0;
  static const int cps = env_int("CGQ_GEMV_CTAS_PER_SM", 4, 1, 4);
  const int slots = cps * sm_count();
  int SPT, Z;
  plan_bands(a.N, a.K, &SPT, &Z);
  const int grid = tiles * Z;
  // fewer CTAs than slots -> deeper rings keep the same number of bytes in flight
  // Ring depth by how many CTAs of the launch share an SM: 4 -> 4 stages, 3 -> 6, <= 2 -> 8 (7 for the fused
  // launches, whose CTAs carry the prologue buffers).  Measured on the token chain / the fused step (B200): the depth
  // of w_out's ring (256 CTAs, 14 stages each) decides how early the neighbouring launches' CTAs fit beside it --
  // chain 787 -> 744 us at 8, step 979 -> 953 us at 7; 9 and more lose again (profiles/r02_ring_depth.txt).
  static const int s_small = env_int("CGQ_GEMV_STAGES_SMALL", 6, 2, 16);     // <= 3 CTAs per SM
  static const int s_two = env_int("CGQ_GEMV_STAGES_2PERSM", 0, 0, 16);      // <= 2 CTAs per SM (0: 8 plain, 7 fused)
  int stages = stages_env > 0 ? stages_env : (grid * 4 <= slots * 3 ? s_small : 4);
  // (one token only: the M > 1 kernel keeps 2 CTAs per SM and loses with deeper rings -- bs-8 chain 1.80 -> 2.17 ms)
  if (stages_env == 0 && grid * 2 <= slots && a.M == 1) stages = s_two > 0 ? s_two : (fu != nullptr ? 7 : 8);
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
  prm.resid = fu != nullptr ? fu->resid : nullptr;
  prm.norm_w = fu != nullptr ? fu->norm_w : nullptr;
  prm.eps = fu != nullptr ? fu->eps : 0.f;
  prm.band_units = per_cta;
  static const int plain_release = env_int("CGQ_HACK_PLAIN_RELEASE", 0, 0, 1);
  prm.plain_release = plain_release;
  // one-shot tensor-parallel hint (cgq_tp_next); only the fused M == 1 launches take it
  memset(&prm.tp, 0, sizeof(prm.tp));
  prm.tp_idx = 0;
  if (g_tp.set) {
    if (fu != nullptr && a.M == 1) {
      prm.tp = g_tp.ctx;
      prm.tp_idx = g_tp.idx;
    }
    g_tp.set = false;
  }
  // one-shot hand-over hint (cgq_handover_next); only the fused M == 1 launches take it
  const Handover hand = g_hand;
  g_hand = Handover{nullptr, 0, nullptr};
  const bool use_hand = fu != nullptr && a.M == 1;
  prm.wait_ctr = use_hand ? hand.wait_ctr : nullptr;
  prm.wait_count = use_hand ? hand.wait_count : 0;
  prm.signal_ctr = use_hand ? hand.signal_ctr : nullptr;
  static const int hand_mode = env_int("CGQ_HAND_MODE", 0, 0, 3), hand_sleep = env_int("CGQ_HAND_SLEEP", 20, 0, 2000);
  prm.hand_mode = static_cast<unsigned>(hand_mode);
  prm.hand_sleep = static_cast<unsigned>(hand_sleep);
  const int pro = fu != nullptr ? fu->prologue : PRO_NONE;
  // one-shot hint: stream the next launch's weights into L2 from this kernel's producers
  const NextHint nh = g_next;
  g_next = NextHint{nullptr, nullptr, 0, 0};
  prm.pf_pieces = 0;
  static const int pf_mb = env_int("CGQ_PF_MB", 0, 0, 512);   // measured slower on B200: off unless asked for
  if (nh.w != nullptr && pf_mb > 0) {
    int spt_n, z_n;
    plan_bands(nh.N, nh.K, &spt_n, &z_n);
    const int per_n = (spt_n + z_n - 1) / z_n;
    const int64_t stage_bytes = static_cast<int64_t>(ROWS + CW * 2) * nh.N;   // weights + scales of one k-stage
    int64_t depth = (static_cast<int64_t>(pf_mb) << 20) / (stage_bytes * z_n);
    if (depth > per_n) depth = per_n;
    if (depth > 0) {
      prm.pf_w = static_cast<const uint8_t*>(nh.w);
      prm.pf_s = static_cast<const uint8_t*>(nh.s);
      prm.pf_N = nh.N;
      prm.pf_rows = nh.K / 2;
      prm.pf_groups = nh.K / 32;
      prm.pf_SPT = spt_n;
      prm.pf_Z = z_n;
      prm.pf_depth = static_cast<int>(depth);
      prm.pf_ppc_w = (ROWS * nh.N + kPfPiece - 1) / kPfPiece;
      prm.pf_ppc = prm.pf_ppc_w + (CW * 2 * nh.N + kPfPiece - 1) / kPfPiece;
      prm.pf_pieces = static_cast<int>(depth) * z_n * prm.pf_ppc;
    }
  }

  constexpr bool kIsHalf = (DT<T>::code == CGQ_DTYPE_F16);
  const bool trick = kIsHalf && !exact;
  if (a.M == 1) {
    const bool imma = arith == W4_ARITH_IMMA;
    if (trick && !imma) return launch_m1<T, kIsHalf>(a, tmW, tmS, prm, grid, stages, pdl, pro, false);
    return launch_m1<T, false>(a, tmW, tmS, prm, grid, stages, pdl, pro, imma);
  }
  // M > 1 in fp16 takes the subnormal-operand arithmetic too (8 % faster M = 8 chain).  It was switched off in round
  // 1 after run-to-run divergence at M >= 5; that was the ring-release race (mbarrier.arrive overtaking the stage's
  // outstanding ld.shared), which the faster variant simply hit more often: with the load-dependent release it is
  // bit-stable over the 200-launch stress under concurrent L2 traffic, and CGQ_HACK_PLAIN_RELEASE=1 brings the
  // divergence back (profiles/r02_rootcause_*.txt).  CGQ_GEMV_TRICK_MGT1=0 selects the exact path.
  static const bool trick_mgt1 = env_int("CGQ_GEMV_TRICK_MGT1", 1, 0, 1) != 0;
  if (trick && trick_mgt1) return launch_inst<T, kIsHalf, false, PRO_NONE>(a, tmW, tmS, prm, grid, stages, pdl);
  return launch_inst<T, false, false, PRO_NONE>(a, tmW, tmS, prm, grid, stages, pdl);
}

}  // namespace

bool w4_gemv_supported(const GemmArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return a.M >= 1 && a.M <= MMAX && a.K % 32 == 0 && a.N % 16 == 0 && al16(a.Wq) &&
         al16(a.scale) && al16(a.A) && a.lda % 8 == 0 && true;
}

int launch_w4_gemv(const GemmArgs& a, int arith) {
  static const bool umma_default = env_int("CGQ_GEMV_UMMA", 0, 0, 1) != 0;
  if (umma_default && arith == W4_ARITH_DEFAULT && a.M == 1) {   // opt-in: integer tcgen05 decode kernel (gemv_w4_umma.cu)
    bool taken = false;
    const int rc = launch_w4_gemv_umma(a, &taken);
    if (rc != CGQ_OK || taken) return rc;
  }
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, arith, nullptr)
                                  : launch_t<__nv_bfloat16>(a, arith, nullptr);
}

// process-wide default arithmetic of the decode kernel: CGQ_GEMV_ARITH (0 imma, 1 exact, 2 subnormal) unless set
// through cgq_set_decode_arith; returns the value in force before the call (set < 0: query only)
int default_w4_arith(int set) {
  static std::atomic<int> cur{env_int("CGQ_GEMV_ARITH", W4_ARITH_IMMA, W4_ARITH_IMMA, W4_ARITH_SUBNORMAL)};
  if (set < W4_ARITH_IMMA || set > W4_ARITH_SUBNORMAL) return cur.load();
  return cur.exchange(set);
}

void set_next_w4_hint(const void* w, const void* s, int N, int K) {
  g_next = NextHint{w, s, N, K};
}

void set_w4_tp(const cgq_tp_ctx& ctx, unsigned idx) {
  g_tp.ctx = ctx;
  g_tp.idx = idx;
  g_tp.set = true;
}

void set_w4_handover(const unsigned* wait_ctr, unsigned wait_count, unsigned* signal_ctr) {
  g_hand = Handover{wait_ctr, wait_count, signal_ctr};
}
int w4_gemv_tiles(int N) { return (N + BN - 1) / BN; }

// M == 1 with a fused prologue (RMSNorm / SiLU-gate on the activation) and residual epilogue.
int launch_w4_gemv_fused(const GemmArgs& a, const GemvFused& fu) {
  return a.dtype == CGQ_DTYPE_F16 ? launch_t<__half>(a, W4_ARITH_DEFAULT, &fu)
                                  : launch_t<__nv_bfloat16>(a, W4_ARITH_DEFAULT, &fu);
}

}  // namespace cgq
