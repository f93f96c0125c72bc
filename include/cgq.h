/*
 * cgq.h — C-ABI of the B200-native (sm_100a) weight-only dequant-matmul path of chatglm-q.
 *
 * The reference (K024/chatglm-q) has no FFI: its GPU path is four Triton kernels bound to
 * the module globals `_dynamic_quant_matmul_impl` / `check_input` of
 *   chatglm_q/int4/qlinear.py:7-17   and   chatglm_q/int8/qlinear.py:6-16.
 * This header DEFINES the boundary a maintainer would bind there (see INTEGRATION.md for the
 * ctypes stub).  Every entry point states which reference interface it replaces.
 *
 * Conventions
 *   - plain pointers + sizes; all data pointers are DEVICE pointers on the current CUDA device;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - launches are asynchronous on `stream`; no host synchronisation inside any call;
 *   - return value: CGQ_OK (0) or a negative CGQ_ERR_* code; `cgq_last_error()` has the text;
 *   - no CPU fallback exists: an input the kernels cannot take is an error, never a slow path.
 *
 * dtype codes (activation / scale / output element type): 0 = IEEE fp16, 1 = bfloat16.
 */
#ifndef CGQ_H_
#define CGQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGQ_OK 0
#define CGQ_ERR_BAD_SHAPE (-1)      /* M,N,K / group inconsistent (reference: AssertionError, int4/triton_ops.py:102-123) */
#define CGQ_ERR_BAD_DTYPE (-2)      /* dtype code not 0/1 (reference asserts a.dtype == b_scale.dtype, :108) */
#define CGQ_ERR_MISALIGNED (-3)     /* pointer / leading dimension alignment not met */
#define CGQ_ERR_CUDA (-4)           /* a CUDA runtime / driver call failed */
#define CGQ_ERR_WORKSPACE (-5)      /* workspace NULL or smaller than cgq_workspace_bytes() */
#define CGQ_ERR_UNSUPPORTED (-6)    /* device is not sm_100 */

#define CGQ_DTYPE_F16 0
#define CGQ_DTYPE_BF16 1

/* Kernel selection for the *_ex entry points (tests / bench / profiling only). */
#define CGQ_IMPL_AUTO 0             /* what cgq_w4a16_gemm / cgq_w8a16_gemm pick */
#define CGQ_IMPL_SIMPLE 1           /* one-thread-per-column CUDA-core kernel, bit-faithful dequant */
#define CGQ_IMPL_GEMV 2             /* TMA-fed cluster mma.sync kernel, M <= 8 (decode), default arithmetic */
#define CGQ_IMPL_GEMV_EXACT 3       /* same, (q-8) converted exactly to fp16 / bf16, f16 MMA */
#define CGQ_IMPL_TC 4               /* tcgen05 tensor-core GEMM (prefill) */
#define CGQ_IMPL_GEMV_UMMA 5        /* M == 1 decode on integer tcgen05 (int8 digits of the activation), opt-in */
#define CGQ_IMPL_GEMV_SUBNORMAL 6   /* decode kernel, fp16 nibbles as subnormal f16 MMA operands (round-1/2 default) */
#define CGQ_IMPL_GEMV_IMMA 7        /* decode kernel, M == 1: IMMA.16832 on base-256 digits of the activation (default);
                                       M > 1 takes the subnormal-operand path */

/* Library / ABI version: (major << 16) | minor. */
int cgq_version(void);

/* Text of the last error raised on the calling thread ("" if none). */
const char* cgq_last_error(void);

/*
 * Bytes of device workspace a GEMM call needs (stream-K partial tiles + self-cleaning tile
 * counters).  The workspace must be zero-filled ONCE after allocation; every call leaves it
 * zeroed again.  One workspace may not be shared by calls running concurrently on different
 * streams.  The value is a constant upper bound, independent of the shape.
 * Reference: none (the Triton kernels use no workspace, int4/triton_ops.py:124-138).
 */
size_t cgq_workspace_bytes(void);

/*
 * C[M,N] = A[M,K] · dequant(Wq, scale) (+ bias), int4 group-quantised weights.
 * Replaces chatglm_q.int4.triton_ops.dynamic_quant_matmul_s4 (int4/triton_ops.py:90-139,
 * kernel :18-87) and, with `bias`, the `out += self.bias` of
 * DynamicQuantizeLinear.forward (int4/qlinear.py:90-94).
 *
 *   A      [M, K]    dtype, row stride `lda` elements, unit column stride
 *   Wq     [K/2, N]  uint8, contiguous; byte (r, n) = nibble(k=2r) | nibble(k=2r+1) << 4,
 *                    value = nibble - 8                        (int4/quantizer.py:25-28)
 *   scale  [K/group, N] dtype, contiguous                      (int4/qlinear.py:84)
 *   bias   [N] dtype or NULL; added AFTER the product is rounded to dtype (two roundings,
 *                    as the reference's separate in-place add)
 *   C      [M, N]    dtype, row stride `ldc` elements
 *   group  must be 32 (the only group size the reference model builds, int4/qlinear.py:5,76)
 */
int cgq_w4a16_gemm(const void* A, int64_t lda, const uint8_t* Wq, const void* scale,
                   const void* bias, void* C, int64_t ldc, int M, int N, int K, int group,
                   int dtype, void* workspace, size_t workspace_bytes, void* stream);

/* Same, with an explicit kernel choice (CGQ_IMPL_*). */
int cgq_w4a16_gemm_ex(const void* A, int64_t lda, const uint8_t* Wq, const void* scale,
                      const void* bias, void* C, int64_t ldc, int M, int N, int K, int group,
                      int dtype, void* workspace, size_t workspace_bytes, void* stream, int impl);

/*
 * C[M,N] = A[M,K] · (Wq^T * scale) (+ bias), int8 per-output-channel weights.
 * Replaces chatglm_q.int8.triton_ops.dynamic_quant_matmul (int8/triton_ops.py:87-127,
 * kernel :13-84) called with `weight.t()` by DynamicQuantizeLinear.forward
 * (int8/qlinear.py:89-93).
 *
 *   Wq     [N, K] int8, contiguous (the module buffer itself, NOT the transposed view)
 *   scale  [N] dtype (may be negative: tests/test_triton_ops.py:12)
 */
int cgq_w8a16_gemm(const void* A, int64_t lda, const int8_t* Wq, const void* scale,
                   const void* bias, void* C, int64_t ldc, int M, int N, int K, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream);

int cgq_w8a16_gemm_ex(const void* A, int64_t lda, const int8_t* Wq, const void* scale,
                      const void* bias, void* C, int64_t ldc, int M, int N, int K, int dtype,
                      void* workspace, size_t workspace_bytes, void* stream, int impl);

/*
 * Unpack the int4 weight to signed integers: out_i8[K, N] = nibble - 8.  Bit-exact restatement
 * of `((x >> shifts) & 0xF).to(int8) - 8` (int4/qlinear.py:29-31).  Test-only surface that
 * carries the "int unpack bit-exact" claim.
 */
int cgq_w4_unpack_i8(const uint8_t* Wq, int8_t* out, int K, int N, void* stream);

/*
 * Dequantise the int4 weight: out[K, N] (dtype) = round_dtype((nibble - 8) * scale[k/group, n]).
 * Bit-exact restatement of chatglm_q.int4.qlinear.unpack_int4 (int4/qlinear.py:20-33).
 */
int cgq_w4_dequant(const uint8_t* Wq, const void* scale, void* out, int K, int N, int group,
                   int dtype, void* stream);

/*
 * Embedding row gather + dequant for the int4 / int8 QEmbedding modules
 * (int4/qlinear.py:122-130: packed along the VOCAB axis, 2 tokens per byte, groups of 32 tokens;
 *  int8/qlinear.py:118-120).  ids are int64 token ids, out is [n_ids, dim] dtype.
 */
int cgq_w4_embedding(const int64_t* ids, int n_ids, const uint8_t* Wq /*[V/2, D]*/,
                     const void* scale /*[V/group, D]*/, void* out, int V, int D, int group,
                     int dtype, void* stream);
int cgq_w8_embedding(const int64_t* ids, int n_ids, const int8_t* Wq /*[V, D]*/,
                     const void* scale /*[D]*/, void* out, int V, int D, int dtype, void* stream);

/*
 * ---- Fused batch-1 decode step --------------------------------------
 * What ChatGLM2Model.forward (chatglm_q/model.py:329-392) does for ONE new token against a KV
 * cache, as a chain of launches that a caller captures into one CUDA graph.  Every launch carries
 * the programmatic-dependent-launch attribute: the next dequant-matmul streams its weights while
 * the current kernel still runs.  All three calls are asynchronous and allocation-free.
 *
 * cgq_w4a16_gemv_fused: C[N] = (resid[N] +) round(prologue(A) · dequant(Wq, scale)) (+ bias), M == 1.
 *   prologue CGQ_PRO_NONE      a = A[0..K)
 *            CGQ_PRO_RMSNORM   a = round(round(A * rsqrt(mean(A^2) + eps)) * norm_w)
 *                              — RMSNorm.forward feeding the linear (model.py:62-73, 231, 244, 381)
 *            CGQ_PRO_SILU_GATE a[k] = round(round(silu(A[k])) * A[K + k]), A has 2K elements
 *                              — GatedFeedForward.forward between w_in and w_out (model.py:200-201)
 *   resid (nullable) is the residual stream: `x = x + h` (model.py:243, 246); C may alias resid.
 */
#define CGQ_PRO_NONE 0
#define CGQ_PRO_RMSNORM 1
#define CGQ_PRO_SILU_GATE 2
int cgq_w4a16_gemv_fused(const void* A, const uint8_t* Wq, const void* scale, const void* bias,
                         const void* resid, void* C, int N, int K, int group, int dtype,
                         int prologue, const void* norm_w, float eps, void* stream);

/*
 * EXPERIMENTAL one-shot hint (no reference counterpart): the NEXT int4 decode launch (M <= 8) issued by
 * the calling thread also streams the leading part (at most CGQ_PF_MB MiB) of THIS weight — the one
 * the launch after it will read — from HBM into L2 with `cp.async.bulk.prefetch.L2`, in the order
 * that launch will consume it, so that HBM keeps streaming across the dependency bubble between two
 * launches.  Measured on B200 it is a net loss (the decode kernel is issue-bound on the SM, not
 * HBM-bound, and the prefetches queue in front of its own TMA loads), so it is DISABLED
 * unless the environment sets CGQ_PF_MB > 0; the call is then a no-op.  Results are never affected.
 * Wq == NULL cancels.
 */
int cgq_prefetch_next_w4(const uint8_t* Wq, const void* scale, int N, int K);

/*
 * EXPERIMENTAL one-shot hint (no reference counterpart): tile-granular hand-over between two consecutive
 * cgq_w4a16_gemv_fused launches instead of the grid-granular `griddepcontrol.wait`.  The NEXT
 * cgq_w4a16_gemv_fused launch issued by the calling thread
 *   - `wait_ctr` != NULL: does not wait for the previous grid to drain; its consumers acquire-poll
 *     *wait_ctr until it reaches `wait_count` (= cgq_w4_gemv_tiles(N) of the launch that produces its input)
 *     before touching the activation (bounded spin: a lost producer yields wrong numbers, not a hung device);
 *   - `signal_ctr` != NULL: release-increments *signal_ctr once per stored 128-column output tile.
 * The caller zeroes the counters on the stream before the first launch that uses them (one counter per
 * producing launch per step) and keeps launches that reuse an activation buffer apart by a full dependency
 * (DESIGN.md §6.1a lists the hazards FusedDecodeModel's schedule was checked against).  Either pointer NULL
 * leaves that side on the default protocol; (NULL, 0, NULL) cancels.  Results are bit-identical.
 */
int cgq_handover_next(const uint32_t* wait_ctr, uint32_t wait_count, uint32_t* signal_ctr);
/* Output tiles (128 columns each) a decode launch with N columns stores -- the count its consumers wait for. */
int cgq_w4_gemv_tiles(int N);

/*
 * ---- Persistent decode program: a chain of batch-1 int4g32 linears in ONE launch --------------------
 * The same linears as cgq_w4a16_gemv_fused (same arithmetic, bit for bit), executed in the given order by
 * persistent workers: every worker's TMA producer walks the whole chain and keeps its shared-memory ring
 * full with the weights of whatever comes next, so HBM keeps streaming while the consumer warps wait in the
 * grid barrier between two dependent linears.  Op i may read what ops < i wrote (A, resid) -- the chain
 * is executed strictly in order.  All buffers must stay valid and in place while the program exists.
 *   cgq_program_create  builds the device-side description (tensor maps, work split) once;
 *   cgq_program_run     = one cudaMemsetAsync + one kernel launch on `stream` (graph-capturable);
 *   cgq_program_status  synchronises the device and reports the worker count and whether a grid barrier
 *                       ever timed out (a lost worker: results are then invalid);
 *   cgq_program_destroy frees it.
 */
typedef struct {
  const uint8_t* Wq;    /* [K/2, N] packed int4 (as cgq_w4a16_gemm) */
  const void* scale;    /* [K/32, N] dtype */
  const void* bias;     /* [N] dtype or NULL */
  const void* A;        /* activation row [K] dtype ([2K] for CGQ_PRO_SILU_GATE) */
  void* C;              /* output row [N] dtype */
  const void* resid;    /* [N] dtype or NULL (may alias C) */
  const void* norm_w;   /* [K] dtype, CGQ_PRO_RMSNORM only */
  int N, K;
  int prologue;         /* CGQ_PRO_* */
  float eps;
} cgq_linear_op;
int cgq_program_create(const cgq_linear_op* ops, int n_ops, int dtype, uint64_t* handle);
int cgq_program_run(uint64_t handle, void* stream);
int cgq_program_status(uint64_t handle, int* workers, int* failed);
int cgq_program_destroy(uint64_t handle);

/*
 * ---- Backward with respect to the activation ----------------------------------------------
 * grad_A[M, K] = grad_out[M, N] . dequant(W)^T -- `DynamicQuantizeMatMul.backward` (int4/qlinear.py:53-64,
 * int8/qlinear.py:41-52; Triton twins int4/triton_ops.py:142-264, int8/triton_ops.py:130-245).  Weight layouts as in
 * the forward entry points (int4: Wq [K/2, N] + scale [K/32, N]; int8: Wq [N, K] + scale [N]).  Every element is
 * dequantised with the reference's single rounding, fp32 accumulation, one final rounding.  CUDA-core kernel, any
 * shape (the reference's Triton kernel needs power-of-two sizes).
 */
int cgq_w4a16_grad_a(const void* grad_out, int64_t ldg, const uint8_t* Wq, const void* scale, void* grad_a, int64_t ldo,
                     int M, int N, int K, int group, int dtype, void* stream);
int cgq_w8a16_grad_a(const void* grad_out, int64_t ldg, const int8_t* Wq, const void* scale, void* grad_a, int64_t ldo,
                     int M, int N, int K, int dtype, void* stream);

/*
 * ---- Tensor parallelism of the fused decode step (no reference counterpart: the reference is single-GPU) ------
 * One process per GPU.  Column-parallel linears (qkv_proj, w_in, lm_head) need no exchange; a ROW-parallel linear
 * (o_proj, w_out: this rank holds a k-slice) exchanges its fp32 partial sums INSIDE the decode kernel's epilogue:
 * every output column is stored as an 8-byte {value, epoch} word into every rank's receive buffer over NVLink
 * (peer-mapped device memory, cgq_ipc_*), each rank then sums the `world` words of the column in rank order, rounds
 * once and adds the residual -- no NCCL call, no extra launch, bit-identical rows on all ranks.
 *   recv[r]   rank r's receive buffer, [2 slots][world][max_n] x 8 bytes, zero-initialised (recv[rank] = own)
 *   step      device int32 incremented once per token BEFORE the step's launches run (cgq_decode_begin_w4 does it
 *             to state[2]); epoch = step * 128 + idx + 1
 *   err       device uint32, set if a peer's word never arrived (bounded spin), or NULL
 *   out[r]    broadcast stores of a column-parallel linear (lm_head): rank r's full output buffer; this rank's N
 *             columns go to out[r][out_offset + n] for every r.  NULL = plain local store to C.
 * cgq_tp_next(ctx, idx): one-shot hint for the NEXT cgq_w4a16_gemv_fused launch of the calling thread; idx = running
 * index of the exchange within the token (consecutive exchanges alternate slots), < 127.  For a broadcast-only launch
 * pass recv[0] = NULL.  cgq_tp_barrier: a one-thread kernel that publishes `epoch` to every rank's `flags[rank]`
 * (release, system scope) and waits until all `world` flags of its own array reached it -- run once after the
 * broadcast linear so that the sampler sees every rank's logits.
 */
typedef struct {
  int world, rank;
  int max_n;
  int out_offset;
  void* recv[8];
  const int* step;
  uint32_t* err;
  void* out[8];
} cgq_tp_ctx;
int cgq_tp_next(const cgq_tp_ctx* ctx, uint32_t idx);
int cgq_tp_barrier(uint32_t* const* flags /*[world] peer-mapped, each [world] uint32*/, int world, int rank,
                   const int* step, uint32_t* err, void* stream);
/* Peer-visible device memory: cudaMalloc + legacy CUDA IPC handle (64 bytes) / open a peer's handle / close / free. */
int cgq_ipc_alloc(size_t bytes, void** ptr, void* handle64);
int cgq_ipc_open(const void* handle64, void** ptr);
int cgq_ipc_close(void* ptr);
int cgq_ipc_free(void* ptr);

/*
 * ---- One-launch decode step ("step program"): embedding row + attention + linears of a whole token -----------
 * The batch-1 token step of the int4g32 model (ChatGLM2Model.forward with past_key_values, model.py:329-392) as ONE
 * persistent cooperative kernel: one CTA per SM, the TMA producer of every CTA streams the weights of the whole
 * step through a deep shared-memory ring (HBM keeps streaming across phase boundaries), a linear is cut into
 * column slices that belong to one CTA each (no cross-CTA reduction), phases are separated by a flat grid barrier
 * (DESIGN.md §3.10).  Ops run strictly in order; op i may read what ops < i wrote.
 *   CGQ_STEP_LINEAR     out[N] = (resid +) round(prologue(A) . dequant(Wq, scale)) (+ bias)   -- as cgq_linear_op
 *   CGQ_STEP_ATTENTION  RoPE, KV append, attention of one query row: A = qkv, C = out           -- as cgq_decode_attention
 *   CGQ_STEP_EMBED      C[N] = int4 QEmbedding row ids[0] of (Wq [V/2, N], scale [V/32, N])      -- as cgq_decode_begin_w4
 * `state` ([1] int32, device; may be NULL without attention ops) = tokens in the KV cache before the step; the
 * kernel reads it at entry and increments it.  fp16 only; N a multiple of 32, K <= 13824.
 *   cgq_step_create / cgq_step_run (memset + one cooperative launch, graph-capturable) / cgq_step_status
 *   (synchronises; CTAs, ring stages, whether a barrier ever timed out) / cgq_step_destroy.
 */
#define CGQ_STEP_LINEAR 0
#define CGQ_STEP_ATTENTION 1
#define CGQ_STEP_EMBED 2
/* LINEAR epilogues.  CGQ_EPI_SILU_PAIR (w_in, model.py:200-201): the N columns are [h | gate]; the op stores
 * C[N/2] = round(round(silu(round(h))) * round(gate)) instead of the N raw outputs, so that the next linear (w_out)
 * takes it with CGQ_PRO_NONE -- the same roundings as CGQ_PRO_SILU_GATE on the raw outputs, computed once. */
#define CGQ_EPI_NONE 0
#define CGQ_EPI_SILU_PAIR 1
typedef struct {
  int kind;             /* CGQ_STEP_* */
  const uint8_t* Wq;    /* LINEAR: [K/2, N]; EMBED: [V/2, N] */
  const void* scale;    /* LINEAR: [K/32, N]; EMBED: [V/32, N] */
  const void* bias;     /* LINEAR: [N] or NULL */
  const void* A;        /* LINEAR: activation row [K] ([2K] for CGQ_PRO_SILU_GATE); ATTENTION: qkv */
  void* C;              /* output row */
  const void* resid;    /* LINEAR: [N] or NULL (may alias C) */
  const void* norm_w;   /* LINEAR: [K], CGQ_PRO_RMSNORM only */
  int N, K;             /* EMBED: N = embedding dim */
  int prologue;         /* CGQ_PRO_* */
  float eps;
  int epilogue;         /* CGQ_EPI_* */
  const void* freqs;    /* ATTENTION: rotary table [max_pos, d_head] */
  void* kcache;         /* ATTENTION: [max_len, n_groups, d_head] */
  void* vcache;
  int n_head, n_groups, d_head, max_len;
  const int64_t* ids;   /* EMBED: [1] token id (device) */
  int V;                /* EMBED: vocabulary size */
} cgq_step_op;
int cgq_step_create(const cgq_step_op* ops, int n_ops, int dtype, int* state, uint64_t* handle);
int cgq_step_run(uint64_t handle, void* stream);
int cgq_step_status(uint64_t handle, int* ctas, int* stages, int* failed);
int cgq_step_destroy(uint64_t handle);

/*
 * First launch of a decode step: x[D] = int4 QEmbedding row of token ids[0]
 * (int4/qlinear.py:122-130) and the device-side position bookkeeping:
 *   state[1] = state[0]  (tokens in the KV cache before this step, used by cgq_decode_attention)
 *   state[0] += 1
 *   state[2] += 1        (token counter: the epoch source of the tensor-parallel exchanges; state has >= 3 ints)
 * The host sets state[0] once after prefill; afterwards the step is a static CUDA graph.
 */
int cgq_decode_begin_w4(const int64_t* ids, const uint8_t* Wq /*[V/2, D]*/, const void* scale,
                        void* x, int V, int D, int group, int dtype, int* state, void* stream);

/* The int8 twins of the fused decode step's entry points: per-channel int8 weights [N, K] (K contiguous) with
 * scales [N] (int8/qlinear.py:77-107), int8 QEmbedding [V, D] with per-column scales [D] (int8/qlinear.py:110-132).  Same
 * prologues / residual epilogue and roundings as the int4 versions; the per-channel scale multiplies the fp32 sum. */
int cgq_w8a16_gemv_fused(const void* A, const int8_t* Wq, const void* scale, const void* bias, const void* resid,
                         void* C, int N, int K, int dtype, int prologue, const void* norm_w, float eps, void* stream);
int cgq_decode_begin_w8(const int64_t* ids, const int8_t* Wq /*[V, D]*/, const void* scale /*[D]*/, void* x, int V,
                        int D, int dtype, int* state, void* stream);

/*
 * ChatGLM2Attention.forward between qkv_proj and o_proj for one new token (model.py:140-174):
 * RoPE of q and k with row (state[1] + 1) of `freqs` (the model's freqs_cis_cache, [max_pos, d_head];
 * position ids are 1-based, model.py:296-297), append k / v to the caches at slot state[1]
 * (layout [max_len, n_groups, d_head], the reference's past_key_values layout without the batch and
 * broadcast axes), multi-query attention over slots 0..state[1], output [n_head * d_head].
 * One CTA per head for max_len <= 384, else a cluster of min(8, ceil(max_len / 128)) CTAs per head.
 * Online softmax in fp32: the scores are rounded to dtype as the reference's matmul does, the
 * probabilities are NOT rounded to dtype as the reference's `.type_as(x)` does (the more accurate side
 * of the parity bar).
 * d_head must be 64 or 128; state[1] must be < max_len (the launch is a no-op otherwise).
 */
int cgq_decode_attention(const void* qkv, const void* freqs, void* kcache, void* vcache, void* out,
                         const int* state, int n_head, int n_groups, int d_head, int max_len,
                         int dtype, void* stream);
/*
 * One-shot hint for the calling thread's NEXT cgq_decode_attention launch: the K / V caches (same layout, 16-byte
 * aligned) the FOLLOWING attention launch of the step will read -- the next layer's.  Their live rows are requested
 * into L2 one layer ahead (a cache row is read once per token, so it always comes from HBM, and a launch's own
 * loads queue behind the weight requests of the neighbouring linears).  NULL cancels.  No effect on results.
 */
void cgq_attention_next_kv(const void* kcache_next, const void* vcache_next);

/*
 * ---- Token sampling ---------------------------------------------------
 * chatglm_q.decoder.top_p_sampling (chatglm_q/decoder.py:12-27) for ONE logits row, in one launch:
 *   probs = softmax(float(logits) / temperature) over the whole vocabulary        (:14)
 *   the top_k largest, descending (ties: lower vocabulary index first)            (:15-17)
 *   probs[(cumsum(probs) - probs) > top_p] = 0;  probs /= sum(probs)              (:20-22)
 *   token = indices[argmax(probs / q)]   -- what torch.multinomial(probs, 1) computes from its
 *           Exp(1) variates q (:25-26); the caller draws q ([min(top_k, V)] fp32, e.g.
 *           torch.empty(k).exponential_(1): the same generator call multinomial makes, so the same
 *           seed gives the reference's token).
 *   logits  [V] dtype (fp16 / bf16: the exact radix select runs on 16-bit keys), 2-byte aligned
 *   q       [k] fp32 or NULL (NULL: only the distribution is produced; `token` must be NULL)
 *   token   [1] int64 or NULL;  probs [k] fp32 or NULL;  indices [k] int64 or NULL,  k = min(top_k, V)
 * Limits: top_k <= 1024; the row must fit in shared memory (V <= ~90 000); temperature > 0, top_p >= 0.
 * Deviation from the reference: the order is taken on the LOGITS, the reference sorts the fp32
 * probabilities -- identical unless two different logits round to the same probability.
 */
int cgq_top_p_sample(const void* logits, int V, int dtype, int top_k, float top_p, float temperature,
                     const float* q, int64_t* token, float* probs, int64_t* indices, void* stream);

/*
 * Profiling aid (no reference counterpart): the NEXT decode-kernel launch issued by the calling
 * thread writes a per-CTA timeline (8 x uint64 %globaltimer stamps per CTA, first 1024 CTAs:
 * entry, producer start, consumer dependency wait passed, first data, loop end, exit,
 * producer prefetch issued, producer done) into `device_buffer` (>= 64 KiB, zeroed by the
 * caller).  One-shot: the pointer is consumed by that launch.  NULL cancels.
 */
void cgq_debug_trace(void* device_buffer);

/*
 * Arithmetic of the int4 decode kernel for launches that do not name one (cgq_w4a16_gemm, cgq_w4a16_gemv_fused,
 * CGQ_IMPL_GEMV): 0 = IMMA.16832 on base-256 digits of the activation at M == 1 (default; CGQ_GEMV_ARITH overrides),
 * 1 = exact (q - 8) conversion + f16 / bf16 MMA, 2 = fp16 nibbles as subnormal MMA operands.  Process-wide; returns
 * the value in force before the call, any other `arith` only queries.  For tests / A-B timing (the persistent
 * programs keep the subnormal arithmetic, their bit-identity tests select it for the launch-per-linear side).
 */
int cgq_set_decode_arith(int arith);

/*
 * The shape-general CUDA-core kernels (CGQ_IMPL_SIMPLE) are what cgq_w4a16_gemm / cgq_w8a16_gemm fall back to for
 * shapes the TMA / tcgen05 kernels cannot take (N % 16 != 0, misaligned pointers, lda % 8 != 0, no workspace).  No
 * real layer of the reference's models should run on them: cgq_simple_fallback_count() counts how often AUTO took
 * them in this process, cgq_forbid_simple(1) (or CGQ_FORBID_SIMPLE=1) turns that into CGQ_ERR_BAD_SHAPE; returns the
 * previous setting.
 */
unsigned long long cgq_simple_fallback_count(void);
int cgq_forbid_simple(int on);

#ifdef __cplusplus
}
#endif
#endif /* CGQ_H_ */
