"""Host side of the dequant-matmul path: the same operator surface as the reference's
`chatglm_q/int4/triton_ops.py` and `chatglm_q/int8/triton_ops.py` (names, argument meaning and
error behaviour), bound to the sm_100a kernels through the C-ABI in include/cgq.h.

torch is used for device memory, streams and nothing else: every function here ends in one
`cgq_*` call with raw device pointers.  There is no CPU / eager fallback — a tensor the kernels
cannot take raises.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib
from ._lib import (IMPL_AUTO, IMPL_GEMV, IMPL_GEMV_EXACT, IMPL_GEMV_IMMA, IMPL_GEMV_SUBNORMAL, IMPL_GEMV_UMMA,  # noqa: F401
                   IMPL_SIMPLE, IMPL_TC)

_DTYPE_CODE = {torch.float16: 0, torch.bfloat16: 1}
_workspaces: dict[tuple[int, int], Tensor] = {}
_ws_bytes = None
# Set by install(): the REFERENCE's own implementations (its Triton kernels / its torch sampler), saved before the
# rebind.  Configurations this library does not build -- fp32 activations, group sizes other than 32, CPU / fp32
# logits or top_k > 1024 in the sampler -- are handed to them unchanged, so that install() never turns a call that
# works in the reference into an error.  Not a fallback of this library's own: without install() those inputs raise.
_delegates: dict[str, object] = {"s4": None, "s8": None, "s4t": None, "s8t": None, "sampler": None}


def check_input(a: Tensor) -> bool:
    """Reference: int4/triton_ops.py:10-11 — the fast path takes any CUDA tensor."""
    return a.get_device() >= 0


def _workspace(device: torch.device, stream: int) -> Tensor:
    """Zero-initialised, self-cleaning stream-K workspace, one per (device, stream)."""
    global _ws_bytes
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _workspaces.get(key)
    if ws is None:
        if _ws_bytes is None:
            _ws_bytes = int(_lib.load().cgq_workspace_bytes())
        ws = torch.zeros(_ws_bytes, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _dtype_code(t: Tensor) -> int:
    code = _DTYPE_CODE.get(t.dtype)
    if code is None:
        raise TypeError(
            f"cgq kernels compute in float16/bfloat16 activations, got {t.dtype} "
            "(no fp32 / CPU fallback exists on this path)")
    return code


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def dynamic_quant_matmul_s4(a: Tensor, b: Tensor, b_scale: Tensor, allow_tf32: bool = None,
                            bias: Tensor = None, impl: int = IMPL_AUTO) -> Tensor:
    """
    a:        (..., K)   float16 / bfloat16, CUDA
    b:        (K//2, N)  uint8, two K-adjacent int4 per byte (low nibble = even k), value = nibble - 8
    b_scale:  (G, N)     same dtype as a, G = K // 32
    returns:  (..., N)   new tensor, dtype of a

    Drop-in for chatglm_q.int4.triton_ops.dynamic_quant_matmul_s4 (int4/triton_ops.py:90-139).
    `allow_tf32` is accepted and ignored (16-bit inputs never used TF32).  `bias` (optional, [N])
    fuses the module's `out += bias` into the epilogue with the same two roundings.
    """
    output_shape = (*a.shape[:-1], b.shape[1])
    a = a.flatten(0, -2)
    assert len(b.shape) == 2
    assert len(b_scale.shape) == 2
    assert a.shape[1] == b.shape[0] * 2
    assert b.shape[1] == b_scale.shape[1]
    assert b.dtype == torch.uint8
    assert a.dtype == b_scale.dtype
    assert b.shape[0] % b_scale.shape[0] == 0
    assert a.get_device() >= 0
    assert b.get_device() == a.get_device(), f"{b.device=}, {a.device=}"
    assert b_scale.get_device() == a.get_device(), f"{b_scale.device=}, {a.device=}"
    M, K = a.shape
    G, N = b_scale.shape
    group = K // G
    if (group != 32 or a.dtype not in _DTYPE_CODE) and _delegates["s4"] is not None and bias is None:
        # fp32 / TF32 activations and other power-of-two groups exist in the reference (int4/triton_ops.py:120-123,
        # loader.create_quant_int4_model(group_size=...)): its own kernel keeps serving them after install()
        return _delegates["s4"](a.reshape(*output_shape[:-1], K), b, b_scale, allow_tf32)
    assert group == 32, f"only the reference model's group size 32 is built, got {group}"
    code = _dtype_code(a)
    if a.stride(1) != 1 or (M > 1 and a.stride(0) < K):
        a = a.contiguous()
    if not b.is_contiguous():
        b = b.contiguous()
    if not b_scale.is_contiguous():
        b_scale = b_scale.contiguous()
    if bias is not None:
        assert bias.shape == (N,) and bias.dtype == a.dtype and bias.get_device() == a.get_device()
        bias = bias.contiguous()
    c = torch.empty((M, N), device=a.device, dtype=a.dtype)
    if M == 0:
        return c.reshape(output_shape)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        stream = torch.cuda.current_stream().cuda_stream
        ws = _workspace(a.device, stream)
        _lib.check(lib.cgq_w4a16_gemm_ex(
            a.data_ptr(), a.stride(0) if M > 1 else K, b.data_ptr(), b_scale.data_ptr(), _ptr(bias),
            c.data_ptr(), N, M, N, K, group, code, ws.data_ptr(), ws.numel(), stream, impl))
    return c.reshape(output_shape)


def dynamic_quant_matmul(a: Tensor, b: Tensor, b_scale: Tensor, allow_tf32: bool = None,
                         bias: Tensor = None, impl: int = IMPL_AUTO) -> Tensor:
    """
    a:        (..., K)  float16 / bfloat16, CUDA
    b:        (K, N)    int8 — the `weight.t()` VIEW (strides (1, K)) of the module's [N, K] buffer
    b_scale:  (N)       same dtype as a
    returns:  (..., N)

    Drop-in for chatglm_q.int8.triton_ops.dynamic_quant_matmul (int8/triton_ops.py:87-127).
    The kernels read the weight K-contiguous, i.e. exactly the module buffer; a `b` that is not such
    a transposed view is copied once into that layout.
    """
    output_shape = (*a.shape[:-1], b.shape[1])
    a = a.flatten(0, -2)
    assert len(b.shape) == 2
    assert len(b_scale.shape) == 1
    assert a.shape[1] == b.shape[0]
    assert b.shape[1] == b_scale.shape[0]
    assert b.dtype == torch.int8
    assert a.dtype == b_scale.dtype
    assert a.get_device() >= 0
    assert b.get_device() == a.get_device(), f"{b.device=}, {a.device=}"
    assert b_scale.get_device() == a.get_device(), f"{b_scale.device=}, {a.device=}"
    M, K = a.shape
    _, N = b.shape
    if a.dtype not in _DTYPE_CODE and _delegates["s8"] is not None and bias is None:
        return _delegates["s8"](a.reshape(*output_shape[:-1], K), b, b_scale, allow_tf32)   # fp32: the reference's kernel
    code = _dtype_code(a)
    if a.stride(1) != 1 or (M > 1 and a.stride(0) < K):
        a = a.contiguous()
    w_nk = b.t()  # [N, K]
    if not w_nk.is_contiguous():
        w_nk = w_nk.contiguous()
    if not b_scale.is_contiguous():
        b_scale = b_scale.contiguous()
    if bias is not None:
        assert bias.shape == (N,) and bias.dtype == a.dtype and bias.get_device() == a.get_device()
        bias = bias.contiguous()
    c = torch.empty((M, N), device=a.device, dtype=a.dtype)
    if M == 0:
        return c.reshape(output_shape)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        stream = torch.cuda.current_stream().cuda_stream
        ws = _workspace(a.device, stream)
        _lib.check(lib.cgq_w8a16_gemm_ex(
            a.data_ptr(), a.stride(0) if M > 1 else K, w_nk.data_ptr(), b_scale.data_ptr(),
            _ptr(bias), c.data_ptr(), N, M, N, K, code, ws.data_ptr(), ws.numel(), stream, impl))
    return c.reshape(output_shape)


def dynamic_quant_matmul_transposed_s4(a: Tensor, b: Tensor, b_scale: Tensor, allow_tf32: bool = None) -> Tensor:
    """Backward of `dynamic_quant_matmul_s4` with respect to its activation: a = grad_out (..., N), b uint8 (K//2, N),
    b_scale (K//32, N) -> grad_A (..., K) = a @ unpack_int4(b, b_scale).T.  Drop-in for
    chatglm_q.int4.triton_ops.dynamic_quant_matmul_transposed_s4 (int4/triton_ops.py:208-264) without its power-of-two
    restrictions; `DynamicQuantizeMatMul.backward` calls it (int4/qlinear.py:53-64)."""
    K = b.shape[0] * 2
    output_shape = (*a.shape[:-1], K)
    a = a.flatten(0, -2)
    assert len(b.shape) == 2 and len(b_scale.shape) == 2
    assert a.shape[1] == b.shape[1] and b.shape[1] == b_scale.shape[1]
    assert b.dtype == torch.uint8 and a.dtype == b_scale.dtype
    assert a.get_device() >= 0 and b.get_device() == a.get_device() and b_scale.get_device() == a.get_device()
    M, N = a.shape
    group = K // b_scale.shape[0]
    if (group != 32 or a.dtype not in _DTYPE_CODE) and _delegates["s4t"] is not None:
        return _delegates["s4t"](a.reshape(*output_shape[:-1], N), b, b_scale, allow_tf32)
    assert group == 32, f"only the reference model's group size 32 is built, got {group}"
    code = _dtype_code(a)
    a, b, b_scale = a.contiguous(), b.contiguous(), b_scale.contiguous()
    c = torch.empty((M, K), device=a.device, dtype=a.dtype)
    if M:
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().cgq_w4a16_grad_a(a.data_ptr(), N, b.data_ptr(), b_scale.data_ptr(), c.data_ptr(), K, M, N,
                                                    K, group, code, torch.cuda.current_stream().cuda_stream))
    return c.reshape(output_shape)


def dynamic_quant_matmul_transposed(a: Tensor, b: Tensor, b_scale: Tensor, allow_tf32: bool = None) -> Tensor:
    """int8 twin (int8/triton_ops.py:196-245, `DynamicQuantizeMatMul.backward` int8/qlinear.py:41-52): a = grad_out
    (..., N), b the (K, N) transposed VIEW of the module's [N, K] int8 buffer, b_scale (N,) -> grad_A (..., K)."""
    K, N = b.shape
    output_shape = (*a.shape[:-1], K)
    a = a.flatten(0, -2)
    assert a.shape[1] == N and b_scale.shape == (N,) and b.dtype == torch.int8 and a.dtype == b_scale.dtype
    assert a.get_device() >= 0 and b.get_device() == a.get_device() and b_scale.get_device() == a.get_device()
    if a.dtype not in _DTYPE_CODE and _delegates["s8t"] is not None:
        return _delegates["s8t"](a.reshape(*output_shape[:-1], N), b, b_scale, allow_tf32)
    code = _dtype_code(a)
    w_nk = b.t()
    if not w_nk.is_contiguous():
        w_nk = w_nk.contiguous()
    a, b_scale = a.contiguous(), b_scale.contiguous()
    M = a.shape[0]
    c = torch.empty((M, K), device=a.device, dtype=a.dtype)
    if M:
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().cgq_w8a16_grad_a(a.data_ptr(), N, w_nk.data_ptr(), b_scale.data_ptr(), c.data_ptr(), K, M,
                                                    N, K, code, torch.cuda.current_stream().cuda_stream))
    return c.reshape(output_shape)


def unpack_int4_i8(x: Tensor) -> Tensor:
    """[K/2, N] uint8 -> [K, N] int8 (nibble - 8); bit-exact twin of int4/qlinear.py:29-31."""
    assert x.dtype == torch.uint8 and x.dim() == 2 and x.get_device() >= 0
    x = x.contiguous()
    K, N = x.shape[0] * 2, x.shape[1]
    out = torch.empty((K, N), dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().cgq_w4_unpack_i8(x.data_ptr(), out.data_ptr(), K, N,
                                               torch.cuda.current_stream().cuda_stream))
    return out


def unpack_int4(x: Tensor, x_scale: Tensor) -> Tensor:
    """[K/2, N] uint8, [G, N] scale -> dequantised [K, N]; bit-exact twin of
    chatglm_q.int4.qlinear.unpack_int4 (int4/qlinear.py:20-33)."""
    assert x.dtype == torch.uint8 and x.dim() == 2 and x.get_device() >= 0
    K = x.shape[0] * 2
    G, N = x_scale.shape
    assert x.shape[1] == N
    assert K % G == 0, f"{K=}, {G=}"
    code = _dtype_code(x_scale)
    x, x_scale = x.contiguous(), x_scale.contiguous()
    out = torch.empty((K, N), dtype=x_scale.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().cgq_w4_dequant(x.data_ptr(), x_scale.data_ptr(), out.data_ptr(), K, N,
                                             K // G, code, torch.cuda.current_stream().cuda_stream))
    return out


def embedding_s4(ids: Tensor, weight: Tensor, weight_scale: Tensor, group_size: int = 32) -> Tensor:
    """QEmbedding.forward of the int4 model (int4/qlinear.py:122-130)."""
    assert ids.dtype == torch.int64 and ids.get_device() >= 0
    code = _dtype_code(weight_scale)
    V, D = weight.shape[0] * 2, weight.shape[1]
    flat = ids.reshape(-1).contiguous()
    out = torch.empty((flat.numel(), D), dtype=weight_scale.dtype, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(_lib.load().cgq_w4_embedding(
            flat.data_ptr(), flat.numel(), weight.data_ptr(), weight_scale.data_ptr(), out.data_ptr(),
            V, D, group_size, code, torch.cuda.current_stream().cuda_stream))
    return out.reshape(*ids.shape, D)


def embedding_s8(ids: Tensor, weight: Tensor, weight_scale: Tensor) -> Tensor:
    """QEmbedding.forward of the int8 model (int8/qlinear.py:118-120)."""
    assert ids.dtype == torch.int64 and ids.get_device() >= 0
    code = _dtype_code(weight_scale)
    V, D = weight.shape
    flat = ids.reshape(-1).contiguous()
    out = torch.empty((flat.numel(), D), dtype=weight_scale.dtype, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.check(_lib.load().cgq_w8_embedding(
            flat.data_ptr(), flat.numel(), weight.data_ptr(), weight_scale.data_ptr(), out.data_ptr(),
            V, D, code, torch.cuda.current_stream().cuda_stream))
    return out.reshape(*ids.shape, D)


# ---------------------------------------------------------------------- fused decode-step pieces
class decode_arith:
    """`with ops.decode_arith(_lib.ARITH_SUBNORMAL): ...` -- arithmetic of the int4 decode kernel for launches that do
    not name one (cgq_set_decode_arith, include/cgq.h).  Process-wide: tests and A/B timing only."""

    def __init__(self, arith: int):
        self.arith = arith

    def __enter__(self):
        self.prev = _lib.load().cgq_set_decode_arith(self.arith)
        return self

    def __exit__(self, *exc):
        _lib.load().cgq_set_decode_arith(self.prev)
        return False


def gemv_fused_s4(a: Tensor, b: Tensor, b_scale: Tensor, bias: Tensor = None, resid: Tensor = None,
                  prologue: int = _lib.PRO_NONE, norm_weight: Tensor = None, eps: float = 0.0,
                  out: Tensor = None) -> Tensor:
    """One-row int4 linear of the fused decode step (cgq_w4a16_gemv_fused, include/cgq.h):
    out[N] = (resid +) round(prologue(a) · dequant(b, b_scale)) (+ bias).

    prologue PRO_RMSNORM: a [K] is the residual stream, normalised as RMSNorm.forward does (model.py:68-73);
    PRO_SILU_GATE: a [2K] is w_in's output, a' = silu(a[:K]) * a[K:] (model.py:200-201).
    `out` may alias `resid` (the in-place residual update of the step)."""
    K, N = b.shape[0] * 2, b.shape[1]
    code = _dtype_code(a)
    assert a.dim() == 1 and a.is_contiguous() and a.numel() == (2 * K if prologue == _lib.PRO_SILU_GATE else K)
    assert b.dtype == torch.uint8 and b.is_contiguous() and b_scale.is_contiguous()
    assert b_scale.shape == (K // 32, N) and b_scale.dtype == a.dtype
    for t in (bias, resid):
        assert t is None or (t.shape == (N,) and t.dtype == a.dtype and t.is_contiguous())
    if prologue == _lib.PRO_RMSNORM:
        assert norm_weight is not None and norm_weight.shape == (K,) and norm_weight.dtype == a.dtype
    if out is None:
        out = torch.empty(N, device=a.device, dtype=a.dtype)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().cgq_w4a16_gemv_fused(
            a.data_ptr(), b.data_ptr(), b_scale.data_ptr(), _ptr(bias), _ptr(resid), out.data_ptr(), N, K, 32,
            code, prologue, _ptr(norm_weight), float(eps), torch.cuda.current_stream().cuda_stream))
    return out


def gemv_fused_s8(a: Tensor, b: Tensor, b_scale: Tensor, bias: Tensor = None, resid: Tensor = None,
                  prologue: int = _lib.PRO_NONE, norm_weight: Tensor = None, eps: float = 0.0,
                  out: Tensor = None) -> Tensor:
    """The int8 twin of `gemv_fused_s4` (cgq_w8a16_gemv_fused): `b` is the module's [N, K] int8 buffer itself
    (int8/qlinear.py:82-87), `b_scale` its [N] per-channel scales."""
    N, K = b.shape
    code = _dtype_code(a)
    assert a.dim() == 1 and a.is_contiguous() and a.numel() == (2 * K if prologue == _lib.PRO_SILU_GATE else K)
    assert b.dtype == torch.int8 and b.is_contiguous() and b_scale.is_contiguous()
    assert b_scale.shape == (N,) and b_scale.dtype == a.dtype
    for t in (bias, resid):
        assert t is None or (t.shape == (N,) and t.dtype == a.dtype and t.is_contiguous())
    if prologue == _lib.PRO_RMSNORM:
        assert norm_weight is not None and norm_weight.shape == (K,) and norm_weight.dtype == a.dtype
    if out is None:
        out = torch.empty(N, device=a.device, dtype=a.dtype)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().cgq_w8a16_gemv_fused(
            a.data_ptr(), b.data_ptr(), b_scale.data_ptr(), _ptr(bias), _ptr(resid), out.data_ptr(), N, K, code, prologue,
            _ptr(norm_weight), float(eps), torch.cuda.current_stream().cuda_stream))
    return out


def decode_attention(qkv: Tensor, freqs: Tensor, k_cache: Tensor, v_cache: Tensor, state: Tensor,
                     n_head: int, n_groups: int, d_head: int) -> Tensor:
    """ChatGLM2Attention.forward between qkv_proj and o_proj for one new token (cgq_decode_attention):
    RoPE with row state[1]+1 of `freqs`, append k/v at slot state[1] of the [max_len, n_groups, d_head]
    caches (in place), attention over slots 0..state[1].  Returns [n_head * d_head]."""
    code = _dtype_code(qkv)
    max_len = k_cache.shape[0]
    assert qkv.numel() == d_head * (n_head + 2 * n_groups) and qkv.is_contiguous()
    assert k_cache.is_contiguous() and v_cache.is_contiguous() and k_cache.numel() == max_len * n_groups * d_head
    assert state.dtype == torch.int32 and state.numel() >= 2 and freqs.is_contiguous() and freqs.dtype == qkv.dtype
    out = torch.empty(n_head * d_head, device=qkv.device, dtype=qkv.dtype)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().cgq_decode_attention(
            qkv.data_ptr(), freqs.data_ptr(), k_cache.data_ptr(), v_cache.data_ptr(), out.data_ptr(),
            state.data_ptr(), n_head, n_groups, d_head, max_len, code, torch.cuda.current_stream().cuda_stream))
    return out


# ---------------------------------------------------------------------- sampling
def _sample_rows(logits: Tensor, top_k: int, top_p: float, temperature: float, want_token: bool,
                 q: Tensor = None):
    assert logits.get_device() >= 0, "top_p_sampling: CUDA logits only (no CPU fallback on this path)"
    assert logits.dim() >= 1 and logits.shape[-1] >= 1
    assert temperature > 0 and top_p >= 0 and top_k >= 1
    code = _dtype_code(logits)
    V = logits.shape[-1]
    k = min(int(top_k), V)
    rows = logits.reshape(-1, V)
    if rows.stride(1) != 1:
        rows = rows.contiguous()
    n = rows.shape[0]
    lib = _lib.load()
    with torch.cuda.device(logits.device):
        stream = torch.cuda.current_stream().cuda_stream
        if want_token:
            # the variates torch.multinomial(probs, 1) draws for itself: empty_like(probs).exponential_(1)
            if q is None:
                q = torch.empty((n, k), dtype=torch.float32, device=logits.device).exponential_(1)
            else:
                assert q.dtype == torch.float32 and q.numel() == n * k and q.device == logits.device
                q = q.reshape(n, k).contiguous()
            token = torch.empty(n, dtype=torch.int64, device=logits.device)
            for r in range(n):
                _lib.check(lib.cgq_top_p_sample(
                    rows[r].data_ptr(), V, code, int(top_k), float(top_p), float(temperature),
                    q[r].data_ptr(), token[r].data_ptr(), None, None, stream))
            return token.reshape(logits.shape[:-1])
        probs = torch.empty((n, k), dtype=torch.float32, device=logits.device)
        indices = torch.empty((n, k), dtype=torch.int64, device=logits.device)
        for r in range(n):
            _lib.check(lib.cgq_top_p_sample(
                rows[r].data_ptr(), V, code, int(top_k), float(top_p), float(temperature),
                None, None, probs[r].data_ptr(), indices[r].data_ptr(), stream))
        return probs.reshape(*logits.shape[:-1], k), indices.reshape(*logits.shape[:-1], k)


def top_p_sampling(logits: Tensor, top_k=100, top_p=0.8, temperature=1.0, *, q: Tensor = None) -> Tensor:
    """Drop-in for chatglm_q.decoder.top_p_sampling (chatglm_q/decoder.py:12-27): logits (..., V) fp16 / bf16 on
    the GPU -> sampled token ids (...) int64.  Two launches per call (the Exp(1) draw torch.multinomial would make,
    then cgq_top_p_sample per row) instead of ~15 and a host synchronisation; with the same torch seed it
    returns the token the reference returns except at fp32 rounding ties (the softmax / cumulative sums are added in
    another order: a probability sitting exactly on the top-p edge, or two candidates whose p/q agree to the last
    bit, may resolve differently).  `q` (keyword-only, not in the reference): the Exp(1) variates
    (..., min(top_k, V)) fp32 to use instead of drawing them -- reproducible sampling for tests."""
    ref = _delegates["sampler"]
    if ref is not None and q is None and (logits.get_device() < 0 or logits.dtype not in _DTYPE_CODE
                                          or min(int(top_k), logits.shape[-1]) > 1024 or logits.shape[-1] > 90000):
        return ref(logits, top_k, top_p, temperature)      # what the one-launch kernel does not take: the reference's own
    return _sample_rows(logits, top_k, top_p, temperature, True, q)


def top_p_distribution(logits: Tensor, top_k=100, top_p=0.8, temperature=1.0) -> tuple[Tensor, Tensor]:
    """The (probs, indices) pair chatglm_q/decoder.py:14-22 hands to torch.multinomial / torch.gather:
    top_k probabilities, descending, masked by top_p and renormalised (fp32), and their vocabulary ids."""
    return _sample_rows(logits, top_k, top_p, temperature, False)


def prefetch_next_s4(b: Tensor, b_scale: Tensor) -> None:
    """One-shot hint (cgq_prefetch_next_w4): the NEXT int4 decode launch of this thread also streams the
    leading part of THIS weight — the one the launch after it will read — from HBM into L2."""
    assert b.dtype == torch.uint8 and b.is_contiguous() and b_scale.is_contiguous() and b.get_device() >= 0
    _lib.check(_lib.load().cgq_prefetch_next_w4(b.data_ptr(), b_scale.data_ptr(), b.shape[1], b.shape[0] * 2))


EPI_NONE, EPI_SILU_PAIR = _lib.EPI_NONE, _lib.EPI_SILU_PAIR


class StepProgram:
    """A decode-step program (cgq_step_*, include/cgq.h): phases -- int4g32 linears with fused prologues / epilogues,
    the RoPE + KV-append + attention of one query row, the embedding row -- executed strictly in order by ONE
    persistent cooperative kernel (csrc/decode_mk.cu; one CTA per SM, weights streamed through a shared-memory
    ring across phase boundaries, a grid barrier between phases).  `FusedDecodeModel` builds the whole token with
    it; this class is the operator-level handle (tests, tools).  Every tensor must stay alive and in place while
    the program exists (references are kept).  float16 only."""

    def __init__(self, dtype: torch.dtype = torch.float16, state: Tensor = None):
        if dtype != torch.float16:
            raise TypeError("the one-launch step is built for float16 (the reference checkpoints' dtype)")
        self.code = _DTYPE_CODE[dtype]
        self.dtype = dtype
        self.state = state            # [>= 1] int32: tokens in the KV cache before the step (attention phases)
        self._ops: list = []
        self._keep: list = [state]
        self._handle = None

    def linear(self, a: Tensor, b: Tensor, b_scale: Tensor, out: Tensor, bias: Tensor = None, resid: Tensor = None,
               prologue: int = _lib.PRO_NONE, norm_weight: Tensor = None, eps: float = 0.0, epilogue: int = 0,
               k: int = None) -> None:
        """out[N] = (resid +) round(prologue(a) . dequant(b, b_scale)) (+ bias); EPI_SILU_PAIR stores
        silu(out[:N/2]) * out[N/2:] into out[:N/2] instead.  `k` (default: from b) only documents intent."""
        assert self._handle is None, "program already built"
        K, N = b.shape[0] * 2, b.shape[1]
        assert k is None or k == K
        assert a.dtype == self.dtype and a.is_contiguous() and a.numel() >= (2 * K if prologue == _lib.PRO_SILU_GATE else K)
        assert b.dtype == torch.uint8 and b.is_contiguous() and b_scale.is_contiguous() and b_scale.shape == (K // 32, N)
        assert out.dtype == self.dtype and out.is_contiguous() and out.numel() >= (N // 2 if epilogue == EPI_SILU_PAIR else N)
        self._ops.append(_lib.StepOp(kind=_lib.STEP_LINEAR, Wq=b.data_ptr(), scale=b_scale.data_ptr(), bias=_ptr(bias),
                                     A=a.data_ptr(), C=out.data_ptr(), resid=_ptr(resid), norm_w=_ptr(norm_weight), N=N,
                                     K=K, prologue=prologue, eps=float(eps), epilogue=epilogue))
        self._keep += [a, b, b_scale, out, bias, resid, norm_weight]

    def attention(self, qkv: Tensor, freqs: Tensor, k_cache: Tensor, v_cache: Tensor, out: Tensor, n_head: int,
                  n_groups: int, d_head: int) -> None:
        assert self._handle is None and self.state is not None, "attention phases need the `state` tensor"
        max_len = k_cache.shape[0]
        assert qkv.is_contiguous() and k_cache.is_contiguous() and v_cache.is_contiguous() and freqs.is_contiguous()
        self._ops.append(_lib.StepOp(kind=_lib.STEP_ATTENTION, A=qkv.data_ptr(), C=out.data_ptr(), freqs=freqs.data_ptr(),
                                     kcache=k_cache.data_ptr(), vcache=v_cache.data_ptr(), n_head=n_head,
                                     n_groups=n_groups, d_head=d_head, max_len=max_len))
        self._keep += [qkv, freqs, k_cache, v_cache, out]

    def embed(self, ids: Tensor, weight: Tensor, weight_scale: Tensor, out: Tensor) -> None:
        assert self._handle is None and ids.dtype == torch.int64
        self._ops.append(_lib.StepOp(kind=_lib.STEP_EMBED, Wq=weight.data_ptr(), scale=weight_scale.data_ptr(),
                                     C=out.data_ptr(), N=weight.shape[1], V=weight.shape[0] * 2, ids=ids.data_ptr()))
        self._keep += [ids, weight, weight_scale, out]

    def build(self) -> "StepProgram":
        import ctypes

        arr = (_lib.StepOp * len(self._ops))(*self._ops)
        handle = ctypes.c_uint64(0)
        dev = next(t for t in self._keep if t is not None).device
        with torch.cuda.device(dev):
            _lib.check(_lib.load().cgq_step_create(arr, len(self._ops), self.code, _ptr(self.state), ctypes.byref(handle)))
        self._handle, self._device = handle.value, dev
        return self

    def run(self) -> None:
        if self._handle is None:
            self.build()
        with torch.cuda.device(self._device):
            _lib.check(_lib.load().cgq_step_run(self._handle, torch.cuda.current_stream().cuda_stream))

    def status(self) -> tuple[int, int, bool]:
        """(CTAs, ring stages per CTA, whether a grid barrier ever timed out) -- synchronises the device."""
        import ctypes

        c, st, f = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.load().cgq_step_status(self._handle, ctypes.byref(c), ctypes.byref(st), ctypes.byref(f)))
        return c.value, st.value, bool(f.value)

    def __del__(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().cgq_step_destroy(self._handle)
            except Exception:  # noqa: BLE001
                pass


class DecodeProgram:
    """A chain of batch-1 int4g32 linears executed by ONE persistent launch (cgq_program_*, include/cgq.h).

    `add()` takes the arguments of `gemv_fused_s4`; every tensor must stay alive and in place while the program
    exists (the object keeps references).  `run()` = one memset + one kernel on the current stream, bit-identical
    to calling `gemv_fused_s4` once per linear.  Measured on B200 it is ~15 % SLOWER than the launch-per-linear
    chain with programmatic dependent launch: opt-in groundwork, not the default decode path."""

    def __init__(self, dtype: torch.dtype):
        self.code = _DTYPE_CODE[dtype]
        self.dtype = dtype
        self._ops: list[_lib.LinearOp] = []
        self._keep: list = []
        self._handle = None

    def add(self, a: Tensor, b: Tensor, b_scale: Tensor, out: Tensor, bias: Tensor = None, resid: Tensor = None,
            prologue: int = _lib.PRO_NONE, norm_weight: Tensor = None, eps: float = 0.0) -> None:
        assert self._handle is None, "program already built"
        K, N = b.shape[0] * 2, b.shape[1]
        assert a.dtype == self.dtype and a.is_contiguous() and a.numel() >= (2 * K if prologue == _lib.PRO_SILU_GATE else K)
        assert b.dtype == torch.uint8 and b.is_contiguous() and b_scale.is_contiguous() and b_scale.shape == (K // 32, N)
        assert out.dtype == self.dtype and out.is_contiguous() and out.numel() >= N
        self._ops.append(_lib.LinearOp(b.data_ptr(), b_scale.data_ptr(), _ptr(bias), a.data_ptr(), out.data_ptr(),
                                       _ptr(resid), _ptr(norm_weight), N, K, prologue, float(eps)))
        self._keep += [a, b, b_scale, out, bias, resid, norm_weight]

    def build(self) -> "DecodeProgram":
        import ctypes

        arr = (_lib.LinearOp * len(self._ops))(*self._ops)
        handle = ctypes.c_uint64(0)
        with torch.cuda.device(self._keep[0].device):
            _lib.check(_lib.load().cgq_program_create(arr, len(self._ops), self.code, ctypes.byref(handle)))
        self._handle = handle.value
        return self

    def run(self) -> None:
        if self._handle is None:
            self.build()
        with torch.cuda.device(self._keep[0].device):
            _lib.check(_lib.load().cgq_program_run(self._handle, torch.cuda.current_stream().cuda_stream))

    def status(self) -> tuple[int, bool]:
        """(number of persistent workers, whether a grid barrier ever timed out) — synchronises the device."""
        import ctypes

        w, f = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.load().cgq_program_status(self._handle, ctypes.byref(w), ctypes.byref(f)))
        return w.value, bool(f.value)

    def __del__(self):
        if getattr(self, "_handle", None) is not None:
            try:
                _lib.load().cgq_program_destroy(self._handle)
            except Exception:
                pass
