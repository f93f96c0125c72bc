"""ORACLE — test infrastructure, NOT product code.

CPU restatement (numpy only, no torch, no CUDA) of the reference algorithm for the int4g32 / int8
weight-only dequant-matmul path of K024/chatglm-q.  Every function cites the reference lines it
restates.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import
this module; `chatglm_q_b200/` never does (a test enforces that).

Parity is PINNED: `tests/test_oracle_golden.py` checks every function here against
`tests/golden/*.npz`, vectors produced by importing the real reference from /root/reference
(`tests/golden/make_golden.py`, committed), plus the known-answer vectors of SURVEY.md §8(c).

Value conventions: 16-bit tensors travel as numpy float16, or — for bfloat16, which numpy lacks —
as float32 arrays whose values are exactly representable in bfloat16 (`round_to(x, "bfloat16")`).
"""
from __future__ import annotations

import numpy as np

GROUP = 32
DTYPES = ("float32", "float16", "bfloat16")


# ----------------------------------------------------------------------------- rounding helpers
def bf16_round(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest-even bfloat16, returned as float32 (what torch .bfloat16() does)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(np.float32)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.float32(np.nan), out)
    return out.reshape(np.shape(x))


def bf16_bits(x: np.ndarray) -> np.ndarray:
    """bfloat16-representable float32 array -> uint16 bit patterns (storage for fixtures)."""
    return (np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) >> 16).astype(np.uint16)


def bf16_from_bits(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round float32 values once to `dtype`; result stays float32 (exactly representable)."""
    x = np.asarray(x, dtype=np.float32)
    if dtype == "float32":
        return x
    if dtype == "float16":
        with np.errstate(over="ignore"):
            return x.astype(np.float16).astype(np.float32)
    if dtype == "bfloat16":
        return bf16_round(x)
    raise ValueError(dtype)


# ----------------------------------------------------------------------------- int4: unpack / dequant
def unpack_int4_i8(x: np.ndarray) -> np.ndarray:
    """[K/2, N] uint8 -> [K, N] int8 = nibble - 8; low nibble is the EVEN k.
    Reference: chatglm_q/int4/qlinear.py:29-31 (`repeat`, `>> [0, 4]`, `& 0xF`, `.to(int8) - 8`)."""
    assert x.dtype == np.uint8 and x.ndim == 2
    k2, n = x.shape
    out = np.empty((k2, 2, n), dtype=np.int8)
    out[:, 0, :] = (x & 0xF).astype(np.int8) - 8
    out[:, 1, :] = (x >> 4).astype(np.int8) - 8
    return out.reshape(2 * k2, n)


def unpack_int4(x: np.ndarray, x_scale: np.ndarray, dtype: str = "float16") -> np.ndarray:
    """Dequantised [K, N] weight: round_dtype((nibble - 8) * scale[k // GROUP_K, n]), one rounding.
    Reference: chatglm_q/int4/qlinear.py:20-33.  `x_scale` holds dtype-representable values."""
    k = x.shape[0] * 2
    g, n = x_scale.shape
    assert x.shape[1] == n
    assert k % g == 0, f"{k=}, {g=}"
    group_k = k // g
    q = unpack_int4_i8(x).astype(np.float32).reshape(g, group_k, n)
    w = q * np.asarray(x_scale, dtype=np.float32)[:, None, :]  # exact in fp32 (4-bit x <=11-bit)
    return round_to(w.reshape(k, n), dtype)


def qmatmul_int4(a: np.ndarray, b: np.ndarray, b_scale: np.ndarray, bias: np.ndarray | None = None,
                 dtype: str = "float16") -> np.ndarray:
    """DynamicQuantizeLinear.forward of the int4 model on its torch path:
    `out = A.matmul(unpack_int4(B, b_scale)); out += bias` (int4/qlinear.py:47-50, 90-94).
    fp32 accumulation, product rounded to dtype, bias added as a second rounded op."""
    w = unpack_int4(b, b_scale, dtype)
    lead = a.shape[:-1]
    a2 = np.asarray(a, dtype=np.float32).reshape(-1, a.shape[-1])
    out = round_to(a2 @ w, dtype)
    if bias is not None:
        out = round_to(out + np.asarray(bias, dtype=np.float32)[None, :], dtype)
    return out.reshape(*lead, w.shape[1])


def quantize_int4(x: np.ndarray, group_k: int = GROUP):
    """Round-to-nearest symmetric int4 per (group of `group_k` rows, column), packed 2 per byte.
    Reference: chatglm_q/int4/quantizer.py:12-29.  x is the [K, N] (in_dim, out_dim) weight."""
    x = np.asarray(x, dtype=np.float32)
    k, n = x.shape
    assert k % group_k == 0
    g = k // group_k
    xg = x.reshape(g, group_k, n)
    w_max = np.abs(xg).max(axis=1, keepdims=True)
    scale = np.maximum(w_max / np.float32(7), np.float32(1e-10)).astype(np.float32)
    q = np.clip(np.rint(xg / scale), -7, 7)
    q = (q + 8).astype(np.uint8).reshape(k, n)
    packed = (q[0::2, :] & 0xF) | ((q[1::2, :] & 0xF) << 4)
    return np.ascontiguousarray(packed.astype(np.uint8)), scale.reshape(g, n)


def qembedding_int4(ids: np.ndarray, weight: np.ndarray, weight_scale: np.ndarray,
                    dtype: str = "float16", group: int = GROUP) -> np.ndarray:
    """int4 QEmbedding.forward — the table is packed along the VOCAB axis (2 tokens per byte).
    Reference: chatglm_q/int4/qlinear.py:122-130."""
    ids = np.asarray(ids, dtype=np.int64)
    scales = np.asarray(weight_scale, dtype=np.float32)[ids // group]
    emb = weight[ids // 2]
    shifts = ((ids % 2) * 4)[..., None].astype(np.uint8)
    q = ((emb >> shifts) & 0xF).astype(np.int8) - 8
    return round_to(q.astype(np.float32) * scales, dtype)


# ----------------------------------------------------------------------------- int8
def qmatmul_int8(a: np.ndarray, w_nk: np.ndarray, w_scale: np.ndarray, bias: np.ndarray | None = None,
                 dtype: str = "float16") -> np.ndarray:
    """DynamicQuantizeLinear.forward of the int8 model on its torch path:
    `A.matmul(weight.t() * weight_scale)` (+ bias) — int8/qlinear.py:38, 89-93.
    `w_nk` is the module buffer [N, K]; `weight.t() * scale` rounds once to dtype per element."""
    assert w_nk.dtype == np.int8 and w_nk.ndim == 2
    w = round_to(w_nk.T.astype(np.float32) * np.asarray(w_scale, dtype=np.float32)[None, :], dtype)
    lead = a.shape[:-1]
    a2 = np.asarray(a, dtype=np.float32).reshape(-1, a.shape[-1])
    out = round_to(a2 @ w, dtype)
    if bias is not None:
        out = round_to(out + np.asarray(bias, dtype=np.float32)[None, :], dtype)
    return out.reshape(*lead, w.shape[1])


def qmatmul_int4_grad_a(grad_out: np.ndarray, b: np.ndarray, b_scale: np.ndarray, dtype: str = "float16") -> np.ndarray:
    """`DynamicQuantizeMatMul.backward` of the int4 model on its torch path:
    `grad_A = grad_out.matmul(unpack_int4(B, b_scale).t())` (int4/qlinear.py:53-64).
    Dequantised weight rounded once per element, fp32 accumulation, one final rounding."""
    w = unpack_int4(b, b_scale, dtype)
    lead = grad_out.shape[:-1]
    g2 = np.asarray(grad_out, dtype=np.float32).reshape(-1, grad_out.shape[-1])
    return round_to(g2 @ w.T, dtype).reshape(*lead, w.shape[0])


def qmatmul_int8_grad_a(grad_out: np.ndarray, w_nk: np.ndarray, w_scale: np.ndarray, dtype: str = "float16") -> np.ndarray:
    """int8 twin: `grad_A = grad_out.matmul((B * b_scale).t())` with B = weight.t() (int8/qlinear.py:41-52)."""
    assert w_nk.dtype == np.int8 and w_nk.ndim == 2
    w = round_to(w_nk.T.astype(np.float32) * np.asarray(w_scale, dtype=np.float32)[None, :], dtype)
    lead = grad_out.shape[:-1]
    g2 = np.asarray(grad_out, dtype=np.float32).reshape(-1, grad_out.shape[-1])
    return round_to(g2 @ w.T, dtype).reshape(*lead, w.shape[0])


def quantize_int8(x: np.ndarray):
    """Per-row abs-max / 127 symmetric int8.  Reference: chatglm_q/int8/quantizer.py:11-19
    (x is the [N, K] (out_dim, in_dim) weight)."""
    x = np.asarray(x, dtype=np.float32)
    w_max = np.abs(x).max(axis=1, keepdims=True)
    scale = np.maximum(w_max / np.float32(127), np.float32(1e-10)).astype(np.float32)
    q = np.clip(np.rint(x / scale), -127, 127).astype(np.int8)
    return q, scale[:, 0]


def qembedding_int8(ids: np.ndarray, weight: np.ndarray, weight_scale: np.ndarray,
                    dtype: str = "float16") -> np.ndarray:
    """int8 QEmbedding.forward: `embedding(input, weight) * weight_scale` (int8/qlinear.py:118-120)."""
    emb = weight[np.asarray(ids, dtype=np.int64)].astype(np.float32)
    return round_to(emb * np.asarray(weight_scale, dtype=np.float32), dtype)


# ----------------------------------------------------------------------------- byte / flop accounting
def algorithmic_bytes(kind: str, m: int, n: int, k: int, bias: bool = False) -> int:
    """the byte count `roofline.achieved` is computed from (16-bit activations)."""
    act = 2 * m * k + 2 * m * n + (2 * n if bias else 0)
    if kind == "int4":
        return k * n // 2 + (k // GROUP) * n * 2 + act
    if kind == "int8":
        return k * n + 2 * n + act
    raise ValueError(kind)


def flops(m: int, n: int, k: int) -> int:
    return 2 * m * n * k
