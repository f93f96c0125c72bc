"""CPU oracle of the fused batch-1 decode step — TEST INFRASTRUCTURE ONLY.

numpy restatement of what `ChatGLM2Model.forward` (reference chatglm_q/model.py:329-392) computes for
ONE new token against a KV cache, with the reference's rounding points for 16-bit activations.  Only
`tests/`, `__graft_entry__.smoke()` and bench.py's CPU legs may import this module; the product path
(chatglm_q_b200/fused_decode.py -> include/cgq.h) never does.

Parity is pinned: tests/golden/decode_tiny.npz holds inputs and logits produced by the REAL reference
model (imported from /root/reference, CPU, float16) by tests/golden/make_golden_decode.py;
tests/test_oracle_golden.py checks this module against them.

All arrays are float32 holding values exactly representable in `dtype` (see qmatmul_oracle.round_to).
"""
from __future__ import annotations

import numpy as np

from . import qmatmul_oracle as orc


def rmsnorm(x: np.ndarray, w: np.ndarray, eps: float, dtype: str) -> np.ndarray:
    """RMSNorm.forward (model.py:68-73): `_norm(x.float()).type_as(x) * weight` — two roundings."""
    x = x.astype(np.float32)
    rstd = np.float32(1.0) / np.sqrt(np.mean(x * x, axis=-1, keepdims=True, dtype=np.float32) + np.float32(eps))
    return orc.round_to(orc.round_to(x * rstd, dtype) * w, dtype)


def silu_gate(u: np.ndarray, dtype: str) -> np.ndarray:
    """GatedFeedForward.forward (model.py:200-201): `act_fn(h) * gate`, h | gate = split(w_in(x))."""
    inner = u.shape[-1] // 2
    h, gate = u[..., :inner].astype(np.float32), u[..., inner:]
    act = orc.round_to(h / (np.float32(1.0) + np.exp(-h)), dtype)        # F.silu computes in fp32
    return orc.round_to(act * gate, dtype)


def rope(x: np.ndarray, freqs_row: np.ndarray, dtype: str) -> np.ndarray:
    """apply_rotary_emb (model.py:47-59) on [..., d_head]: complex product of (x[2j], x[2j+1]) with
    (freqs[2j], freqs[2j+1]); torch multiplies complex-half in fp32 and rounds once."""
    a, b = x[..., 0::2].astype(np.float32), x[..., 1::2].astype(np.float32)
    c, s = freqs_row[0::2].astype(np.float32), freqs_row[1::2].astype(np.float32)
    out = np.empty_like(x, dtype=np.float32)
    out[..., 0::2] = a * c - b * s
    out[..., 1::2] = a * s + b * c
    return orc.round_to(out, dtype)


def attention_decode(qkv: np.ndarray, freqs_row: np.ndarray, k_cache: np.ndarray, v_cache: np.ndarray,
                     n_head: int, n_groups: int, d_head: int, dtype: str):
    """ChatGLM2Attention.forward between qkv_proj and o_proj for one query row (model.py:140-174).
    k_cache / v_cache: [n_past, n_groups, d_head].  Returns (out [n_head*d_head], k_new, v_new)."""
    q = qkv[: n_head * d_head].reshape(n_head, d_head)
    k = qkv[n_head * d_head: (n_head + n_groups) * d_head].reshape(n_groups, d_head)
    v = qkv[(n_head + n_groups) * d_head:].reshape(n_groups, d_head)
    q = rope(q, freqs_row, dtype)
    k = rope(k, freqs_row, dtype)
    keys = np.concatenate([k_cache, k[None]], axis=0)          # [L, g, d]
    vals = np.concatenate([v_cache, v[None]], axis=0)
    q = orc.round_to(q * (np.float32(1.0) / np.float32(np.sqrt(d_head))), dtype)   # q / sqrt(d) via the reciprocal
    hpg = n_head // n_groups
    out = np.empty((n_head, d_head), dtype=np.float32)
    for h in range(n_head):
        g = h // hpg                                            # q.view(.., n_groups, n_head//n_groups, ..), :144
        s = orc.round_to(keys[:, g, :].astype(np.float32) @ q[h], dtype)            # matmul -> dtype, :164
        e = np.exp(s - s.max(), dtype=np.float32)
        p = orc.round_to(e / e.sum(dtype=np.float32), dtype)                        # softmax fp32 -> dtype, :168
        out[h] = orc.round_to(p @ vals[:, g, :].astype(np.float32), dtype)          # :171
    return out.reshape(-1), k, v


def decode_step(w: dict, token: int, kv: list, cfg: dict, dtype: str = "float16"):
    """One token through the tiny int4 model described by `w` (see make_golden_decode.py for the keys).
    kv: list of (k_cache [n_past, g, d], v_cache) per layer.  Returns (logits [V], new kv)."""
    n_head, n_groups, d_head, eps = cfg["n_head"], cfg["n_groups"], cfg["d_head"], cfg["eps"]
    n_past = kv[0][0].shape[0]
    freqs_row = w["freqs"][n_past + 1]                          # position ids are 1-based, model.py:296-297
    x = orc.qembedding_int4(np.array([token]), w["emb_w"], w["emb_s"], dtype)[0]
    new_kv = []
    for i in range(cfg["n_layers"]):
        p = f"l{i}_"
        h = rmsnorm(x, w[p + "attn_ln"], eps, dtype)
        qkv = orc.qmatmul_int4(h[None], w[p + "qkv_w"], w[p + "qkv_s"], w[p + "qkv_b"], dtype)[0]
        ao, k_new, v_new = attention_decode(qkv, freqs_row, kv[i][0], kv[i][1], n_head, n_groups, d_head, dtype)
        new_kv.append((np.concatenate([kv[i][0], k_new[None]]), np.concatenate([kv[i][1], v_new[None]])))
        o = orc.qmatmul_int4(ao[None], w[p + "o_w"], w[p + "o_s"], None, dtype)[0]
        x = orc.round_to(x + o, dtype)                           # x = x + h, model.py:243
        h = rmsnorm(x, w[p + "ffn_ln"], eps, dtype)
        u = orc.qmatmul_int4(h[None], w[p + "win_w"], w[p + "win_s"], None, dtype)[0]
        d = orc.qmatmul_int4(silu_gate(u, dtype)[None], w[p + "wout_w"], w[p + "wout_s"], None, dtype)[0]
        x = orc.round_to(x + d, dtype)                           # model.py:246
    h = rmsnorm(x, w["final_ln"], eps, dtype)
    logits = orc.qmatmul_int4(h[None], w["lm_w"], w["lm_s"], None, dtype)[0]
    return logits, new_kv
