"""Shared helpers for the parity tests (numpy <-> torch conversions, the parity bar)."""
from __future__ import annotations

import numpy as np

from oracle import qmatmul_oracle as orc

# Parity bar for the matmul (BASELINE.json north_star "fp16 accum within 1e-2 rel", made testable in
# SURVEY.md §8(c)):  |got - ref| <= RTOL * |ref| + RTOL * rms(ref)
RTOL = 1e-2
# bfloat16 keeps 8 significant bits: ONE ulp is 2^-8..2^-7 = 0.39..0.78 % of the value, and the
# reference's two roundings (product, then `out += bias`) can each flip by one ulp when the fp32
# sums differ in the last bits.  The 1e-2 bar of the north star is stated for fp16; bf16 outputs
# are held to 2e-2 (about 2.5 bf16 ulps).
RTOL_BF16 = 2e-2


def rtol_for(dtype: str) -> float:
    return RTOL_BF16 if dtype == "bfloat16" else RTOL


def load16(arr: np.ndarray, dtype: str) -> np.ndarray:
    """Fixture array -> float32 values (bf16 fixtures are uint16 bit patterns)."""
    if dtype == "bfloat16":
        return orc.bf16_from_bits(arr)
    return arr.astype(np.float32)


def assert_parity(got: np.ndarray, ref: np.ndarray, what: str = "", rtol: float = RTOL):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, f"{what}: shape {got.shape} vs {ref.shape}"
    assert np.isfinite(got).all(), f"{what}: non-finite output"
    rms = float(np.sqrt(np.mean(ref * ref))) if ref.size else 0.0
    err = np.abs(got - ref)
    bound = rtol * np.abs(ref) + rtol * rms
    bad = err > bound
    assert not bad.any(), (
        f"{what}: {int(bad.sum())}/{bad.size} elements outside rtol={rtol} "
        f"(max err {err.max():.4g}, rms(ref) {rms:.4g}, worst ratio {(err / np.maximum(bound, 1e-30)).max():.3g})")


def to_torch(x: np.ndarray, dtype: str, device="cuda"):
    import torch

    td = {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32}[dtype]
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(device=device, dtype=td)


def from_torch(t) -> np.ndarray:
    return t.detach().float().cpu().numpy()


def make_int4_case(seed: int, m: int, k: int, n: int, kind: str = "Q", dtype: str = "float16"):
    """SURVEY §8(d) config-2 generators: set Q = quantiser-shaped, set R = adversarial random bytes
    (covers nibble 0 => -8, signed and tiny scales, one zero-scale column block)."""
    rng = np.random.default_rng(seed)
    a = orc.round_to(rng.standard_normal((m, k)), dtype)
    if kind == "Q":
        w = (rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32)
        bq, s = orc.quantize_int4(w)
    else:
        bq = rng.integers(0, 256, size=(k // 2, n), dtype=np.uint8)
        s = (rng.random((k // 32, n)) * 0.02 - 0.01).astype(np.float32)
        s[:, : min(16, n)] = 0.0
    return a, bq, orc.round_to(s, dtype)


def make_int8_case(seed: int, m: int, k: int, n: int, kind: str = "Q", dtype: str = "float16"):
    rng = np.random.default_rng(seed)
    a = orc.round_to(rng.standard_normal((m, k)), dtype)
    if kind == "Q":
        w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
        q, s = orc.quantize_int8(w)
    else:
        q = rng.integers(-128, 128, size=(n, k), dtype=np.int8)
        s = (rng.standard_normal(n) / 256 / 8).astype(np.float32)  # signed (tests/test_triton_ops.py:12)
    return a, q, orc.round_to(s, dtype)


def load_decode_golden():
    """tests/golden/decode_tiny.npz (reference model, CPU fp16) -> (weights dict, cfg dict, fixture)."""
    from pathlib import Path

    fx = dict(np.load(Path(__file__).resolve().parent / "golden" / "decode_tiny.npz"))
    n_head, n_groups, d_head, n_layers, hidden, inner, vocab, max_seq = (int(v) for v in fx["cfg"])
    cfg = dict(n_head=n_head, n_groups=n_groups, d_head=d_head, n_layers=n_layers, hidden=hidden, inner=inner,
               vocab=vocab, max_seq=max_seq, eps=float(fx["eps"]))
    w = {}
    for k, v in fx.items():
        if k.endswith(("_w",)):
            w[k] = v
        elif k.endswith(("_s", "_b", "_ln")) or k in ("final_ln", "freqs"):
            w[k] = v.astype(np.float32)
    for i in range(n_layers):
        w.setdefault(f"l{i}_qkv_b", None)
    return w, cfg, fx
