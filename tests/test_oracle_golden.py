"""CPU suite: pin the oracle (numpy + C restatements) to vectors produced by the real reference.

Golden vectors: tests/golden/reference_vectors.npz, config1_int8.npz (made by make_golden.py from
/root/reference).  Reference tests mirrored: tests/test_triton_ops_int4.py:11-22 (fp32, 1e-4) and
tests/test_triton_ops.py:9-17 (int8, signed scales, 1e-4)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import c_oracle
from oracle import qmatmul_oracle as orc
from util import load16

ROOT = Path(__file__).resolve().parent.parent
DTYPES = ["float32", "float16", "bfloat16"]


def test_known_answer_vectors(golden):
    # low nibble = even k, value = nibble - 8, scale switches at k = 32
    out = orc.unpack_int4(np.array([[0xA3]], np.uint8), np.array([[0.5]], np.float32), "float32")
    assert out.ravel().tolist() == [-2.5, 1.0]
    assert np.array_equal(out, golden["ka1_out"])
    out = orc.unpack_int4(np.array([[0x00], [0xFF]], np.uint8), np.array([[1.0]], np.float32), "float32")
    assert out.ravel().tolist() == [-8, -8, 7, 7]
    assert np.array_equal(out, golden["ka2_out"])
    out = orc.unpack_int4(golden["ka3_bytes"], golden["ka3_scale"], "float32")
    assert out[30:34].tolist() == [[-6, -8], [-5, -8], [-10, -20], [50, 100]]
    assert np.array_equal(out, golden["ka3_out"])


def test_unpack_i8_bit_exact(golden):
    assert np.array_equal(orc.unpack_int4_i8(golden["unpack_bytes"]), golden["unpack_i8"])
    assert np.array_equal(c_oracle.w4_unpack_i8(golden["unpack_bytes"]), golden["unpack_i8"])


@pytest.mark.parametrize("dtype", DTYPES)
def test_unpack_int4_bit_exact(golden, dtype):
    scale = load16(golden[f"unpack_scale_{dtype}"], dtype)
    want = load16(golden[f"unpack_out_{dtype}"], dtype)
    got = orc.unpack_int4(golden["unpack_bytes"], scale, dtype)
    assert np.array_equal(got, want)
    assert np.array_equal(c_oracle.w4_dequant(golden["unpack_bytes"], scale, dtype), want)


def test_quantize_int4_matches_reference(golden):
    q, s = orc.quantize_int4(golden["q4_in"])
    assert np.array_equal(q, golden["q4_bytes"])
    assert np.array_equal(s, golden["q4_scale"])
    # nibbles in [1, 15]; an all-zero group packs to 0x88 with the clamped scale
    nib = np.concatenate([(q & 0xF).ravel(), (q >> 4).ravel()])
    assert nib.min() >= 1 and nib.max() <= 15
    assert (q[16:32, 3] == 0x88).all() and s[1, 3] == np.float32(1e-10)
    assert orc.round_to(s[1:2, 3:4], "float16")[0, 0] == 0.0


def test_quantize_int8_matches_reference(golden):
    q, s = orc.quantize_int8(golden["q8_in"])
    assert np.array_equal(q, golden["q8_q"])
    assert np.array_equal(s, golden["q8_scale"])


@pytest.mark.parametrize("dtype", DTYPES)
def test_int4_linear_forward(golden, dtype):
    x = load16(golden[f"l4_x_{dtype}"], dtype)
    s = load16(golden[f"l4_scale_{dtype}"], dtype)
    b = load16(golden[f"l4_bias_{dtype}"], dtype)
    want = load16(golden[f"l4_y_{dtype}"], dtype)
    got = orc.qmatmul_int4(x, golden["l4_bytes"], s, b, dtype)
    got_c = c_oracle.w4a16_gemm(x, golden["l4_bytes"], s, b, dtype)
    # torch's CPU matmul may order the K=128 sum differently: allow one rounding step of the dtype
    tol = {"float32": 2e-6, "float16": 2e-3, "bfloat16": 1.6e-2}[dtype]
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol)
    np.testing.assert_allclose(got_c, want, rtol=tol, atol=tol)


def test_int4_reference_test_shape(golden):
    # tests/test_triton_ops_int4.py:20-22 criterion: atol = rtol = 1e-4 in fp32
    got = orc.qmatmul_int4(golden["t4_a"], golden["t4_bytes"], golden["t4_scale"], None, "float32")
    assert np.allclose(got, golden["t4_y"], atol=1e-4, rtol=1e-4)
    got_c = c_oracle.w4a16_gemm(golden["t4_a"], golden["t4_bytes"], golden["t4_scale"], None, "float32")
    assert np.allclose(got_c, golden["t4_y"], atol=1e-4, rtol=1e-4)


def test_int8_reference_test_shape(golden):
    # tests/test_triton_ops.py:14-17: A @ (B * B_scale), signed scales, 1e-4
    w_nk = np.ascontiguousarray(golden["t8_b_kn"].T)
    got = orc.qmatmul_int8(golden["t8_a"], w_nk, golden["t8_scale"], None, "float32")
    assert np.allclose(got, golden["t8_y"], atol=1e-4, rtol=1e-4)
    got_c = c_oracle.w8a16_gemm(golden["t8_a"], w_nk, golden["t8_scale"], None, "float32")
    assert np.allclose(got_c, golden["t8_y"], atol=1e-4, rtol=1e-4)


@pytest.mark.parametrize("dtype", DTYPES)
def test_int8_linear_forward(golden, dtype):
    x = load16(golden[f"l8_x_{dtype}"], dtype)
    s = load16(golden[f"l8_scale_{dtype}"], dtype)
    b = load16(golden[f"l8_bias_{dtype}"], dtype)
    want = load16(golden[f"l8_y_{dtype}"], dtype)
    tol = {"float32": 2e-6, "float16": 2e-3, "bfloat16": 1.6e-2}[dtype]
    np.testing.assert_allclose(orc.qmatmul_int8(x, golden["l8_q"], s, b, dtype), want, rtol=tol, atol=tol)
    np.testing.assert_allclose(c_oracle.w8a16_gemm(x, golden["l8_q"], s, b, dtype), want, rtol=tol, atol=tol)


def test_qembedding(golden):
    s4 = golden["e4_scale"].astype(np.float32)
    got = orc.qembedding_int4(golden["e4_ids"], golden["e4_bytes"], s4, "float16")
    assert np.array_equal(got, golden["e4_y"].astype(np.float32))
    # QEmbedding(ids) == unpack_int4(W, S)[ids]
    assert np.array_equal(got, orc.unpack_int4(golden["e4_bytes"], s4, "float16")[golden["e4_ids"]])
    got8 = orc.qembedding_int8(golden["e4_ids"], golden["e8_q"], golden["e8_scale"].astype(np.float32), "float16")
    assert np.array_equal(got8, golden["e8_y"].astype(np.float32))


def test_backward_grad_a_matches_reference_autograd():
    """grad_A oracles against autograd through the real reference's torch path (tests/golden/make_golden_backward.py);
    fp32, the reference tests' own 1e-4 criterion (tests/test_triton_ops_int4.py:24-37)."""
    g = np.load(ROOT / "tests" / "golden" / "backward.npz")
    for tag in ("a", "b"):
        got = orc.qmatmul_int4_grad_a(g[f"s4_{tag}_grad_out"], g[f"s4_{tag}_b"], g[f"s4_{tag}_scale"], "float32")
        want = g[f"s4_{tag}_grad_a"]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
        got = orc.qmatmul_int8_grad_a(g[f"s4_{tag}_grad_out"], g[f"s8_{tag}_w"], g[f"s8_{tag}_scale"], "float32")
        want = g[f"s8_{tag}_grad_a"]
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4 * np.abs(want).max()


def test_config1_int8_plumbing():
    """BASELINE.json configs[0]: int8 QLinear forward (128,4096)x(4096,4096) on the CPU path."""
    fx = dict(np.load(ROOT / "tests" / "golden" / "config1_int8.npz"))
    assert bool(fx["exact"])  # reference: forward(x) == x @ (q.t() * s) exactly
    r1 = np.random.default_rng(1)
    W = (r1.standard_normal((4096, 4096)) / 64).astype(np.float32)
    X = r1.standard_normal((128, 4096)).astype(np.float32)
    q, s = orc.quantize_int8(W)
    assert np.array_equal(q[::64, ::64], fx["q_sub"]) and np.array_equal(s[::64], fx["s_sub"])
    y = c_oracle.w8a16_gemm(X, q, s, None, "float32")
    np.testing.assert_allclose(y[::16, ::64], fx["y_sub"], rtol=1e-4, atol=1e-4)
    assert abs(float(y.astype(np.float64).sum()) - float(fx["y_checksum"])) < 1e-2 * 128
    relerr = np.linalg.norm(y - X @ W.T) / np.linalg.norm(X @ W.T)
    assert abs(relerr - float(fx["relerr"])) < 1e-4  # ~8.7e-3


def test_byte_accounting():
    # BASELINE.md §3 table
    assert abs(orc.algorithmic_bytes("int4", 1, 4608, 4096) / 1e6 - 10.634) < 1e-3
    assert abs(orc.algorithmic_bytes("int4", 1, 27392, 4096) / 1e6 - 63.174) < 1e-3
    assert abs(orc.algorithmic_bytes("int4", 2048, 13696, 4096) / 1e6 - 104.432) < 1e-3
    assert orc.flops(128, 4608, 4096) == 2 * 128 * 4608 * 4096


def test_decode_step_oracle_matches_reference_model():
    """oracle/decode_oracle.py against the REAL reference model (CPU fp16, tests/golden/make_golden_decode.py):
    prefill KV -> greedy decode steps; logits within the parity bar, same greedy tokens, KV rows equal up to
    one fp16 ulp (the reference rounds the same fp32 sums)."""
    from oracle import decode_oracle as dec
    from util import assert_parity, load_decode_golden

    w, cfg, fx = load_decode_golden()
    kv = [(fx[f"prefill_k{i}"].astype(np.float32), fx[f"prefill_v{i}"].astype(np.float32))
          for i in range(cfg["n_layers"])]
    for step, tok in enumerate(fx["step_tokens"]):
        logits, kv = dec.decode_step(w, int(tok), kv, cfg, "float16")
        ref = fx["step_logits"][step].astype(np.float32)
        assert_parity(logits, ref, f"decode oracle step {step}", rtol=2e-2)
        assert int(logits.argmax()) == int(ref.argmax())
    for i in range(cfg["n_layers"]):
        assert_parity(kv[i][0], fx[f"final_k{i}"].astype(np.float32), f"k cache layer {i}", rtol=2e-2)
        assert_parity(kv[i][1], fx[f"final_v{i}"].astype(np.float32), f"v cache layer {i}", rtol=2e-2)


# ---------------------------------------------------------------------- sampler
def _sampling_cases():
    fx = dict(np.load(Path(__file__).resolve().parent / "golden" / "sampling.npz"))
    for entry in fx["cases"]:
        name, dtype = str(entry).split(":")
        bits = fx[f"{name}_logits_bits"]
        logits = orc.bf16_from_bits(bits) if dtype == "bfloat16" else bits.view(np.float16).astype(np.float32)
        top_k, top_p, temp = fx[f"{name}_params"]
        yield name, dtype, logits, int(top_k), float(top_p), float(temp), fx


def test_sampling_oracle_matches_reference_distribution_and_token():
    """oracle/sampling_oracle.py against what the unmodified chatglm_q.decoder.top_p_sampling handed to
    torch.multinomial / torch.gather (tests/golden/make_golden_sampling.py), and against the token it returned."""
    from oracle import sampling_oracle as so

    n = 0
    for name, _, logits, top_k, top_p, temp, fx in _sampling_cases():
        p, idx = so.top_p_distribution(logits, top_k, top_p, temp)
        ref_p, ref_idx = fx[f"{name}_probs"], fx[f"{name}_indices"]
        assert p.shape == ref_p.shape, name
        np.testing.assert_allclose(p, ref_p, rtol=2e-6, atol=1e-9, err_msg=name)
        # the same probabilities at every rank; the same ids wherever the probability is not tied
        assert np.array_equal(logits[idx], logits[ref_idx]), name
        untied = np.ones(len(idx), bool)
        lv = logits[ref_idx]
        untied[1:] &= lv[1:] != lv[:-1]
        untied[:-1] &= lv[:-1] != lv[1:]
        edge = logits[ref_idx[-1]]
        untied &= lv != edge                      # the boundary value may have more holders than slots
        assert np.array_equal(idx[untied], ref_idx[untied]), name
        assert so.sample_with(ref_p, ref_idx, fx[f"{name}_q"]) == int(fx[f"{name}_token"]), name
        n += 1
    assert n >= 7
