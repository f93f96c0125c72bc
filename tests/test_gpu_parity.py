"""GPU suite (-m gpu): the CUDA path, called through the C-ABI (chatglm_q_b200.ops -> libcgq.so),
against the oracle on identical seeded inputs, the committed golden fixtures from the real
reference, and size-independent properties at BASELINE.json's full sizes.

Parity bar: integer unpack and the dequantised weight are BIT-EXACT; the matmul is within
1e-2 (util.assert_parity: |got-ref| <= 1e-2*|ref| + 1e-2*rms(ref)), BASELINE.json north_star.
"""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import qmatmul_oracle as orc
from util import (assert_parity, rtol_for, from_torch, load16, make_int4_case, make_int8_case, to_torch)

pytestmark = pytest.mark.gpu

from chatglm_q_b200 import int4 as m4  # noqa: E402
from chatglm_q_b200 import int8 as m8  # noqa: E402
from chatglm_q_b200 import ops  # noqa: E402

DEV = "cuda"
IMPLS4 = {"auto": ops.IMPL_AUTO, "simple": ops.IMPL_SIMPLE, "gemv": ops.IMPL_GEMV,
          "gemv_exact": ops.IMPL_GEMV_EXACT, "tc": ops.IMPL_TC, "umma": ops.IMPL_GEMV_UMMA,
          "gemv_subnormal": ops.IMPL_GEMV_SUBNORMAL, "gemv_imma": ops.IMPL_GEMV_IMMA}


def u8(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


def run4(a, bq, s, dtype, bias=None, impl=ops.IMPL_AUTO):
    out = ops.dynamic_quant_matmul_s4(to_torch(a, dtype), u8(bq), to_torch(s, dtype),
                                      bias=None if bias is None else to_torch(bias, dtype), impl=impl)
    torch.cuda.synchronize()
    return from_torch(out)


def run8(a, q_nk, s, dtype, bias=None, impl=ops.IMPL_AUTO):
    w = u8(q_nk)  # [N, K] module buffer; the op takes the transposed view like the reference forward
    out = ops.dynamic_quant_matmul(to_torch(a, dtype), w.t(), to_torch(s, dtype),
                                   bias=None if bias is None else to_torch(bias, dtype), impl=impl)
    torch.cuda.synchronize()
    return from_torch(out)


# ------------------------------------------------------------------ bit-exact pieces
def test_unpack_i8_bit_exact(golden):
    got = ops.unpack_int4_i8(u8(golden["unpack_bytes"])).cpu().numpy()
    assert np.array_equal(got, golden["unpack_i8"])
    rng = np.random.default_rng(7)
    b = rng.integers(0, 256, size=(2048, 272), dtype=np.uint8)
    assert np.array_equal(ops.unpack_int4_i8(u8(b)).cpu().numpy(), orc.unpack_int4_i8(b))


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_dequant_bit_exact(golden, dtype):
    scale = load16(golden[f"unpack_scale_{dtype}"], dtype)
    want = load16(golden[f"unpack_out_{dtype}"], dtype)
    got = from_torch(ops.unpack_int4(u8(golden["unpack_bytes"]), to_torch(scale, dtype)))
    assert np.array_equal(got, want)
    # random bytes (2048, 64), includes nibble 0 and tiny / negative scales
    rng = np.random.default_rng(11)
    b = rng.integers(0, 256, size=(2048, 64), dtype=np.uint8)
    s = orc.round_to(rng.standard_normal((128, 64)) * 0.01, dtype)
    got = from_torch(ops.unpack_int4(u8(b), to_torch(s, dtype)))
    assert np.array_equal(got, orc.unpack_int4(b, s, dtype))


def test_embeddings_bit_exact(golden):
    ids = torch.from_numpy(golden["e4_ids"]).to(DEV)
    got = from_torch(ops.embedding_s4(ids, u8(golden["e4_bytes"]),
                                      torch.from_numpy(golden["e4_scale"]).to(DEV)))
    assert np.array_equal(got, golden["e4_y"].astype(np.float32))
    got8 = from_torch(ops.embedding_s8(ids, u8(golden["e8_q"]), torch.from_numpy(golden["e8_scale"]).to(DEV)))
    assert np.array_equal(got8, golden["e8_y"].astype(np.float32))


# ------------------------------------------------------------------ golden fixtures from the reference
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("impl", ["auto", "simple", "gemv_exact"])
def test_int4_linear_golden(golden, dtype, impl):
    x = load16(golden[f"l4_x_{dtype}"], dtype)
    s = load16(golden[f"l4_scale_{dtype}"], dtype)
    b = load16(golden[f"l4_bias_{dtype}"], dtype)
    want = load16(golden[f"l4_y_{dtype}"], dtype)
    got = run4(x, golden["l4_bytes"], s, dtype, bias=b, impl=IMPLS4[impl])
    assert_parity(got, want, f"int4 golden {dtype} {impl}", rtol=rtol_for(dtype))


def test_int4_reference_test_shape(golden):
    """tests/test_triton_ops_int4.py:11-22 inputs, run in fp16 (the kernels take 16-bit activations)."""
    a = orc.round_to(golden["t4_a"], "float16")
    s = orc.round_to(golden["t4_scale"], "float16")
    got = run4(a, golden["t4_bytes"], s, "float16")
    assert_parity(got, orc.qmatmul_int4(a, golden["t4_bytes"], s, None, "float16"), "int4 ref-test shape")
    # and against the fp32 result the reference test itself expects, at the 16-bit parity bar
    assert_parity(got, golden["t4_y"], "int4 ref-test shape vs fp32 golden", rtol=2e-2)


def test_int8_reference_test_shape(golden):
    """tests/test_triton_ops.py:9-17: M=10 (ragged row block), signed scales."""
    a = orc.round_to(golden["t8_a"], "float16")
    s = orc.round_to(golden["t8_scale"], "float16")
    w_nk = np.ascontiguousarray(golden["t8_b_kn"].T)
    got = run8(a, w_nk, s, "float16")
    assert_parity(got, orc.qmatmul_int8(a, w_nk, s, None, "float16"), "int8 ref-test shape")


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_int8_linear_golden(golden, dtype):
    x = load16(golden[f"l8_x_{dtype}"], dtype)
    s = load16(golden[f"l8_scale_{dtype}"], dtype)
    b = load16(golden[f"l8_bias_{dtype}"], dtype)
    want = load16(golden[f"l8_y_{dtype}"], dtype)
    for impl in (ops.IMPL_AUTO, ops.IMPL_SIMPLE):
        assert_parity(run8(x, golden["l8_q"], s, dtype, bias=b, impl=impl), want, f"int8 golden {dtype}", rtol=rtol_for(dtype))


# ------------------------------------------------------------------ decode kernels vs oracle, real shapes
SHAPES4 = [(4096, 4608), (4096, 4096), (13696, 4096), (4096, 1280), (512, 256), (4096, 6848)]


@pytest.mark.parametrize("impl", ["gemv", "gemv_exact", "gemv_subnormal", "gemv_imma", "simple"])
@pytest.mark.parametrize("kind", ["Q", "R"])
@pytest.mark.parametrize("m", [1, 2, 5, 8])
def test_int4_decode_shapes(impl, kind, m):
    for (k, n) in SHAPES4:
        a, bq, s = make_int4_case(1234 + 1000 * m + n, m, k, n, kind)
        want = c_oracle.w4a16_gemm(a, bq, s, None, "float16")
        got = run4(a, bq, s, "float16", impl=IMPLS4[impl])
        assert_parity(got, want, f"int4 {impl} {kind} M={m} K={k} N={n}")


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("kind", ["Q", "R"])
def test_int4_decode_umma(kind, dtype):
    """Integer-tcgen05 batch-1 decode kernel (activation as three int8 digits, exact int32 group sums)."""
    for (k, n, with_bias) in [(4096, 4608, True), (4096, 4096, False), (13696, 4096, False), (512, 256, False),
                              (4096, 27392, False), (4096, 1296, True)]:
        a, bq, s = make_int4_case(77 + n + k, 1, k, n, kind, dtype)
        bias = orc.round_to(np.random.default_rng(n).standard_normal(n) * 0.1, dtype) if with_bias else None
        want = c_oracle.w4a16_gemm(a, bq, s, bias, dtype)
        got = run4(a, bq, s, dtype, bias=bias, impl=ops.IMPL_GEMV_UMMA)
        assert_parity(got, want, f"int4 umma {kind} {dtype} K={k} N={n}", rtol=rtol_for(dtype))
    # wide dynamic range inside a group (digits keep 21 bits below the group maximum) and a one-hot row
    k, n = 4096, 512
    a, bq, s = make_int4_case(5, 1, k, n, "Q", dtype)
    a[0, ::7] *= 1e-3
    a[0, 5::64] *= 300.0
    a = orc.round_to(a, dtype)
    assert_parity(run4(a, bq, s, dtype, impl=ops.IMPL_GEMV_UMMA), c_oracle.w4a16_gemm(a, bq, s, None, dtype),
                  f"int4 umma dynamic range {dtype}", rtol=rtol_for(dtype))
    e = np.zeros((1, k), dtype=np.float32)
    e[0, 1234] = 1.0
    got = run4(e, bq, s, dtype, impl=ops.IMPL_GEMV_UMMA)
    assert np.array_equal(got, orc.unpack_int4(bq, s, dtype)[1234][None, :])


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_int4_decode_big_shapes(dtype):
    """w_in (N=27392) and a 1/8 slice of lm_head width, bias on the qkv shape, bf16 too."""
    for (m, k, n, with_bias) in [(1, 4096, 27392, False), (1, 4096, 4608, True), (8, 4096, 8128, True)]:
        a, bq, s = make_int4_case(99 + n + m, m, k, n, "Q", dtype)
        bias = orc.round_to(np.random.default_rng(n).standard_normal(n) * 0.1, dtype) if with_bias else None
        want = c_oracle.w4a16_gemm(a, bq, s, bias, dtype)
        assert_parity(run4(a, bq, s, dtype, bias=bias), want, f"int4 big {dtype} M={m} N={n}", rtol=rtol_for(dtype))


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_int4_imma_digits_are_fp32_exact(dtype):
    """The integer-MMA arithmetic (base-256 digits of the activation) against the fp64 value of the
    same sum of products, BEFORE the final rounding matters: activations spanning 20 binades inside one quantisation
    group (outliers next to tiny values), a subnormal-only group, an all-zero group, a ragged last column tile.  The
    result must equal the correctly rounded fp64 sum up to 1 ulp of the 16-bit output (fp32 accumulation error only)."""
    k, n = 4096, 4608 + 16
    rng = np.random.default_rng(77)
    a = rng.standard_normal((1, k)) * np.exp2(rng.integers(-14, 6, size=(1, k)))
    a[0, 64:96] = 0.0                                   # all-zero group
    a[0, 96:128] = rng.integers(-3, 4, size=32) * 2.0 ** -24 if dtype == "float16" else 1e-30   # subnormals / vanishing
    a[0, 1000] = 3.0e4                                  # outlier: 2^15 next to 2^-14 in one group
    a = orc.round_to(a, dtype)
    _, bq, s = make_int4_case(78, 1, k, n, "R", dtype)
    w = orc.unpack_int4_i8(bq).astype(np.float64).reshape(k // 32, 32, n) * s.astype(np.float64)[:, None, :]
    exact = a.astype(np.float64) @ w.reshape(k, n)
    for impl in ("gemv_imma", "auto"):
        got = run4(a, bq, s, dtype, impl=IMPLS4[impl]).astype(np.float64)
        ulp = np.maximum(np.abs(exact), 2.0 ** -14) * (2.0 ** -10 if dtype == "float16" else 2.0 ** -7)
        assert (np.abs(got - exact) <= 0.5001 * ulp + 2e-6 * np.abs(a).astype(np.float64) @ np.abs(w.reshape(k, n))).all(), impl
    # non-finite activations poison the output like the reference's matmul does (never a finite garbage value)
    a2 = a.copy()
    a2[0, 5] = np.inf
    assert not np.isfinite(run4(a2, bq, s, dtype, impl=IMPLS4["gemv_imma"])).any()
    a2[0, 5] = np.nan
    assert np.isnan(run4(a2, bq, s, dtype, impl=IMPLS4["gemv_imma"])).all()


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_int4_imma_ragged_shapes(dtype):
    """One-token integer-MMA kernel on ragged geometry: K from one quantisation group up to several k-stages with a
    partial last stage, N from one 16-column block to ragged last tiles, bias on every other case."""
    rng = np.random.default_rng(5)
    for k in (32, 64, 96, 160, 384, 416, 1056):
        for n in (16, 48, 144, 272, 400):
            a, bq, s = make_int4_case(1000 + k + n, 1, k, n, "R", dtype)
            bias = orc.round_to(rng.standard_normal(n) * 0.1, dtype) if (k + n) % 64 == 0 else None
            got = run4(a, bq, s, dtype, bias=bias, impl=IMPLS4["gemv_imma"])
            assert_parity(got, c_oracle.w4a16_gemm(a, bq, s, bias, dtype), f"imma {dtype} K={k} N={n}", rtol=rtol_for(dtype))


@pytest.mark.parametrize("m", [2, 3, 5, 8])
def test_int8_mx_ragged_shapes(m):
    """w8_gemv_mx_kernel (2 .. 8 tokens, activations in the TMA ring): ragged K / N, strided activation rows, bias."""
    rng = np.random.default_rng(6)
    for (k, n) in [(128, 64), (136, 80), (1000, 200), (4096, 1000), (13696, 4096), (272, 4608)]:
        a, q, s = make_int8_case(2000 + k + n + m, m, k, n, "R")
        bias = orc.round_to(rng.standard_normal(n) * 0.1, "float16") if n % 3 == 0 else None
        want = c_oracle.w8a16_gemm(a, q, s, bias, "float16")
        assert_parity(run8(a, q, s, "float16", bias=bias), want, f"int8 mx M={m} K={k} N={n}")
    # row-strided activation (lda > K): the tensor map carries the stride
    k, n = 512, 256
    a, q, s = make_int8_case(77, m, k, n, "Q")
    wide = torch.zeros((m, k + 64), dtype=torch.float16, device=DEV)
    wide[:, :k] = to_torch(a, "float16")
    got = ops.dynamic_quant_matmul(wide[:, :k], u8(q).t(), to_torch(s, "float16"))
    assert_parity(from_torch(got), c_oracle.w8a16_gemm(a, q, s, None, "float16"), f"int8 mx strided M={m}")


@pytest.mark.parametrize("kind", ["Q", "R"])
@pytest.mark.parametrize("m", [1, 3, 8])
def test_int8_decode_shapes(kind, m):
    for (k, n) in [(4096, 4608), (4096, 4096), (13696, 4096), (128, 256), (4096, 1000)]:
        a, q, s = make_int8_case(4321 + 1000 * m + n, m, k, n, kind)
        want = c_oracle.w8a16_gemm(a, q, s, None, "float16")
        for impl in (ops.IMPL_AUTO, ops.IMPL_SIMPLE):
            assert_parity(run8(a, q, s, "float16", impl=impl), want, f"int8 {kind} M={m} K={k} N={n} impl={impl}")


def test_int8_bf16_and_bias():
    a, q, s = make_int8_case(5, 4, 4096, 4608, "Q", "bfloat16")
    bias = orc.round_to(np.random.default_rng(3).standard_normal(4608) * 0.1, "bfloat16")
    assert_parity(run8(a, q, s, "bfloat16", bias=bias), c_oracle.w8a16_gemm(a, q, s, bias, "bfloat16"), "int8 bf16",
                  rtol=rtol_for("bfloat16"))


# ------------------------------------------------------------------ prefill-sized M (M > 8)
@pytest.mark.parametrize("m", [9, 32, 128, 300])
def test_int4_prefill_m(m):
    k, n = 4096, 1280
    a, bq, s = make_int4_case(77 + m, m, k, n, "Q")
    assert_parity(run4(a, bq, s, "float16"), c_oracle.w4a16_gemm(a, bq, s, None, "float16"), f"int4 M={m}")


@pytest.mark.parametrize("m", [10, 128])
def test_int8_prefill_m(m):
    k, n = 4096, 768
    a, q, s = make_int8_case(78 + m, m, k, n, "Q")
    assert_parity(run8(a, q, s, "float16"), c_oracle.w8a16_gemm(a, q, s, None, "float16"), f"int8 M={m}")


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("kind", ["Q", "R"])
def test_int4_tcgen05_shapes(kind, dtype):
    """tcgen05 prefill kernel: every token-block size (MB 32/64/128/256), ragged M and N tails,
    K = 13696 (214 k-stages), bias, and M <= 8 forced through the tensor-core path."""
    for (m, k, n, with_bias) in [(9, 4096, 1280, False), (33, 512, 256, True), (100, 4096, 4608, True),
                                 (128, 13696, 4096, False), (300, 4096, 1296, True), (5, 4096, 512, False),
                                 (520, 1024, 2064, False)]:
        a, bq, s = make_int4_case(500 + m + n, m, k, n, kind, dtype)
        bias = orc.round_to(np.random.default_rng(n).standard_normal(n) * 0.1, dtype) if with_bias else None
        want = c_oracle.w4a16_gemm(a, bq, s, bias, dtype)
        got = run4(a, bq, s, dtype, bias=bias, impl=ops.IMPL_TC)
        assert_parity(got, want, f"int4 tc {kind} {dtype} M={m} K={k} N={n}", rtol=rtol_for(dtype))


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("kind", ["Q", "R"])
def test_int8_tcgen05_shapes(kind, dtype):
    for (m, k, n, with_bias) in [(10, 128, 256, False), (40, 4096, 1000, True), (128, 13696, 4096, False),
                                 (300, 4096, 4608, True), (3, 4096, 512, False), (700, 1040, 2048, False)]:
        a, q, s = make_int8_case(600 + m + n, m, k, n, kind, dtype)
        bias = orc.round_to(np.random.default_rng(n).standard_normal(n) * 0.1, dtype) if with_bias else None
        want = c_oracle.w8a16_gemm(a, q, s, bias, dtype)
        got = run8(a, q, s, dtype, bias=bias, impl=ops.IMPL_TC)
        assert_parity(got, want, f"int8 tc {kind} {dtype} M={m} K={k} N={n}", rtol=rtol_for(dtype))


def test_int4_tcgen05_matches_simple_bitwise_dequant():
    """The tensor-core path feeds the reference's own rounded weights: against the bit-faithful CUDA-core
    kernel only the fp32 summation order differs, so the outputs agree to ~1 ulp of fp16."""
    a, bq, s = make_int4_case(901, 256, 4096, 4096, "Q")
    y_tc = run4(a, bq, s, "float16", impl=ops.IMPL_TC)
    y_s = run4(a, bq, s, "float16", impl=ops.IMPL_SIMPLE)
    rms = float(np.sqrt(np.mean(y_s * y_s)))
    assert np.abs(y_tc - y_s).max() <= 2e-3 * rms + 2e-3 * np.abs(y_s).max()
    # deterministic
    assert np.array_equal(y_tc, run4(a, bq, s, "float16", impl=ops.IMPL_TC))


# ------------------------------------------------------------------ edge cases
def test_edge_shapes_and_strides():
    # N not a multiple of 16 -> shape-general kernel; leading dims are flattened; empty batch
    a, bq, s = make_int4_case(3, 6, 64, 40, "R")
    want = orc.qmatmul_int4(a, bq, s, None, "float16")
    assert_parity(run4(a, bq, s, "float16"), want, "N=40")
    x = to_torch(a, "float16").reshape(2, 3, 64)
    out = ops.dynamic_quant_matmul_s4(x, u8(bq), to_torch(s, "float16"))
    assert out.shape == (2, 3, 40)
    assert_parity(from_torch(out).reshape(6, 40), want, "leading dims")
    empty = ops.dynamic_quant_matmul_s4(x[:0], u8(bq), to_torch(s, "float16"))
    assert empty.shape == (0, 3, 40)
    # row-strided activations (a slice of a wider tensor), decode kernel
    a2, bq2, s2 = make_int4_case(4, 4, 4096, 512, "Q")
    wide = torch.zeros(4, 8192, dtype=torch.float16, device=DEV)
    wide[:, :4096] = to_torch(a2, "float16")
    got = from_torch(ops.dynamic_quant_matmul_s4(wide[:, :4096], u8(bq2), to_torch(s2, "float16")))
    assert_parity(got, c_oracle.w4a16_gemm(a2, bq2, s2, None, "float16"), "row-strided A")
    # the output is a fresh tensor the caller may modify in place (reference: `out += self.bias`)
    o1 = ops.dynamic_quant_matmul_s4(to_torch(a2, "float16"), u8(bq2), to_torch(s2, "float16"))
    o1 += 1.0


def test_unsupported_inputs_raise():
    a, bq, s = make_int4_case(3, 2, 64, 32, "Q")
    with pytest.raises(TypeError):  # fp32 activations: no fallback exists
        ops.dynamic_quant_matmul_s4(to_torch(a, "float32"), u8(bq), to_torch(s, "float32"))
    with pytest.raises(AssertionError):  # group size 64
        ops.dynamic_quant_matmul_s4(to_torch(a, "float16"), u8(bq), to_torch(s[:1], "float16"))
    with pytest.raises(AssertionError):  # weight on another device (CPU)
        ops.dynamic_quant_matmul_s4(to_torch(a, "float16"), torch.from_numpy(bq), to_torch(s, "float16"))


# ------------------------------------------------------------------ size-independent properties, full sizes
@pytest.mark.parametrize("n", [4608, 13696, 27392, 65024])
def test_properties_full_size_int4(n):
    k = 4096
    g = torch.Generator(device=DEV).manual_seed(n)
    bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
    s = (torch.rand((k // 32, n), device=DEV, generator=g) * 0.02 - 0.01).half()
    a = torch.randn((1, k), device=DEV, generator=g).half()
    y1 = ops.dynamic_quant_matmul_s4(a, bq, s)
    y2 = ops.dynamic_quant_matmul_s4(a, bq, s)
    assert torch.equal(y1, y2), "deterministic reduction: two launches must agree bit-for-bit"
    assert torch.isfinite(y1).all()
    # scaling every scale by 2 doubles the result exactly (power of two) wherever the fp16 result is
    # a normal number (a subnormal result rounds on a fixed grid, so round(2x) != 2*round(x) there)
    normal = y1.abs() >= 2.0 ** -13
    assert torch.equal(ops.dynamic_quant_matmul_s4(a, bq, s * 2)[normal], (y1 * 2)[normal])
    # one-hot activation selects a dequantised weight row bit-exactly
    w = ops.unpack_int4(bq, s)
    for kk in (0, 1, 2047, 4095):
        e = torch.zeros((1, k), dtype=torch.float16, device=DEV)
        e[0, kk] = 1.0
        assert torch.equal(ops.dynamic_quant_matmul_s4(e, bq, s)[0], w[kk]), f"one-hot k={kk}"
    # all nibbles == 8 is the zero weight
    # (exactly with the exact-dequant kernel; the subnormal-trick kernel computes
    #  s*(2^24*sum(a*q') - 8*sum(a)) whose two fp32 sums cancel to ~1e-7 of |a|_1, far inside the bar)
    zero_w = torch.full_like(bq, 0x88)
    assert (ops.dynamic_quant_matmul_s4(a, zero_w, s, impl=ops.IMPL_GEMV_EXACT) == 0).all()
    assert (ops.dynamic_quant_matmul_s4(a, zero_w, s, impl=ops.IMPL_SIMPLE) == 0).all()
    z = ops.dynamic_quant_matmul_s4(a, zero_w, s)
    assert float(z.float().abs().max()) <= 1e-4 * float(y1.float().pow(2).mean().sqrt())
    # fast kernel vs the bit-faithful CUDA-core kernel on the same inputs
    ys = ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
    assert_parity(from_torch(y1), from_torch(ys), f"gemv vs simple N={n}")
    # the repeated-run workspace stays clean: a different shape right after still agrees
    a8 = torch.randn((8, k), device=DEV, generator=g).half()
    y8 = ops.dynamic_quant_matmul_s4(a8, bq, s)
    assert_parity(from_torch(y8), from_torch(ops.dynamic_quant_matmul_s4(a8, bq, s, impl=ops.IMPL_SIMPLE)),
                  f"M=8 N={n}")


def test_properties_full_size_int8():
    k, n = 4096, 27392
    g = torch.Generator(device=DEV).manual_seed(8)
    w = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=DEV, generator=g)
    s = (torch.randn(n, device=DEV, generator=g) / 2048).half()
    a = torch.randn((1, k), device=DEV, generator=g).half()
    y1 = ops.dynamic_quant_matmul(a, w.t(), s)
    assert torch.equal(y1, ops.dynamic_quant_matmul(a, w.t(), s))
    normal = y1.abs() >= 2.0 ** -13   # see test_properties_full_size_int4
    assert torch.equal(ops.dynamic_quant_matmul(a, w.t(), s * 2)[normal], (y1 * 2)[normal])
    e = torch.zeros((1, k), dtype=torch.float16, device=DEV)
    e[0, 77] = 1.0
    want = (w[:, 77].float() * s.float()).half()
    assert torch.equal(ops.dynamic_quant_matmul(e, w.t(), s)[0], want)
    assert_parity(from_torch(y1), from_torch(ops.dynamic_quant_matmul(a, w.t(), s, impl=ops.IMPL_SIMPLE)), "int8 full")


# ------------------------------------------------------------------ module-level seam (S2)
def test_module_forward_matches_oracle():
    k, n, m = 4096, 4608, 3
    a, bq, s = make_int4_case(21, m, k, n, "Q")
    bias = orc.round_to(np.random.default_rng(2).standard_normal(n) * 0.02, "float16")
    lin = m4.DynamicQuantizeLinear(k, n, bias=True, device=DEV, dtype=torch.float16)
    lin.apply_weights_(u8(bq), to_torch(s, "float16"), to_torch(bias, "float16"))
    with torch.no_grad():
        y = lin(to_torch(a, "float16").reshape(1, m, k))
    assert y.shape == (1, m, n)
    assert_parity(from_torch(y)[0], c_oracle.w4a16_gemm(a, bq, s, bias, "float16"), "W4Linear")
    a8, q8, s8 = make_int8_case(22, m, k, n, "Q")
    lin8 = m8.DynamicQuantizeLinear(k, n, bias=False, device=DEV, dtype=torch.float16)
    lin8.apply_weights_(u8(q8), to_torch(s8, "float16"))
    with torch.no_grad():
        y8 = lin8(to_torch(a8, "float16"))
    assert_parity(from_torch(y8), c_oracle.w8a16_gemm(a8, q8, s8, None, "float16"), "W8Linear")
    # differentiable in the activation (DynamicQuantizeMatMul.backward, int4/qlinear.py:53-64)
    at = to_torch(a, "float16").requires_grad_()
    go = orc.round_to(np.random.default_rng(3).standard_normal((m, n)) * 0.1, "float16")
    m4.dynamic_quant_matmul(at, u8(bq), to_torch(s, "float16")).backward(to_torch(go, "float16"))
    assert_parity(from_torch(at.grad), orc.qmatmul_int4_grad_a(go, bq, s, "float16"), "int4 autograd grad_A")


# ------------------------------------------------------------------ the CUDA-core kernels are a net, not a path
def test_real_layer_shapes_never_take_the_simple_kernels():
    """AUTO silently took the one-thread-per-column kernel for shapes the TMA kernels reject.  It is
    counted now (cgq_simple_fallback_count) and can be forbidden (cgq_forbid_simple / CGQ_FORBID_SIMPLE=1): every
    ChatGLM2-6B layer shape at every M stays on the TMA / tcgen05 kernels, an N % 16 != 0 shape is counted, and with
    the net forbidden it raises instead of running slowly."""
    from chatglm_q_b200 import _lib
    lib = _lib.load()
    before = lib.cgq_simple_fallback_count()
    for (k, n) in [(4096, 4608), (4096, 4096), (13696, 4096), (4096, 27392), (4096, 65024)]:
        bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV)
        s = (torch.rand((k // 32, n), device=DEV) * 0.02 - 0.01).half()
        q8 = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=DEV)
        s8 = (torch.rand(n, device=DEV) * 0.01).half()
        for m in (1, 3, 8, 9, 128):
            a = torch.randn((m, k), device=DEV).half()
            ops.dynamic_quant_matmul_s4(a, bq, s)
            ops.dynamic_quant_matmul(a, q8.t(), s8)
    torch.cuda.synchronize()
    assert lib.cgq_simple_fallback_count() == before, "a real layer shape fell back to the CUDA-core kernel"
    k, n = 512, 200                                      # N % 16 != 0: only the shape-general kernel takes it
    a, bq, s = make_int4_case(9, 2, k, n, "Q")
    got = run4(a, bq, s, "float16")
    assert lib.cgq_simple_fallback_count() == before + 1
    assert_parity(got, c_oracle.w4a16_gemm(a, bq, s, None, "float16"), "int4 N=200 on the net")
    prev = lib.cgq_forbid_simple(1)
    try:
        with pytest.raises(RuntimeError, match="CGQ_FORBID_SIMPLE"):
            run4(a, bq, s, "float16")
        run4(a, bq, s, "float16", impl=ops.IMPL_SIMPLE)   # naming the kernel explicitly is still allowed
    finally:
        lib.cgq_forbid_simple(prev)


# ------------------------------------------------------------------ backward
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_grad_a_matches_oracle_and_reference_golden(dtype):
    """cgq_w4a16_grad_a / cgq_w8a16_grad_a against the oracle on the golden inputs of the real reference's autograd
    (tests/golden/backward.npz) and on shapes the reference's Triton backward cannot take (N = 13696: not a power of
    two, int4/triton_ops.py:226-229)."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "backward.npz")
    for tag in ("a", "b"):
        go = orc.round_to(g[f"s4_{tag}_grad_out"], dtype)
        sc = orc.round_to(g[f"s4_{tag}_scale"], dtype)
        got = ops.dynamic_quant_matmul_transposed_s4(to_torch(go, dtype), u8(g[f"s4_{tag}_b"]), to_torch(sc, dtype))
        assert_parity(from_torch(got), orc.qmatmul_int4_grad_a(go, g[f"s4_{tag}_b"], sc, dtype), f"int4 grad_A {tag}", rtol=rtol_for(dtype))
        if dtype == "float16":   # fp32 golden of the reference itself, fp16 rounding of inputs / output inside the bar
            assert_parity(from_torch(got), g[f"s4_{tag}_grad_a"], f"int4 grad_A {tag} vs reference fp32", rtol=2e-2)
        s8 = orc.round_to(g[f"s8_{tag}_scale"], dtype)
        w = torch.from_numpy(g[f"s8_{tag}_w"]).to(DEV)
        got = ops.dynamic_quant_matmul_transposed(to_torch(go, dtype), w.t(), to_torch(s8, dtype))
        assert_parity(from_torch(got), orc.qmatmul_int8_grad_a(go, g[f"s8_{tag}_w"], s8, dtype), f"int8 grad_A {tag}", rtol=rtol_for(dtype))
    # a real layer shape with leading dims: w_out of ChatGLM2-6B seen from its output side
    k, n, m = 13696, 4096, 6
    a, bq, s = make_int4_case(31, m, k, n, "Q", dtype)
    go = orc.round_to(np.random.default_rng(5).standard_normal((2, 3, n)) * 0.05, dtype)
    got = ops.dynamic_quant_matmul_transposed_s4(to_torch(go, dtype), u8(bq), to_torch(s, dtype))
    assert got.shape == (2, 3, k)
    assert_parity(from_torch(got), orc.qmatmul_int4_grad_a(go, bq, s, dtype), "int4 grad_A w_out", rtol=rtol_for(dtype))
    empty = ops.dynamic_quant_matmul_transposed_s4(to_torch(go[:0], dtype), u8(bq), to_torch(s, dtype))
    assert empty.shape == (0, 3, k)


def test_install_routes_reference_backward():
    """After install() the unmodified reference `DynamicQuantizeMatMul.backward` (int4/qlinear.py:53-64) runs on
    cgq_w4a16_grad_a, also for shapes its own Triton backward asserts on."""
    import sys
    from pathlib import Path
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    sys.path.insert(0, str(ref))
    from chatglm_q.int4 import qlinear as rq4
    from chatglm_q_b200.install import install, uninstall
    install("chatglm_q")
    try:
        k, n, m = 256, 208, 4
        a, bq, s = make_int4_case(41, m, k, n, "Q")
        at = to_torch(a, "float16").requires_grad_()
        go = orc.round_to(np.random.default_rng(6).standard_normal((m, n)) * 0.1, "float16")
        rq4.dynamic_quant_matmul(at, u8(bq), to_torch(s, "float16")).backward(to_torch(go, "float16"))
        assert_parity(from_torch(at.grad), orc.qmatmul_int4_grad_a(go, bq, s, "float16"), "reference autograd on cgq")
    finally:
        uninstall("chatglm_q")


# ------------------------------------------------------------------ graph-captured decode step
@pytest.mark.parametrize("batch", [1, 4])
def test_graph_decode_matches_eager_reference(batch):
    """The CUDA-graph wrapper must reproduce the unmodified reference model token by token (greedy),
    driven by the unmodified reference decoder loop contract: model(input_ids=, past_key_values=); batch 4 is
    BASELINE config 4's `model.forward(input_ids [B, L])` decode (a static window of B rows, M = B linears)."""
    import sys
    from pathlib import Path
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    sys.path.insert(0, str(ref))
    from chatglm_q.loader import create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q_b200.graph_decode import GraphDecodeModel
    from chatglm_q_b200.install import install, uninstall

    install("chatglm_q")
    try:
        cfg = ChatGLM2Config(hidden_size=512, inner_hidden_size=1024, head_hidden_size=64, num_multi_query_groups=2,
                             num_attention_heads=8, num_layers=3, vocab_size=1024, max_sequence_length=256)
        with torch.device(DEV):
            model = create_quant_int4_model(cfg, 32, torch.float16)
        g = torch.Generator(device=DEV).manual_seed(3)
        with torch.no_grad():
            for mod in model.modules():
                if isinstance(mod, DynamicQuantizeLinear):
                    k, n = mod.in_features, mod.out_features
                    w = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
                    sc = (torch.rand((k // 32, n), device=DEV, generator=g) * 0.5 + 0.75) / (4.4 * k ** 0.5)
                    mod.apply_weights_(w, sc.half(), torch.zeros(n, device=DEV).half() if mod.bias is not None else None)
                elif isinstance(mod, QEmbedding):
                    mod.weight.copy_(torch.randint(0, 256, mod.weight.shape, dtype=torch.uint8, device=DEV, generator=g))
                    mod.weight_scale.copy_((torch.rand(mod.weight_scale.shape, device=DEV, generator=g) * 0.2 + 0.05).half())
        model.eval()
        prompt = torch.tensor([[5, 17, 300, 42, 7, 99, 1000], [8, 1, 2, 3, 4, 5, 6], [900, 800, 700, 600, 500, 400, 300],
                               [11, 12, 13, 14, 15, 16, 17]][:batch], device=DEV)
        wrapped = GraphDecodeModel(model, max_len=64)
        with torch.no_grad():
            _, lg_e, kv_e = model(input_ids=prompt)
            _, lg_g, kv_g = wrapped(input_ids=prompt, past_key_values=None)
            assert torch.equal(lg_e, lg_g)
            tok = lg_e[:, -1].argmax(-1).reshape(batch, 1)
            for step in range(20):
                _, lg_e, kv_e = model(input_ids=tok, past_key_values=kv_e)
                _, lg_g, kv_g = wrapped(input_ids=tok, past_key_values=kv_g)
                assert wrapped.graph is not None and wrapped._eager_kv is None, "the step must be a graph replay"
                a, b = lg_e[:, -1].float(), lg_g[:, -1].float()
                assert torch.isfinite(b).all()
                err = (a - b).abs().max().item()
                assert err <= 2e-2 * a.abs().max().item() + 1e-3, f"step {step}: logits differ by {err}"
                assert torch.equal(a.argmax(-1), b.argmax(-1)), f"step {step}: greedy token differs"
                tok = a.argmax(-1).reshape(batch, 1)
    finally:
        uninstall("chatglm_q")


# ------------------------------------------------------------------ repeated launches, multi-wave grids
@pytest.mark.parametrize("n", [65024, 32768])
def test_decode_kernel_stress(n):
    """Every M <= 8 of the decode kernel, launched 200 times on grids larger than one wave of co-resident CTAs
    (508 / 512 CTAs) WHILE a second stream streams copies through L2 (so ring refills hit both L2 and HBM, and the
    SMs are shared with foreign CTAs), must agree with the bit-faithful CUDA-core kernel and with itself bit for
    bit on EVERY launch (this caught the ring-release race of DESIGN.md §3.1 and the M >= 5 divergence of the
    subnormal-operand variant, which is root-caused in DESIGN.md §3.1a)."""
    k = 4096
    g = torch.Generator(device=DEV).manual_seed(n)
    bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
    s = (torch.rand((k // 32, n), device=DEV, generator=g) * 0.02 - 0.01).half()
    side = torch.cuda.Stream()
    ha = torch.empty(96 << 20, dtype=torch.uint8, device=DEV)
    hb = torch.empty_like(ha)
    reps = 200
    for m in (8, 7, 6, 5, 4, 2, 1):
        a = torch.randn((m, k), device=DEV, generator=g).half()
        ref = ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
        ys = []
        for rep in range(reps):
            if rep % 2 == 0:
                with torch.cuda.stream(side):
                    hb.copy_(ha, non_blocking=True)      # ~30 us of L2 / HBM traffic beside the launches
            ys.append(ops.dynamic_quant_matmul_s4(a, bq, s))
        torch.cuda.synchronize()
        assert_parity(from_torch(ys[0]), from_torch(ref), f"stress M={m} N={n}")
        stack = torch.stack(ys)
        same = (stack == stack[0]).all(dim=(1, 2))
        assert bool(same.all()), f"M={m} N={n}: launches {torch.nonzero(~same).flatten().tolist()[:8]} differ from launch 0"


# ------------------------------------------------------------------ tensor-parallel shard shapes
@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_shard_shapes_on_the_decode_kernel(world):
    """Every per-rank linear of the TP plan (column slices of qkv / w_in / lm_head, k-row slices of o_proj / w_out:
    K = 512 .. 6848, 54/53-group splits at T=8) through the decode kernel at M=1 against the CUDA-core kernel."""
    from chatglm_q_b200 import tp

    g = torch.Generator(device=DEV).manual_seed(world)
    seen = set()
    for rank in {0, world - 1}:
        plan = tp.plan_block(world, rank)
        shapes = [(4096, plan.qkv.n_out(4608)), (plan.o.k_in(4096), 4096), (4096, plan.w_in.n_out(27392)),
                  (plan.w_out.k_in(13696), 4096), (4096, plan.lm_head.n_out(65024))]
        for k, n in shapes:
            if (k, n) in seen:
                continue
            seen.add((k, n))
            bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
            s = ((torch.rand((k // 32, n), device=DEV, generator=g) * 0.5 + 0.75) / (4.4 * k ** 0.5)).half()
            a = torch.randn((1, k), device=DEV, generator=g).half()
            y = ops.dynamic_quant_matmul_s4(a, bq, s)
            assert_parity(from_torch(y), from_torch(ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)),
                          f"T={world} rank {rank} K={k} N={n}")
