"""CPU model of the integer-MMA arithmetic of the one-token int4 decode kernel (csrc/gemv_w4.cu `kImma`,
`put_digits` in csrc/w4_dev.cuh), restated in numpy with the kernel's own bit tricks and checked
against the fp64 value of the same sum -- no GPU needed.  What it pins:
  * the per-group power-of-two scaling from the 16-bit patterns (fp16 and bf16), the poison rule for inf / NaN;
  * `(X + 0x00808080) ^ 0x00808080` = the four SIGNED base-256 digits of the int32 X (least significant first);
  * even k enter as the nibble q, odd k as 16 q against the activation divided by 16;
  * the accumulator that starts from 0x4B400000 - 8 (sum_even d + 16 sum_odd d) IS the fp32 number 1.5 * 2^23 + the
    exact integer group sum (so one FADD recovers it), for every digit;
  * the result equals the exact sum to fp32-accumulation accuracy."""
import numpy as np
import pytest

from oracle import qmatmul_oracle as orc

MAGIC_I = 0x4B400000
MAGIC_F = np.float32(12582912.0)


def patterns16(a: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "float16":
        return a.astype(np.float16).view(np.uint16)
    return orc.bf16_bits(a)


def put_digits(group: np.ndarray, dtype: str):
    """32 activations of one quantisation group -> (digits [4, 32] int8 with odd k pre-divided by 16, inv scale)."""
    pat = patterns16(group, dtype) & 0x7FFF
    mx = int(pat.max())
    if dtype == "float16":
        e = mx >> 10
        bad, mb = e == 31, max(e, 1) + 112
    else:
        e = mx >> 7
        bad, mb = e == 255, max(e, 30)
    if bad:
        mb = 127
    sc = np.float32(2.0) ** np.float32(283 - mb - 127)
    xs = group.astype(np.float32) * np.where(np.arange(32) % 2 == 1, sc * np.float32(0.0625), sc).astype(np.float32)
    with np.errstate(invalid="ignore"):
        x = np.clip(np.rint(xs.astype(np.float64)), -2 ** 31, 2 ** 31 - 1)
    x = np.where(np.isfinite(xs), x, 0).astype(np.int64).astype(np.int32)          # cvt.rni.sat.s32.f32 (NaN -> 0)
    z = ((x.view(np.uint32) + np.uint32(0x00808080)) ^ np.uint32(0x00808080)).astype(np.uint32)
    digits = np.stack([((z >> (8 * d)) & 0xFF).astype(np.uint8).view(np.int8) for d in range(4)])
    inv = np.float32(np.nan) if bad else np.float32(2.0) ** np.float32(mb - 29 - 127)
    return digits, inv, x


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_digits_reassemble_the_scaled_integer(dtype):
    rng = np.random.default_rng(1)
    for _ in range(200):
        g = orc.round_to(rng.standard_normal(32) * np.exp2(rng.integers(-20, 12, 32)), dtype)
        digits, inv, x = put_digits(g, dtype)
        re = sum(digits[d].astype(np.int64) << (8 * d) for d in range(4))
        assert np.array_equal(re, x.astype(np.int64))
        # the scaled group maximum lands in [2^29, 2^30) (an odd k carries it divided by 16): no int32 overflow
        top = (np.abs(x.astype(np.int64)) * np.where(np.arange(32) % 2 == 1, 16, 1)).max()
        assert (2 ** 29 - 16 <= top < 2 ** 30 + 16) or not g.any()
        # elements within 2^-19 (even k) / 2^-15 (odd k: divided by 16) of the group maximum are exact
        scale = float(inv)
        back = x.astype(np.float64) * scale * np.where(np.arange(32) % 2 == 1, 16.0, 1.0)
        big = np.abs(g) >= np.abs(g).max() * 2.0 ** -15
        assert np.array_equal(back[big], g.astype(np.float64)[big])
        assert np.abs(back - g).max() <= np.abs(g).max() * 2.0 ** -25


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_group_sum_through_the_magic_accumulator(dtype):
    """One k-group of one weight column, exactly as the MMAs see it: unsigned bytes q / 16 q against signed digits,
    accumulator start 0x4B400000 - 8 (...), reinterpreted as fp32."""
    rng = np.random.default_rng(2)
    for _ in range(300):
        g = orc.round_to(rng.standard_normal(32) * np.exp2(rng.integers(-10, 8, 32)), dtype)
        q = rng.integers(0, 16, 32)                                    # nibbles of one column, k = 0 .. 31
        digits, inv, _ = put_digits(g, dtype)
        a_bytes = np.where(np.arange(32) % 2 == 1, 16 * q, q).astype(np.int64)          # & 0x0F0F0F0F / & 0xF0F0F0F0
        neg = np.where(np.arange(32) % 2 == 1, -128, -8).astype(np.int64)               # the constant rows of the extra MMA
        total = 0.0
        for d in range(4):
            start = MAGIC_I + int((neg * digits[d].astype(np.int64)).sum())
            acc = start + int((a_bytes * digits[d].astype(np.int64)).sum())              # s32 accumulator of the IMMA
            assert abs(acc - MAGIC_I) < 2 ** 22, "the sum must stay inside the binade of 1.5 * 2^23"
            e = np.array([acc], dtype=np.int32).view(np.float32)[0] - MAGIC_F           # one FADD: the exact integer
            want = int((((q - 8) * np.where(np.arange(32) % 2 == 1, 16, 1)) * digits[d].astype(np.int64)).sum())
            assert float(e) == want
            total += float(e) * 256.0 ** d
        exact = float(((q - 8).astype(np.float64) * g.astype(np.float64)).sum())
        got = total * float(inv)
        assert abs(got - exact) <= 2.0 ** -22 * float(np.abs((q - 8) * g.astype(np.float64)).sum()) + 1e-300


def test_nonfinite_groups_are_poisoned():
    g = np.zeros(32, dtype=np.float32)
    g[3] = np.inf
    assert np.isnan(put_digits(g, "float16")[1])
    g[3] = np.nan
    assert np.isnan(put_digits(g, "bfloat16")[1])
    assert put_digits(np.zeros(32, np.float32), "float16")[1] > 0          # an all-zero group is fine (digits 0)
