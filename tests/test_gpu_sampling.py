"""GPU suite (-m gpu) of the one-launch token sampler (cgq_top_p_sample): against
oracle/sampling_oracle.py, against the fixture captured from the REAL reference function
(tests/golden/sampling.npz) and, when the pip-installed reference is present (baseline/_ref), against the
unmodified `chatglm_q.decoder.top_p_sampling` running on the same GPU with the same torch seed.
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import qmatmul_oracle as orc
from oracle import sampling_oracle as so

pytestmark = pytest.mark.gpu

from chatglm_q_b200 import ops  # noqa: E402
from chatglm_q_b200._lib import CgqError  # noqa: E402

DEV = "cuda"
TD = {"float16": torch.float16, "bfloat16": torch.bfloat16}


def _fixture_cases():
    fx = dict(np.load(Path(__file__).resolve().parent / "golden" / "sampling.npz"))
    out = []
    for entry in fx["cases"]:
        name, dtype = str(entry).split(":")
        out.append((name, dtype, fx))
    return out


def _logits_from_bits(bits, dtype):
    t = torch.from_numpy(bits.view(np.int16).copy()).to(DEV).view(TD[dtype])
    f32 = orc.bf16_from_bits(bits) if dtype == "bfloat16" else bits.view(np.float16).astype(np.float32)
    return t, f32


def _check_distribution(logits_f32, p, idx, top_k, top_p, temp, what):
    """CUDA (probs, indices) against the oracle: identical ids (both break ties by the lower id), probabilities
    to fp32 rounding of a 65 024-term sum."""
    ref_p, ref_idx = so.top_p_distribution(logits_f32, top_k, top_p, temp)
    assert p.shape == ref_p.shape and idx.shape == ref_idx.shape, what
    assert np.array_equal(idx, ref_idx), f"{what}: ids differ at ranks {np.nonzero(idx != ref_idx)[0][:8]}"
    np.testing.assert_allclose(p, ref_p, rtol=1e-5, atol=1e-9, err_msg=what)
    assert np.array_equal(p == 0, ref_p == 0), f"{what}: top-p mask differs"
    assert abs(float(p.sum()) - 1.0) < 1e-5, what


@pytest.mark.parametrize("case", _fixture_cases(), ids=lambda c: c[0])
def test_sampler_matches_reference_fixture(case):
    name, dtype, fx = case
    top_k, top_p, temp = fx[f"{name}_params"]
    top_k, top_p, temp = int(top_k), float(top_p), float(temp)
    logits, f32 = _logits_from_bits(fx[f"{name}_logits_bits"], dtype)
    p, idx = ops.top_p_distribution(logits, top_k, top_p, temp)
    p, idx = p.cpu().numpy(), idx.cpu().numpy()
    _check_distribution(f32, p, idx, top_k, top_p, temp, name)
    # the reference's own numbers: same probability at every rank, same ids where the sort had no tie to break
    ref_p, ref_idx = fx[f"{name}_probs"], fx[f"{name}_indices"]
    np.testing.assert_allclose(p, ref_p, rtol=1e-5, atol=1e-9, err_msg=name)
    assert np.array_equal(f32[idx], f32[ref_idx]), name
    # the reference's Exp(1) variates -> the token the reference returned (or, on a tie, one of equal probability)
    q = torch.from_numpy(fx[f"{name}_q"]).to(DEV)
    tok = int(ops.top_p_sampling(logits, top_k, top_p, temp, q=q).item())
    assert tok == so.sample_with(p, idx, fx[f"{name}_q"]), name
    assert f32[tok] == f32[int(fx[f"{name}_token"])], name


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("v,top_k,top_p,temp", [
    (65024, 100, 0.8, 1.0), (65024, 1, 0.8, 1.0), (65024, 1024, 0.999, 1.0), (65023, 100, 0.3, 0.5),
    (2049, 100, 0.8, 1.7), (9, 100, 0.8, 1.0), (1, 5, 0.8, 1.0), (70000, 40, 0.0, 1.0)])
def test_sampler_vs_oracle_seeded(dtype, v, top_k, top_p, temp):
    g = torch.Generator().manual_seed(v * 7 + top_k)
    logits = (torch.randn(v + 1, generator=g) * 2.5).to(TD[dtype]).to(DEV)
    for off in (0, 1):                      # off = 1: a 2-byte aligned row (scalar staging path)
        row = logits[off:off + v]
        p, idx = ops.top_p_distribution(row, top_k, top_p, temp)
        _check_distribution(row.float().cpu().numpy(), p.cpu().numpy(), idx.cpu().numpy(), top_k, top_p, temp,
                            f"{dtype} V={v} k={top_k} off={off}")


def test_sampler_heavy_ties_take_lowest_ids():
    v = 65024
    logits = torch.zeros(v, dtype=torch.float16, device=DEV)
    p, idx = ops.top_p_distribution(logits, 100, 0.8, 1.0)
    assert np.array_equal(idx.cpu().numpy(), np.arange(100))
    np.testing.assert_allclose(p.cpu().numpy(), 0.01, rtol=1e-5)    # 100 / 65 024 of the mass: nothing is masked
    logits[40000] = 1.0
    logits[5:20] = -1.0
    p, idx = ops.top_p_distribution(logits, 100, 1.0, 1.0)
    want = np.concatenate([[40000], np.arange(5), np.arange(20, 114)])
    assert np.array_equal(idx.cpu().numpy(), want)
    # -inf rows are legal logits (masked vocabulary): never selected while finite ones remain
    logits = torch.full((v,), float("-inf"), dtype=torch.float16, device=DEV)
    logits[[7, 99, 64000]] = torch.tensor([0.5, 2.0, 1.0], dtype=torch.float16, device=DEV)
    p, idx = ops.top_p_distribution(logits, 3, 1.0, 1.0)
    assert idx.cpu().tolist() == [99, 64000, 7]
    assert abs(float(p.sum()) - 1) < 1e-6


def test_sampler_batched_rows_and_determinism():
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(2, 3, 5000, generator=g) * 3).half().to(DEV)
    p, idx = ops.top_p_distribution(logits, 50, 0.9, 1.0)
    assert p.shape == (2, 3, 50) and idx.shape == (2, 3, 50)
    for b in range(2):
        for r in range(3):
            _check_distribution(logits[b, r].float().cpu().numpy(), p[b, r].cpu().numpy(), idx[b, r].cpu().numpy(),
                                50, 0.9, 1.0, f"row {b},{r}")
    q = torch.empty(2, 3, 50, device=DEV).exponential_(1)
    t1 = ops.top_p_sampling(logits, 50, 0.9, 1.0, q=q)
    t2 = ops.top_p_sampling(logits, 50, 0.9, 1.0, q=q)
    assert t1.shape == (2, 3) and t1.dtype == torch.int64 and torch.equal(t1, t2)
    for b in range(2):
        for r in range(3):
            assert int(t1[b, r]) == so.sample_with(p[b, r].cpu().numpy(), idx[b, r].cpu().numpy(), q[b, r].cpu().numpy())


def test_sampler_frequencies_follow_the_distribution():
    g = torch.Generator().manual_seed(11)
    logits = (torch.randn(65024, generator=g) * 4).half().to(DEV)
    p, idx = ops.top_p_distribution(logits, 100, 0.8, 1.0)
    p, idx = p.cpu().numpy(), idx.cpu().numpy()
    torch.manual_seed(5)
    n = 4000
    toks = torch.stack([ops.top_p_sampling(logits, 100, 0.8, 1.0) for _ in range(n)]).cpu().numpy()
    kept = idx[p > 0]
    assert np.isin(toks, kept).all()
    top = int(idx[0])
    freq = float((toks == top).mean())
    sigma = np.sqrt(p[0] * (1 - p[0]) / n)
    assert abs(freq - p[0]) < 5 * sigma + 1e-3, (freq, p[0])


def _reference_sampler():
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import chatglm_q.decoder as dec

    return dec


def test_sampler_same_seed_same_token_as_unmodified_reference():
    """The unmodified reference function on the same GPU: with the same torch seed both draw the same Exp(1)
    variates (torch.multinomial's own `empty_like(probs).exponential_(1)`), so the tokens agree -- up to the
    reference sort's arbitrary order among tied probabilities and fp32 rounding of near-equal p / q ratios."""
    dec = _reference_sampler()
    g = torch.Generator().manual_seed(23)
    same, n = 0, 0
    for trial in range(40):
        scale = (1.0, 2.5, 5.0, 8.0)[trial % 4]
        logits = (torch.randn(65024, generator=g) * scale).half().to(DEV)
        top_k, top_p, temp = ((100, 0.8, 1.0), (50, 0.95, 0.8), (100, 0.6, 1.5))[trial % 3]
        torch.manual_seed(1000 + trial)
        t_ref = int(dec.top_p_sampling(logits, top_k, top_p, temp).item())
        torch.manual_seed(1000 + trial)
        t_new = int(ops.top_p_sampling(logits, top_k, top_p, temp).item())
        n += 1
        same += t_ref == t_new
        if t_ref != t_new:       # only a tie in the reference's sort may move the token
            assert float(logits[t_ref]) == float(logits[t_new]), (trial, t_ref, t_new)
    assert same >= int(0.8 * n), f"{same}/{n} tokens equal"


def test_install_rebinds_the_decoder_sampler():
    dec = _reference_sampler()
    from chatglm_q_b200.install import install, uninstall

    original = dec.top_p_sampling
    install("chatglm_q", sampler=True)
    try:
        assert dec.top_p_sampling is ops.top_p_sampling
    finally:
        uninstall("chatglm_q")
    assert dec.top_p_sampling is original


def test_sampler_rejects_what_it_cannot_take():
    logits = torch.randn(1000, device=DEV)
    with pytest.raises(TypeError):
        ops.top_p_sampling(logits)                                  # fp32 logits: no 32-bit select is built
    with pytest.raises(AssertionError):
        ops.top_p_sampling(torch.randn(1000).half())                # CPU tensor: no fallback
    with pytest.raises(CgqError):
        ops.top_p_distribution(torch.randn(5000, device=DEV).half(), top_k=2000)
    with pytest.raises(CgqError):
        ops.top_p_distribution(torch.randn(200000, device=DEV).half())   # row larger than shared memory
