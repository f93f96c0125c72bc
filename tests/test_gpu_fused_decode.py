"""GPU suite (-m gpu) of the fused batch-1 decode step: the fused int4 linear
(RMSNorm / SiLU-gate prologue, residual epilogue), the RoPE + KV + attention kernel and the whole step
behind the model call signature — against oracle/decode_oracle.py on seeded inputs, against the golden
fixture made by the REAL reference model (tests/golden/decode_tiny.npz) and, when the pip-installed
reference is present (baseline/_ref), against the unmodified reference model running on the same GPU.
"""
import os
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import decode_oracle as dec
from oracle import qmatmul_oracle as orc
from util import assert_parity, from_torch, load_decode_golden, make_int4_case, rtol_for, to_torch

pytestmark = pytest.mark.gpu

from chatglm_q_b200 import ops  # noqa: E402
from chatglm_q_b200._lib import ARITH_SUBNORMAL, PRO_NONE, PRO_RMSNORM, PRO_SILU_GATE  # noqa: E402
from chatglm_q_b200.fused_decode import FusedDecodeModel, _FusedCache, accelerate  # noqa: E402

DEV = "cuda"


def u8(x):
    return torch.from_numpy(np.ascontiguousarray(x)).to(DEV)


# ------------------------------------------------------------------ fused linear
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("k,n", [(4096, 4608), (4096, 27392), (512, 768), (4096 + 2048, 256), (1056, 144)])
def test_gemv_fused_rmsnorm_bias(dtype, k, n):
    a, bq, s = make_int4_case(31, 1, k, n, "Q", dtype)
    rng = np.random.default_rng(5)
    x = orc.round_to(a[0] * 3.0, dtype)
    nw = orc.round_to(1.0 + 0.2 * rng.standard_normal(k), dtype)
    bias = orc.round_to(rng.standard_normal(n) * 0.05, dtype)
    got = ops.gemv_fused_s4(to_torch(x, dtype), u8(bq), to_torch(s, dtype), bias=to_torch(bias, dtype),
                            prologue=PRO_RMSNORM, norm_weight=to_torch(nw, dtype), eps=1e-5)
    want = orc.qmatmul_int4(dec.rmsnorm(x, nw, 1e-5, dtype)[None], bq, s, bias, dtype)[0]
    assert_parity(from_torch(got), want, f"rmsnorm+linear {dtype} K={k} N={n}", rtol=rtol_for(dtype))
    # plain prologue == the module-level op on the same row
    plain = ops.gemv_fused_s4(to_torch(x, dtype), u8(bq), to_torch(s, dtype), bias=to_torch(bias, dtype))
    mod = ops.dynamic_quant_matmul_s4(to_torch(x, dtype)[None], u8(bq), to_torch(s, dtype), bias=to_torch(bias, dtype))
    assert torch.equal(plain, mod[0])


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("k,n", [(13696, 4096), (1024, 512), (384, 256)])
def test_gemv_fused_silu_gate_residual(dtype, k, n):
    _, bq, s = make_int4_case(32, 1, k, n, "Q", dtype)
    rng = np.random.default_rng(6)
    u = orc.round_to(rng.standard_normal(2 * k) * 1.5, dtype)
    resid = orc.round_to(rng.standard_normal(n), dtype)
    x = to_torch(resid, dtype)
    got = ops.gemv_fused_s4(to_torch(u, dtype), u8(bq), to_torch(s, dtype), resid=x, prologue=PRO_SILU_GATE, out=x)
    assert got.data_ptr() == x.data_ptr()      # in-place residual update, as the step uses it
    d = orc.qmatmul_int4(dec.silu_gate(u, dtype)[None], bq, s, None, dtype)[0]
    want = orc.round_to(resid + d, dtype)
    assert_parity(from_torch(got), want, f"silu-gate+linear+resid {dtype} K={k} N={n}", rtol=rtol_for(dtype))


def test_gemv_fused_prologue_pieces_bit_exact():
    """With a unit weight (one column per k, nibble 9 = +1, scale 1) the linear returns its own input row, so
    the fused RMSNorm / SiLU-gate values themselves can be compared with the oracle's.  bfloat16 takes the
    exact (q - 8) dequant, so the row passes through unchanged; what may differ is the device's rsqrt / exp
    against numpy's in the last fp32 bit (two roundings => at most 2 bf16 ulps, and rarely)."""
    k, dt = 256, "bfloat16"
    rng = np.random.default_rng(9)
    x = orc.round_to(rng.standard_normal(k) * 2.0, dt)
    nw = orc.round_to(1.0 + 0.3 * rng.standard_normal(k), dt)
    q = np.full((k, k), 8, dtype=np.uint8)
    q[np.arange(k), np.arange(k)] = 9
    bq = (q[0::2] | (q[1::2] << 4)).astype(np.uint8)
    s = np.ones((k // 32, k), dtype=np.float32)
    ident = ops.gemv_fused_s4(to_torch(x, dt), u8(bq), to_torch(s, dt))
    assert np.array_equal(from_torch(ident), x)
    got = ops.gemv_fused_s4(to_torch(x, dt), u8(bq), to_torch(s, dt), prologue=PRO_RMSNORM,
                            norm_weight=to_torch(nw, dt), eps=1e-5)
    want = dec.rmsnorm(x, nw, 1e-5, dt)
    diff = np.abs(from_torch(got) - want)
    assert (diff <= np.abs(want) * 2.0 ** -6 + 1e-7).all() and (diff == 0).mean() > 0.97, (diff.max(), (diff == 0).mean())
    u = orc.round_to(rng.standard_normal(2 * k) * 2.0, dt)
    got = ops.gemv_fused_s4(to_torch(u, dt), u8(bq), to_torch(s, dt), prologue=PRO_SILU_GATE)
    want = dec.silu_gate(u, dt)
    diff = np.abs(from_torch(got) - want)
    assert (diff <= np.abs(want) * 2.0 ** -6 + 1e-7).all() and (diff == 0).mean() > 0.97, (diff.max(), (diff == 0).mean())


# ------------------------------------------------------------------ attention
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("d_head,n_head,n_groups,n_past,window", [
    (128, 32, 2, 0, 0), (128, 32, 2, 1, 0), (128, 32, 2, 159, 0), (128, 32, 2, 700, 0), (64, 8, 2, 37, 0),
    (64, 4, 4, 5, 0),
    # a large KV window deals the context to a cluster of 8 CTAs per head: nearly empty, uneven, full
    (128, 32, 2, 5, 1024), (128, 32, 2, 130, 1024), (128, 32, 2, 1021, 1024), (64, 8, 2, 300, 512)])
def test_decode_attention_vs_oracle(dtype, d_head, n_head, n_groups, n_past, window):
    rng = np.random.default_rng(100 + n_past)
    max_len = window or max(n_past + 3, 16)
    qkv = orc.round_to(rng.standard_normal(d_head * (n_head + 2 * n_groups)), dtype)
    kc = orc.round_to(rng.standard_normal((max_len, n_groups, d_head)), dtype)
    vc = orc.round_to(rng.standard_normal((max_len, n_groups, d_head)), dtype)
    pos = np.arange(max_len + 2, dtype=np.float64)[:, None] * (1.0 / 10000 ** (np.arange(0, d_head // 2, 2) / (d_head // 2)))
    table = np.concatenate([np.stack([np.cos(pos), np.sin(pos)], -1), np.stack([np.ones_like(pos), np.zeros_like(pos)], -1)],
                           axis=-2).reshape(max_len + 2, d_head)
    table = orc.round_to(table.astype(np.float32), dtype)
    state = torch.tensor([n_past + 1, n_past], dtype=torch.int32, device=DEV)
    kct, vct = to_torch(kc, dtype), to_torch(vc, dtype)
    got = ops.decode_attention(to_torch(qkv, dtype), to_torch(table, dtype), kct, vct, state, n_head, n_groups, d_head)
    want, k_new, v_new = dec.attention_decode(qkv, table[n_past + 1], kc[:n_past], vc[:n_past], n_head, n_groups,
                                              d_head, dtype)
    assert_parity(from_torch(got), want, f"attention {dtype} d={d_head} L={n_past}", rtol=rtol_for(dtype))
    # the new cache rows are the roped key / raw value (one fp32 product-sum rounded once: <= 1 ulp apart)
    assert_parity(from_torch(kct[n_past]), k_new, "k cache row", rtol=rtol_for(dtype))
    assert np.array_equal(from_torch(vct[n_past]), v_new)
    assert np.array_equal(from_torch(kct[:n_past]), kc[:n_past]) and np.array_equal(from_torch(vct[n_past + 1:]), vc[n_past + 1:])


# ------------------------------------------------------------------ the whole step
def _lin(w, s, b=None):
    m = SimpleNamespace(weight=u8(w), weight_scale=to_torch(s, "float16"), bias=None if b is None else to_torch(b, "float16"))
    return m


def _duck_model(w, cfg):
    """A stand-in with the attribute layout of chatglm_q.model.ChatGLM2Model holding the golden weights."""
    layers = []
    for i in range(cfg["n_layers"]):
        p = f"l{i}_"
        layers.append(SimpleNamespace(
            attn_ln=SimpleNamespace(weight=to_torch(w[p + "attn_ln"], "float16"), eps=cfg["eps"]),
            ffn_ln=SimpleNamespace(weight=to_torch(w[p + "ffn_ln"], "float16"), eps=cfg["eps"]),
            attn=SimpleNamespace(qkv_proj=_lin(w[p + "qkv_w"], w[p + "qkv_s"], w[p + "qkv_b"]),
                                 o_proj=_lin(w[p + "o_w"], w[p + "o_s"])),
            ffn=SimpleNamespace(w_in=_lin(w[p + "win_w"], w[p + "win_s"]), w_out=_lin(w[p + "wout_w"], w[p + "wout_s"]))))
    config = SimpleNamespace(hidden_size=cfg["hidden"], inner_hidden_size=cfg["inner"], head_hidden_size=cfg["d_head"],
                             num_multi_query_groups=cfg["n_groups"], num_attention_heads=cfg["n_head"],
                             num_layers=cfg["n_layers"], vocab_size=cfg["vocab"], max_sequence_length=cfg["max_seq"])
    return SimpleNamespace(config=config, layers=layers,
                           final_ln=SimpleNamespace(weight=to_torch(w["final_ln"], "float16"), eps=cfg["eps"]),
                           lm_head=_lin(w["lm_w"], w["lm_s"]),
                           word_embedding=SimpleNamespace(weight=u8(w["emb_w"]), weight_scale=to_torch(w["emb_s"], "float16")),
                           freqs_cis_cache=to_torch(w["freqs"], "float16"))


@pytest.mark.parametrize("one_launch", [True, False])
def test_fused_step_matches_reference_golden(one_launch):
    """Weights, prefill KV and greedy tokens of the REAL reference model (CPU fp16 fixture): the fused CUDA step
    -- as ONE persistent launch (cgq_step_*) and as the PDL chain of launches -- must reproduce its logits within
    the parity bar and pick the same tokens; also against the numpy oracle."""
    w, cfg, fx = load_decode_golden()
    model = _duck_model(w, cfg)
    fused = FusedDecodeModel(model, max_len=32, one_launch=one_launch)
    n0 = len(fx["prompt"])
    kv = tuple((to_torch(fx[f"prefill_k{i}"], "float16").reshape(1, n0, cfg["n_groups"], 1, cfg["d_head"]),
                to_torch(fx[f"prefill_v{i}"], "float16").reshape(1, n0, cfg["n_groups"], 1, cfg["d_head"]))
               for i in range(cfg["n_layers"]))
    fused._import_kv(kv, torch.device("cuda", torch.cuda.current_device()))
    handle = _FusedCache(fused)
    okv = [(fx[f"prefill_k{i}"].astype(np.float32), fx[f"prefill_v{i}"].astype(np.float32)) for i in range(cfg["n_layers"])]
    for step, tok in enumerate(fx["step_tokens"]):
        _, logits, handle = fused(input_ids=torch.tensor([[int(tok)]], device=DEV), past_key_values=handle)
        torch.cuda.synchronize()
        got = from_torch(logits[0, -1])
        ref = fx["step_logits"][step].astype(np.float32)
        assert_parity(got, ref, f"fused step {step} vs reference model")
        top2 = np.sort(ref)[-2:]
        if top2[1] - top2[0] > 2e-2 * np.abs(ref).max():      # greedy token, unless the reference itself is a near tie
            assert int(got.argmax()) == int(ref.argmax())
        want, okv = dec.decode_step(w, int(tok), okv, cfg, "float16")
        assert_parity(got, want, f"fused step {step} vs oracle")
    assert fused.one_launch == one_launch, fused.one_launch_refused      # the requested path is the one that ran
    assert fused.launches_per_step() == (1 if one_launch else 5 * cfg["n_layers"] + 2)
    n1 = n0 + len(fx["step_tokens"])
    for i, (k, v) in enumerate(fused._export_kv()):
        assert k.shape == (1, n1, cfg["n_groups"], 1, cfg["d_head"])
        assert_parity(from_torch(k[0, :, :, 0]), fx[f"final_k{i}"].astype(np.float32), f"k cache {i}")
        assert_parity(from_torch(v[0, :, :, 0]), fx[f"final_v{i}"].astype(np.float32), f"v cache {i}")


def _random_ref_model(cfg_kwargs, seed=3):
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q.loader import create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config

    cfg = ChatGLM2Config(**cfg_kwargs)
    with torch.device(DEV):
        model = create_quant_int4_model(cfg, 32, torch.float16)
    g = torch.Generator(device=DEV).manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                k, n = mod.in_features, mod.out_features
                wq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
                sc = (torch.rand((k // 32, n), device=DEV, generator=g) * 0.5 + 0.75) / (4.4 * k ** 0.5)
                b = (torch.randn(n, device=DEV, generator=g) * 0.05).half() if mod.bias is not None else None
                mod.apply_weights_(wq, sc.half(), b)
            elif isinstance(mod, QEmbedding):
                mod.weight.copy_(torch.randint(0, 256, mod.weight.shape, dtype=torch.uint8, device=DEV, generator=g))
                mod.weight_scale.copy_((torch.rand(mod.weight_scale.shape, device=DEV, generator=g) * 0.2 + 0.05).half())
        for name, p in model.named_parameters():
            if name.endswith("ln.weight"):
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, device=DEV, generator=g)).half())
    return model.eval()


@pytest.mark.parametrize("cfg_kwargs", [
    dict(hidden_size=512, inner_hidden_size=1024, head_hidden_size=64, num_multi_query_groups=2, num_attention_heads=8,
         num_layers=3, vocab_size=1024, max_sequence_length=256),
    dict(hidden_size=4096, inner_hidden_size=13696, head_hidden_size=128, num_multi_query_groups=2,
         num_attention_heads=32, num_layers=2, vocab_size=65024, max_sequence_length=512),   # ChatGLM2-6B layer shapes
])
def test_fused_decode_matches_unmodified_reference_model(cfg_kwargs):
    """Driven exactly as ChatGLMDecoder.generate does (decoder.py:81-84): prefill, then one token per call.  The
    fused step must track the UNMODIFIED reference model (same kernels behind its QLinear modules) greedily."""
    from chatglm_q_b200.install import install, uninstall

    model = _random_ref_model(cfg_kwargs)
    install("chatglm_q")
    try:
        prompt = torch.tensor([[5, 17, 300, 42, 7, 99, 1000]], device=DEV)
        fused = accelerate(model, max_len=40)
        assert isinstance(fused, FusedDecodeModel)
        with torch.no_grad():
            _, lg_e, kv_e = model(input_ids=prompt)
            _, lg_f, kv_f = fused(input_ids=prompt, past_key_values=None)
            assert torch.equal(lg_e, lg_f)
            tok = lg_e[0, -1].argmax().reshape(1, 1)
            for step in range(40):     # runs past max_len: the last steps take the exported-cache fallback
                _, lg_e, kv_e = model(input_ids=tok, past_key_values=kv_e)
                _, lg_f, kv_f = fused(input_ids=tok, past_key_values=kv_f)
                a, b = lg_e[0, -1].float(), lg_f[0, -1].float()
                assert torch.isfinite(b).all()
                assert_parity(b.cpu().numpy(), a.cpu().numpy(), f"step {step} logits", rtol=2e-2)
                top2 = a.topk(2).values
                if (top2[0] - top2[1]).item() > 2e-2 * a.abs().max().item():
                    assert a.argmax().item() == b.argmax().item(), f"step {step}: greedy token differs"
                tok = a.argmax().reshape(1, 1)
        assert not fused.one_launch and fused.launches_per_step() == 5 * len(model.layers) + 2   # the default path
    finally:
        uninstall("chatglm_q")


@pytest.mark.parametrize("cfg_kwargs", [
    dict(hidden_size=512, inner_hidden_size=1024, head_hidden_size=64, num_multi_query_groups=2, num_attention_heads=8,
         num_layers=3, vocab_size=1024, max_sequence_length=256),
    dict(hidden_size=4096, inner_hidden_size=13696, head_hidden_size=128, num_multi_query_groups=2,
         num_attention_heads=32, num_layers=3, vocab_size=65024, max_sequence_length=512),   # ChatGLM2-6B layer shapes
])
@pytest.mark.skipif(os.environ.get("CGQ_TEST_HANDOVER", "0") != "1",
                    reason="experimental protocol, not yet measured on a B200: set CGQ_TEST_HANDOVER=1")
def test_fused_step_tile_handover_bit_identical(cfg_kwargs):
    """EXPERIMENTAL hand-over protocol (cgq_handover_next): o_proj -> w_in -> w_out -> next qkv / lm_head wait on
    per-tile counters instead of griddepcontrol.wait.  Same kernels, same arithmetic: logits and KV caches of 30
    replayed steps must equal the default protocol's bit for bit."""
    from chatglm_q_b200.install import install, uninstall

    model = _random_ref_model(cfg_kwargs)
    install("chatglm_q")
    try:
        prompt = torch.tensor([[5, 17, 300, 42, 7, 99, 1000]], device=DEV)
        plain = FusedDecodeModel(model, max_len=64, handover=False)
        hand = FusedDecodeModel(model, max_len=64, handover=True)
        with torch.no_grad(), ops.decode_arith(ARITH_SUBNORMAL):   # the kHand kernels keep the subnormal-operand arithmetic
            _, lg_p, kv_p = plain(input_ids=prompt, past_key_values=None)
            _, lg_h, kv_h = hand(input_ids=prompt, past_key_values=None)
            tok = lg_p[0, -1].argmax().reshape(1, 1)
            for step in range(30):
                _, lg_p, kv_p = plain(input_ids=tok, past_key_values=kv_p)
                _, lg_h, kv_h = hand(input_ids=tok, past_key_values=kv_h)
                assert torch.equal(lg_p, lg_h), f"step {step}: logits differ under the hand-over protocol"
                tok = lg_p[0, -1].argmax().reshape(1, 1)
            torch.cuda.synchronize()
            assert hand.graph is not None and plain.graph is not None
            for (kp, vp), (kh, vh) in zip(plain.kv, hand.kv):
                assert torch.equal(kp, kh) and torch.equal(vp, vh)
            # every producing launch announced exactly its tiles in the last replay
            want = []
            for layer in model.layers:
                want += [(layer.attn.o_proj.weight.shape[1] + 127) // 128, (layer.ffn.w_in.weight.shape[1] + 127) // 128,
                         (layer.ffn.w_out.weight.shape[1] + 127) // 128]
            assert hand.ctr[:, 0].cpu().tolist() == want
    finally:
        uninstall("chatglm_q")


def test_speculative_next_step_is_exact():
    """FusedDecodeModel(speculate=True) + its sampler(): the next step is started from the device-resident token
    before the host sees it.  Driven like ChatGLMDecoder.generate with CPU ids (decoder device=None): same seed ->
    the same tokens and bit-identical logits as the plain fused model with ops.top_p_sampling, across the end of the
    static KV window; a caller that feeds ANOTHER token gets the step taken back and redone."""
    from chatglm_q_b200.install import install, uninstall

    cfg_kwargs = dict(hidden_size=512, inner_hidden_size=1024, head_hidden_size=64, num_multi_query_groups=2,
                      num_attention_heads=8, num_layers=3, vocab_size=1024, max_sequence_length=256)
    model = _random_ref_model(cfg_kwargs)
    install("chatglm_q")
    try:
        def run(fm, sampler, n, on_host, swap_at=None):
            torch.manual_seed(11)
            ids = torch.tensor([[5, 17, 300, 42, 7]])
            with torch.no_grad():
                _, lg, kv = fm(input_ids=ids if on_host else ids.to(DEV), past_key_values=None)
                toks, logs = [], []
                for i in range(n):
                    tok = int(sampler(lg[0, -1], 50, 0.9, 1.0).item())
                    if swap_at is not None and i == swap_at:
                        tok = (tok + 1) % 1024            # not what was sampled: the speculation must be undone
                    toks.append(tok)
                    nxt = torch.tensor([[tok]])
                    _, lg, kv = fm(input_ids=nxt if on_host else nxt.to(DEV), past_key_values=kv)
                    logs.append(lg[0, -1].clone())
            torch.cuda.synchronize()
            return toks, logs

        for swap in (None, 7):
            plain = FusedDecodeModel(model, max_len=48)
            spec = FusedDecodeModel(model, max_len=48, speculate=True)
            t_a, l_a = run(plain, ops.top_p_sampling, 60, False, swap)      # runs past the 48-row window
            t_b, l_b = run(spec, spec.sampler(), 60, True, swap)
            assert t_a == t_b, f"tokens differ (swap={swap})"
            for i, (a, b) in enumerate(zip(l_a, l_b)):
                assert torch.equal(a, b), f"step {i}: logits differ under speculation (swap={swap})"
        # one sample per step: the logits buffer already belongs to the next step
        spec = FusedDecodeModel(model, max_len=48, speculate=True)
        samp = spec.sampler()
        with torch.no_grad():
            _, lg, kv = spec(input_ids=torch.tensor([[5, 17, 300]]), past_key_values=None)
            tok = int(samp(lg[0, -1]).item())
            _, lg, kv = spec(input_ids=torch.tensor([[tok]]), past_key_values=kv)
            t1 = samp(lg[0, -1])
            assert t1.device.type == "cpu" and spec._spec
            with pytest.raises(RuntimeError):
                samp(lg[0, -1])
    finally:
        uninstall("chatglm_q")


def test_fused_decode_rejects_unsupported_models():
    w, cfg, _ = load_decode_golden()
    model = _duck_model(w, cfg)
    model.lm_head.weight = model.lm_head.weight.to(torch.int8)       # an int8 model is not this path
    with pytest.raises(TypeError):
        FusedDecodeModel(model)


# ------------------------------------------------------------------ persistent decode program (opt-in)
def test_decode_program_bit_identical_to_per_linear_launches():
    """A two-block chain (RMSNorm / plain / SiLU-gate prologues, bias, in-place residuals, every band split
    Z = 8 / 2 / 1) as ONE persistent launch must equal the launch-per-linear results bit for bit, twice in a row."""
    H, I, V, QKV = 4096, 13696, 65024, 4608
    g = torch.Generator(device=DEV).manual_seed(5)

    def lin(k, n, bias=False):
        w = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=DEV, generator=g)
        s = ((torch.rand((k // 32, n), device=DEV, generator=g) * 0.5 + 0.75) / (4.4 * k ** 0.5)).half()
        b = (torch.randn(n, device=DEV, generator=g) * 0.05).half() if bias else None
        return w, s, b

    blocks = [dict(qkv=lin(H, QKV, True), o=lin(H, H), w_in=lin(H, 2 * I), w_out=lin(I, H),
                   ln1=(1 + 0.2 * torch.randn(H, device=DEV, generator=g)).half(),
                   ln2=(1 + 0.2 * torch.randn(H, device=DEV, generator=g)).half()) for _ in range(2)]
    head, lnf = lin(H, V), (1 + 0.2 * torch.randn(H, device=DEV, generator=g)).half()
    x0 = torch.randn(H, device=DEV, generator=g).half()

    def buffers():
        return dict(x=x0.clone(), qkv=torch.zeros(QKV, device=DEV).half(), u=torch.zeros(2 * I, device=DEV).half(),
                    logits=torch.zeros(V, device=DEV).half())

    def chain(emit, bf):
        for b in blocks:
            emit(bf["x"], *b["qkv"][:2], bf["qkv"], bias=b["qkv"][2], prologue=PRO_RMSNORM, norm_weight=b["ln1"], eps=1e-5)
            emit(bf["qkv"], *b["o"][:2], bf["x"], resid=bf["x"])                # attention stand-in: q columns
            emit(bf["x"], *b["w_in"][:2], bf["u"], prologue=PRO_RMSNORM, norm_weight=b["ln2"], eps=1e-5)
            emit(bf["u"], *b["w_out"][:2], bf["x"], resid=bf["x"], prologue=PRO_SILU_GATE)
        emit(bf["x"], *head[:2], bf["logits"], prologue=PRO_RMSNORM, norm_weight=lnf, eps=1e-5)

    ref = buffers()
    with ops.decode_arith(ARITH_SUBNORMAL):     # the program keeps the subnormal-operand f16 arithmetic
        chain(lambda a, w, s, out, **kw: ops.gemv_fused_s4(a[:2 * w.shape[0] * (2 if kw.get("prologue") == PRO_SILU_GATE else 1)],
                                                            w, s, out=out, **kw), ref)
    torch.cuda.synchronize()
    got = buffers()
    prog = ops.DecodeProgram(torch.float16)
    chain(lambda a, w, s, out, **kw: prog.add(a, w, s, out, **kw), got)
    prog.build()
    for rep in range(2):
        got["x"].copy_(x0)
        prog.run()
        workers, failed = prog.status()
        assert not failed and workers % 8 == 0 and workers >= 8
        assert torch.isfinite(got["logits"].float()).all()
        for k in ("x", "qkv", "u", "logits"):
            assert torch.equal(got[k], ref[k]), f"run {rep}: {k} differs from the launch-per-linear chain"


# ------------------------------------------------------------------ one-launch step program (cgq_step_*)
def _w4(seed, k, n, dtype="float16"):
    _, bq, s = make_int4_case(seed, 1, k, n, "Q", dtype)
    return bq, s


@pytest.mark.parametrize("k,n", [(4096, 4608), (4096, 4096), (13696, 4096), (512, 768), (1056, 160), (4096, 65024)])
def test_step_program_linear_prologues_vs_oracle(k, n):
    """Single-phase step programs (one persistent launch each): RMSNorm + bias, plain + residual (in place),
    SiLU-gate prologue, against the numpy oracle -- every real ChatGLM2-6B K / N plus ragged small shapes
    (K not a multiple of the 512-k ring stage, one slice per CTA and fewer)."""
    dt = "float16"
    rng = np.random.default_rng(k + n)
    bq, s = _w4(41, k, n)
    x = orc.round_to(rng.standard_normal(k) * 2.0, dt)
    nw = orc.round_to(1.0 + 0.2 * rng.standard_normal(k), dt)
    bias = orc.round_to(rng.standard_normal(n) * 0.05, dt)
    resid = orc.round_to(rng.standard_normal(n), dt)
    W, S = u8(bq), to_torch(s, dt)
    # (a) RMSNorm prologue + bias
    out = torch.zeros(n, device=DEV, dtype=torch.float16)
    prog = ops.StepProgram()
    prog.linear(to_torch(x, dt), W, S, out, bias=to_torch(bias, dt), prologue=PRO_RMSNORM, norm_weight=to_torch(nw, dt), eps=1e-5)
    prog.run()
    ctas, stages, failed = prog.status()
    assert not failed and ctas >= 1 and stages >= 2
    want = orc.qmatmul_int4(dec.rmsnorm(x, nw, 1e-5, dt)[None], bq, s, bias, dt)[0]
    assert_parity(from_torch(out), want, f"step program rmsnorm+bias K={k} N={n}")
    first = out.clone()
    prog.run()
    torch.cuda.synchronize()
    assert torch.equal(out, first), "two runs of the same program differ"
    # (b) plain prologue, residual added in place
    xr = to_torch(resid, dt)
    prog = ops.StepProgram()
    prog.linear(to_torch(x, dt), W, S, xr, resid=xr)
    prog.run()
    assert not prog.status()[2]
    want = orc.round_to(resid + orc.qmatmul_int4(x[None], bq, s, None, dt)[0], dt)
    assert_parity(from_torch(xr), want, f"step program plain+resid K={k} N={n}")
    # (c) SiLU-gate prologue
    u = orc.round_to(rng.standard_normal(2 * k) * 1.5, dt)
    out = torch.zeros(n, device=DEV, dtype=torch.float16)
    prog = ops.StepProgram()
    prog.linear(to_torch(u, dt), W, S, out, prologue=PRO_SILU_GATE)
    prog.run()
    assert not prog.status()[2]
    want = orc.qmatmul_int4(dec.silu_gate(u, dt)[None], bq, s, None, dt)[0]
    assert_parity(from_torch(out), want, f"step program silu-gate K={k} N={n}")


@pytest.mark.parametrize("k,inner", [(4096, 13696), (256, 384), (512, 1024)])
def test_step_program_swiglu_pair_epilogue_then_w_out(k, inner):
    """w_in with CGQ_EPI_SILU_PAIR followed by w_out (+ residual) in ONE program, two dependent phases with a grid
    barrier between them: u = silu(h) * gate must be BIT-equal to the oracle's roundings of this kernel's own w_in
    output, and the block's result within the bar of the oracle."""
    dt = "float16"
    rng = np.random.default_rng(inner)
    w_in, s_in = _w4(51, k, 2 * inner)
    w_out, s_out = _w4(52, inner, k)
    x = orc.round_to(rng.standard_normal(k), dt)
    resid = orc.round_to(rng.standard_normal(k), dt)
    raw = torch.zeros(2 * inner, device=DEV, dtype=torch.float16)
    plain = ops.StepProgram()
    plain.linear(to_torch(x, dt), u8(w_in), to_torch(s_in, dt), raw)
    plain.run()
    u = torch.zeros(2 * inner, device=DEV, dtype=torch.float16)
    y = to_torch(resid, dt)
    prog = ops.StepProgram()
    prog.linear(to_torch(x, dt), u8(w_in), to_torch(s_in, dt), u, epilogue=ops.EPI_SILU_PAIR)
    prog.linear(u, u8(w_out), to_torch(s_out, dt), y, resid=y, k=inner)
    prog.run()
    assert not prog.status()[2]
    got_u = from_torch(u[:inner])
    want_u = dec.silu_gate(from_torch(raw), dt)
    diff = np.abs(got_u - want_u)
    # device exp / divide vs numpy: the value is rounded to fp16 twice, a last-fp32-bit difference shows rarely
    assert (diff <= np.abs(want_u) * 2.0 ** -9 + 1e-7).all() and (diff == 0).mean() > 0.97, (diff.max(), (diff == 0).mean())
    h = orc.qmatmul_int4(x[None], w_in, s_in, None, dt)[0]
    want = orc.round_to(resid + orc.qmatmul_int4(dec.silu_gate(h, dt)[None], w_out, s_out, None, dt)[0], dt)
    assert_parity(from_torch(y), want, f"step program swiglu block K={k} inner={inner}")


def test_step_program_rejects_what_it_cannot_take():
    x = torch.zeros(64, device=DEV, dtype=torch.float16)
    w = torch.zeros((32, 48), device=DEV, dtype=torch.uint8)          # N = 48 is not a multiple of the 32-column slice
    s = torch.zeros((2, 48), device=DEV, dtype=torch.float16)
    prog = ops.StepProgram()
    prog.linear(x, w, s, torch.zeros(48, device=DEV, dtype=torch.float16))
    with pytest.raises(Exception, match="multiple of 32"):
        prog.build()
    with pytest.raises(TypeError):
        ops.StepProgram(torch.bfloat16)


# ------------------------------------------------------------------ int8 model through the fused step (BASELINE config 4)
@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("k,n", [(4096, 4608), (4096, 4096), (13696, 4096), (4096, 27392), (512, 768), (6848, 4096), (1056, 144)])
def test_gemv_fused_s8_vs_oracle(dtype, k, n):
    """cgq_w8a16_gemv_fused (cluster / DSMEM int8 decode kernel): RMSNorm + bias, plain + residual in place,
    SiLU-gate, on the real layer shapes (incl. K = 6848: a ragged last 128-k stage) against the numpy oracle; the
    plain launch must equal the module-level op bit for bit."""
    from util import make_int8_case

    a, q, s = make_int8_case(61, 1, k, n, "Q", dtype)
    rng = np.random.default_rng(k * 3 + n)
    x = orc.round_to(a[0] * 2.0, dtype)
    nw = orc.round_to(1.0 + 0.2 * rng.standard_normal(k), dtype)
    bias = orc.round_to(rng.standard_normal(n) * 0.05, dtype)
    resid = orc.round_to(rng.standard_normal(n), dtype)
    W, S = torch.from_numpy(q).to(DEV), to_torch(s, dtype)
    got = ops.gemv_fused_s8(to_torch(x, dtype), W, S, bias=to_torch(bias, dtype), prologue=PRO_RMSNORM,
                            norm_weight=to_torch(nw, dtype), eps=1e-5)
    want = orc.qmatmul_int8(dec.rmsnorm(x, nw, 1e-5, dtype)[None], q, s, bias, dtype)[0]
    assert_parity(from_torch(got), want, f"int8 rmsnorm+linear {dtype} K={k} N={n}", rtol=rtol_for(dtype))
    xr = to_torch(resid, dtype)
    got = ops.gemv_fused_s8(to_torch(x, dtype), W, S, resid=xr, out=xr)
    want = orc.round_to(resid + orc.qmatmul_int8(x[None], q, s, None, dtype)[0], dtype)
    assert_parity(from_torch(got), want, f"int8 linear+resid {dtype} K={k} N={n}", rtol=rtol_for(dtype))
    plain = ops.gemv_fused_s8(to_torch(x, dtype), W, S, bias=to_torch(bias, dtype))
    mod = ops.dynamic_quant_matmul(to_torch(x, dtype)[None], W.t(), S, bias=to_torch(bias, dtype))
    assert torch.equal(plain, mod[0])
    u = orc.round_to(rng.standard_normal(2 * k) * 1.5, dtype)
    got = ops.gemv_fused_s8(to_torch(u, dtype), W, S, prologue=PRO_SILU_GATE)
    want = orc.qmatmul_int8(dec.silu_gate(u, dtype)[None], q, s, None, dtype)[0]
    assert_parity(from_torch(got), want, f"int8 silu-gate+linear {dtype} K={k} N={n}", rtol=rtol_for(dtype))


def _random_ref_int8_model(cfg_kwargs, seed=4):
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    from chatglm_q.int8.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q.loader import create_quant_int8_model
    from chatglm_q.model import ChatGLM2Config

    cfg = ChatGLM2Config(**cfg_kwargs)
    with torch.device(DEV):
        model = create_quant_int8_model(cfg, torch.float16)
    g = torch.Generator(device=DEV).manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                n, k = mod.out_features, mod.in_features
                wq = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=DEV, generator=g)
                sc = (torch.rand(n, device=DEV, generator=g) * 0.5 + 0.75) / (73.0 * k ** 0.5)
                b = (torch.randn(n, device=DEV, generator=g) * 0.05).half() if mod.bias is not None else None
                mod.apply_weights_(wq, sc.half(), b)
            elif isinstance(mod, QEmbedding):
                mod.weight.copy_(torch.randint(-127, 128, mod.weight.shape, dtype=torch.int8, device=DEV, generator=g))
                mod.weight_scale.copy_((torch.rand(mod.weight_scale.shape, device=DEV, generator=g) * 0.01 + 0.005).half())
        for name, p in model.named_parameters():
            if name.endswith("ln.weight"):
                p.copy_((1.0 + 0.2 * torch.randn(p.shape, device=DEV, generator=g)).half())
    return model.eval()


@pytest.mark.parametrize("cfg_kwargs", [
    dict(hidden_size=512, inner_hidden_size=1024, head_hidden_size=64, num_multi_query_groups=2, num_attention_heads=8,
         num_layers=3, vocab_size=1024, max_sequence_length=256),
    dict(hidden_size=4096, inner_hidden_size=13696, head_hidden_size=128, num_multi_query_groups=2,
         num_attention_heads=32, num_layers=2, vocab_size=65024, max_sequence_length=512),   # ChatGLM2-6B layer shapes
])
def test_fused_decode_int8_matches_unmodified_reference_model(cfg_kwargs):
    """The int8 model through the fused step, driven as ChatGLMDecoder.generate does, against the UNMODIFIED reference
    int8 model (same kernels behind its QLinear modules) greedily."""
    from chatglm_q_b200.install import install, uninstall

    model = _random_ref_int8_model(cfg_kwargs)
    install("chatglm_q")
    try:
        prompt = torch.tensor([[5, 17, 300, 42, 7, 99, 1000]], device=DEV)
        fused = accelerate(model, max_len=40)
        assert isinstance(fused, FusedDecodeModel) and fused.kind == "w8"
        with torch.no_grad():
            _, lg_e, kv_e = model(input_ids=prompt)
            _, lg_f, kv_f = fused(input_ids=prompt, past_key_values=None)
            assert torch.equal(lg_e, lg_f)
            tok = lg_e[0, -1].argmax().reshape(1, 1)
            for step in range(24):
                _, lg_e, kv_e = model(input_ids=tok, past_key_values=kv_e)
                _, lg_f, kv_f = fused(input_ids=tok, past_key_values=kv_f)
                a, b = lg_e[0, -1].float(), lg_f[0, -1].float()
                assert torch.isfinite(b).all()
                assert_parity(b.cpu().numpy(), a.cpu().numpy(), f"int8 step {step} logits", rtol=2e-2)
                top2 = a.topk(2).values
                if (top2[0] - top2[1]).item() > 2e-2 * a.abs().max().item():
                    assert a.argmax().item() == b.argmax().item(), f"step {step}: greedy token differs"
                tok = a.argmax().reshape(1, 1)
    finally:
        uninstall("chatglm_q")
