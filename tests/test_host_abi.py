"""CPU suite: the C-ABI library loads and exports every symbol include/cgq.h declares, host-side
validation mirrors the reference's AssertionErrors, the module mirrors keep the reference's
state-dict contract, and the product never routes through the oracle."""
import re
import types
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "cgq.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cgq_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from chatglm_q_b200 import _lib

    lib = _lib.load()  # no compute call: needs no GPU
    names = header_symbols()
    assert len(names) >= 11
    for name in names:
        assert hasattr(lib, name), f"libcgq.so does not export {name}"
        assert name in _lib.SYMBOLS, f"ctypes binding lacks {name}"
    assert lib.cgq_version() >= 1
    assert lib.cgq_workspace_bytes() > 0


def test_abi_argument_errors_without_gpu():
    """Bad arguments are rejected before any CUDA work (status codes of include/cgq.h)."""
    from chatglm_q_b200 import _lib

    lib = _lib.load()
    P = 0x10000  # never dereferenced: validation fails first
    # group != 32
    rc = lib.cgq_w4a16_gemm(P, 64, P, P, None, P, 8, 1, 8, 64, 16, 0, None, 0, None)
    assert rc == -1 and b"group" in lib.cgq_last_error()
    # dtype code
    rc = lib.cgq_w4a16_gemm(P, 64, P, P, None, P, 8, 1, 8, 64, 32, 7, None, 0, None)
    assert rc == -2
    rc = lib.cgq_w8a16_gemm(P, 64, P, P, None, P, 8, 1, 8, -5, 0, None, 0, None)
    assert rc == -1
    rc = lib.cgq_w4_unpack_i8(P, P, 7, 8, None)  # odd K
    assert rc == -1
    with pytest.raises(_lib.CgqError):
        _lib.check(rc)
    # sampler: top_k < 1, temperature <= 0, dtype code, top_k > 1024, row too large, token without variates
    assert lib.cgq_top_p_sample(P, 1000, 0, 0, 0.8, 1.0, None, None, P, P, None) == -1
    assert lib.cgq_top_p_sample(P, 1000, 0, 10, 0.8, 0.0, None, None, P, P, None) == -1
    assert lib.cgq_top_p_sample(P, 1000, 7, 10, 0.8, 1.0, None, None, P, P, None) == -2
    assert lib.cgq_top_p_sample(P, 5000, 0, 2000, 0.8, 1.0, None, None, P, P, None) == -1
    assert lib.cgq_top_p_sample(P, 200000, 0, 10, 0.8, 1.0, None, None, P, P, None) == -1
    assert lib.cgq_top_p_sample(P, 1000, 0, 10, 0.8, 1.0, None, P, None, None, None) == -1
    assert b"variates" in lib.cgq_last_error()
    # backward: group != 32, output row shorter than K, dtype code; the decode-arithmetic / simple-kernel switches
    # are pure host state
    assert lib.cgq_w4a16_grad_a(P, 16, P, P, P, 64, 2, 16, 64, 16, 0, None) == -1 and b"group" in lib.cgq_last_error()
    assert lib.cgq_w4a16_grad_a(P, 16, P, P, P, 32, 2, 16, 64, 32, 0, None) == -1
    assert lib.cgq_w8a16_grad_a(P, 16, P, P, P, 64, 2, 16, 64, 9, None) == -2
    prev = lib.cgq_set_decode_arith(_lib.ARITH_SUBNORMAL)
    assert lib.cgq_set_decode_arith(-1) == _lib.ARITH_SUBNORMAL          # query only
    assert lib.cgq_set_decode_arith(prev) == _lib.ARITH_SUBNORMAL and lib.cgq_set_decode_arith(-1) == prev
    was = lib.cgq_forbid_simple(1)
    assert lib.cgq_forbid_simple(was) == 1 and isinstance(lib.cgq_simple_fallback_count(), int)


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]|qmatmul_oracle|liboracle", re.M)
    for f in (ROOT / "chatglm_q_b200").rglob("*"):
        if f.suffix in {".py", ".cu", ".cuh", ".h"}:
            assert not pat.search(f.read_text()), f"{f} references the oracle"


def test_sampler_host_checks():
    """ops.top_p_sampling keeps the reference signature (decoder.py:12) and refuses what the kernel cannot take."""
    import inspect

    import torch

    from chatglm_q_b200 import ops

    sig = inspect.signature(ops.top_p_sampling)
    assert [(n, p.default) for n, p in sig.parameters.items() if p.kind is p.POSITIONAL_OR_KEYWORD] == [
        ("logits", inspect.Parameter.empty), ("top_k", 100), ("top_p", 0.8), ("temperature", 1.0)]
    with pytest.raises(AssertionError):
        ops.top_p_sampling(torch.zeros(100, dtype=torch.float16))          # CPU tensor: no fallback


def test_host_checks_mirror_reference_asserts():
    from chatglm_q_b200 import ops

    a = torch.zeros(2, 64, dtype=torch.float16)
    b = torch.zeros(32, 16, dtype=torch.uint8)
    s = torch.zeros(2, 16, dtype=torch.float16)
    assert ops.check_input(a) is False  # CPU tensor: the reference would take its torch fallback
    with pytest.raises(AssertionError):  # int4/triton_ops.py:109 `assert a.get_device() >= 0`
        ops.dynamic_quant_matmul_s4(a, b, s)
    with pytest.raises(AssertionError):  # :104 shape mismatch
        ops.dynamic_quant_matmul_s4(torch.zeros(2, 60, dtype=torch.float16), b, s)
    with pytest.raises(AssertionError):  # :106 dtype of B
        ops.dynamic_quant_matmul_s4(a, b.to(torch.int8), s)
    with pytest.raises(AssertionError):  # :107 a.dtype == b_scale.dtype
        ops.dynamic_quant_matmul_s4(a, b, s.float())
    with pytest.raises(AssertionError):  # int8/triton_ops.py:98 scale must be 1-D
        ops.dynamic_quant_matmul(a, torch.zeros(64, 16, dtype=torch.int8), s)


def test_module_mirrors_keep_state_dict_contract():
    from chatglm_q_b200 import int4, int8

    m4 = int4.DynamicQuantizeLinear(4096, 4608, bias=True, dtype=torch.float16, device="meta")
    sd = m4.state_dict()
    assert list(sd) == ["weight", "weight_scale", "bias"]
    assert sd["weight"].shape == (2048, 4608) and sd["weight"].dtype == torch.uint8
    assert sd["weight_scale"].shape == (128, 4608) and sd["weight_scale"].dtype == torch.float16
    assert sd["bias"].shape == (4608,)
    assert not list(m4.parameters())  # buffers, not parameters (int4/qlinear.py:83-88)
    m4n = int4.DynamicQuantizeLinear(13696, 4096, bias=False, dtype=torch.float16, device="meta")
    assert list(m4n.state_dict()) == ["weight", "weight_scale"] and m4n.bias is None
    assert m4n.weight_scale.shape == (428, 4096)
    with pytest.raises(AssertionError):
        int4.DynamicQuantizeLinear(100, 8)  # in_features % group_size
    m8 = int8.DynamicQuantizeLinear(4096, 27392, bias=False, dtype=torch.float16, device="meta")
    assert m8.weight.shape == (27392, 4096) and m8.weight.dtype == torch.int8
    assert m8.weight_scale.shape == (27392,)
    e4 = int4.QEmbedding(65024, 4096, dtype=torch.float16, device="meta")
    assert e4.weight.shape == (32512, 4096) and e4.weight_scale.shape == (2032, 4096)
    e8 = int8.QEmbedding(65024, 4096, dtype=torch.float16, device="meta")
    assert e8.weight.shape == (65024, 4096) and e8.weight_scale.shape == (4096,)
    # apply_weights_ fills in place (quantiser scripts rely on it)
    m = int4.DynamicQuantizeLinear(64, 16, bias=True, dtype=torch.float16)
    m.apply_weights_(torch.full((32, 16), 0x88, dtype=torch.uint8), torch.ones(2, 16), torch.zeros(16))
    assert int(m.weight[0, 0]) == 0x88 and float(m.weight_scale[1, 3]) == 1.0


def test_install_rebinds_reference_globals():
    """Seam S1: chatglm_q.int4.qlinear._dynamic_quant_matmul_impl / check_input are module globals."""
    import sys

    from chatglm_q_b200 import install, ops

    pkg = "fake_chatglm_q"
    mods = {}
    for name in (pkg, f"{pkg}.int4", f"{pkg}.int4.qlinear", f"{pkg}.int8", f"{pkg}.int8.qlinear"):
        mods[name] = types.ModuleType(name)
        mods[name].__path__ = []
    sentinel = object()
    for leaf in (f"{pkg}.int4.qlinear", f"{pkg}.int8.qlinear"):
        mods[leaf]._dynamic_quant_matmul_impl = sentinel
        mods[leaf].check_input = None
        mods[leaf].KERNEL_IMPL = "none"
    sys.modules.update(mods)
    try:
        install.install(pkg)
        assert mods[f"{pkg}.int4.qlinear"]._dynamic_quant_matmul_impl is ops.dynamic_quant_matmul_s4
        assert mods[f"{pkg}.int8.qlinear"]._dynamic_quant_matmul_impl is ops.dynamic_quant_matmul
        assert mods[f"{pkg}.int4.qlinear"].check_input is ops.check_input
        assert mods[f"{pkg}.int4.qlinear"].KERNEL_IMPL == "cgq_b200"
        install.uninstall(pkg)
        assert mods[f"{pkg}.int4.qlinear"]._dynamic_quant_matmul_impl is sentinel
        assert mods[f"{pkg}.int8.qlinear"].KERNEL_IMPL == "none"
        # decoder.top_p_sampling is a module global too (decoder.py:12, resolved at :85): opt-in rebind
        dec = types.ModuleType(f"{pkg}.decoder")
        dec.top_p_sampling = sentinel
        sys.modules[dec.__name__] = mods[dec.__name__] = dec
        install.install(pkg)
        assert dec.top_p_sampling is sentinel
        install.install(pkg, sampler=True)
        assert dec.top_p_sampling is ops.top_p_sampling
        install.uninstall(pkg)
        assert dec.top_p_sampling is sentinel
        bound = lambda logits, top_k=100, top_p=0.8, temperature=1.0: None   # noqa: E731  e.g. FusedDecodeModel.sampler()
        install.install(pkg, sampler=bound)
        assert dec.top_p_sampling is bound
        install.uninstall(pkg)
        assert dec.top_p_sampling is sentinel
    finally:
        for name in mods:
            sys.modules.pop(name, None)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from chatglm_q_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libcgq.so")
    with pytest.raises(ImportError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_fused_decode_wrapper_host_logic():
    """FusedDecodeModel / accelerate() decide on the host which models the fused step can take (no GPU needed):
    int4g32 buffers + fp16/bf16 + head size 64/128 -> fused; anything else -> TypeError / the graph wrapper."""
    from types import SimpleNamespace as NS

    import torch

    from chatglm_q_b200.fused_decode import FusedDecodeModel, accelerate
    from chatglm_q_b200.graph_decode import GraphDecodeModel

    def lin(k, n, wdtype=torch.uint8, sdtype=torch.float16):
        w = torch.zeros((k // 2, n), dtype=wdtype) if wdtype == torch.uint8 else torch.zeros((n, k), dtype=wdtype)
        s = torch.zeros((k // 32, n), dtype=sdtype) if wdtype == torch.uint8 else torch.zeros(n, dtype=sdtype)
        return NS(weight=w, weight_scale=s, bias=None)

    def model(head=64, **kw):
        cfg = NS(hidden_size=256, inner_hidden_size=384, head_hidden_size=head, num_multi_query_groups=2,
                 num_attention_heads=256 // head, num_layers=1, vocab_size=256, max_sequence_length=64)
        layer = NS(attn_ln=NS(weight=torch.ones(256), eps=1e-5), ffn_ln=NS(weight=torch.ones(256), eps=1e-5),
                   attn=NS(qkv_proj=lin(256, 256 + 4 * head, **kw), o_proj=lin(256, 256, **kw)),
                   ffn=NS(w_in=lin(256, 768, **kw), w_out=lin(384, 256, **kw)))
        return NS(config=cfg, layers=[layer], final_ln=NS(weight=torch.ones(256), eps=1e-5), lm_head=lin(256, 256, **kw),
                  word_embedding=NS(weight=torch.zeros((128, 256), dtype=torch.uint8),
                                    weight_scale=torch.zeros((8, 256), dtype=torch.float16)),
                  freqs_cis_cache=torch.zeros((64, head), dtype=torch.float16))

    fused = FusedDecodeModel(model(), max_len=1000)
    assert fused.max_len == 63 and fused.launches_per_step() == 7          # window clipped to the rotary table
    # opt-in modes are off by default; the bound sampler defers to ops.top_p_sampling unless a fused step is live
    assert not fused.speculate and not fused.handover and not fused._spec
    spec = FusedDecodeModel(model(), max_len=32, speculate=True)
    assert spec.speculate and callable(spec.sampler())
    assert not spec._can_speculate(torch.zeros(256, dtype=torch.float16))   # nothing captured yet
    assert isinstance(accelerate(model()), FusedDecodeModel)
    for bad in (model(wdtype=torch.int8), model(sdtype=torch.float32), model(head=32)):
        try:
            FusedDecodeModel(bad)
        except TypeError:
            pass
        else:
            raise AssertionError("FusedDecodeModel accepted a model the fused step cannot take")
        assert isinstance(accelerate(bad), GraphDecodeModel)


def test_fused_wrapper_prefill_last_position_only():
    """FusedDecodeModel(last_logits_only=True): the prefill goes through the UNMODIFIED reference forward with lm_head
    applied to the last position only.  Runs the real reference model on CPU (its torch path):
    logits [1, 1, V] bit-equal to the last row of the full forward, the module tree restored afterwards."""
    import sys

    import torch

    ref = ROOT / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q.int4.quantizer import quantize_int4
    from chatglm_q.loader import create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config

    from chatglm_q_b200.fused_decode import FusedDecodeModel

    torch.manual_seed(3)
    cfg = ChatGLM2Config(hidden_size=128, inner_hidden_size=256, head_hidden_size=64, num_multi_query_groups=2,
                         num_attention_heads=2, num_layers=2, vocab_size=256, max_sequence_length=64)
    model = create_quant_int4_model(cfg, 32, torch.float32)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                q, s = quantize_int4(torch.randn(mod.in_features, mod.out_features) / mod.in_features ** 0.5)
                mod.apply_weights_(q, s, torch.zeros(mod.out_features) if mod.bias is not None else None)
            elif isinstance(mod, QEmbedding):
                q, s = quantize_int4(torch.randn(cfg.vocab_size, cfg.hidden_size))
                mod.apply_weights_(q, s)
    model.eval()
    # fp32 CPU model: the fused STEP cannot take it (TypeError), the prefill wrapper logic is dtype-agnostic
    with pytest.raises(TypeError):
        FusedDecodeModel(model)
    for m in model.modules():
        if isinstance(m, (DynamicQuantizeLinear, QEmbedding)):
            m.weight_scale.data = m.weight_scale.data.half()
    model.half()
    wrapped = FusedDecodeModel(model, max_len=32, last_logits_only=True)
    head = model.lm_head
    ids = torch.tensor([[5, 17, 200, 42, 7]])
    with torch.no_grad():
        try:
            _, full, _ = model(input_ids=ids)
        except RuntimeError as e:           # a CPU build of torch without half matmul: nothing to compare
            pytest.skip(f"reference CPU forward in fp16 unavailable: {e}")
        _, last, handle = wrapped(input_ids=ids, past_key_values=None)
    assert last.shape == (1, 1, cfg.vocab_size) and full.shape == (1, 5, cfg.vocab_size)
    assert torch.equal(last[0, 0], full[0, -1])
    assert model.lm_head is head and wrapped.n_valid == 5


def _tiny_ref_int4_model(torch):
    """Random tiny int4g32 ChatGLM2 built by the reference's own factory, fp16 on CPU (its torch path)."""
    import sys

    ref = ROOT / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q.int4.quantizer import quantize_int4
    from chatglm_q.loader import create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config

    torch.manual_seed(5)
    cfg = ChatGLM2Config(hidden_size=128, inner_hidden_size=256, head_hidden_size=64, num_multi_query_groups=2,
                         num_attention_heads=2, num_layers=2, vocab_size=256, max_sequence_length=64)
    model = create_quant_int4_model(cfg, 32, torch.float32)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                q, s = quantize_int4(torch.randn(mod.in_features, mod.out_features) / mod.in_features ** 0.5)
                mod.apply_weights_(q, s, torch.zeros(mod.out_features) if mod.bias is not None else None)
            elif isinstance(mod, QEmbedding):
                q, s = quantize_int4(torch.randn(cfg.vocab_size, cfg.hidden_size))
                mod.apply_weights_(q, s)
    for m in model.modules():
        if isinstance(m, (DynamicQuantizeLinear, QEmbedding)):
            m.weight_scale.data = m.weight_scale.data.half()
    model.half().eval()
    try:
        with torch.no_grad():
            model(input_ids=torch.tensor([[1, 2]]))
    except RuntimeError as e:
        pytest.skip(f"reference CPU forward in fp16 unavailable: {e}")
    return cfg, model


@pytest.mark.parametrize("wrapper", ["fused", "graph"])
def test_wrappers_batch_and_long_prompt_stay_on_the_reference_path(wrapper):
    """calls the static one-row window cannot take -- a batch of 2, a prompt longer than the window, a
    multi-token call after either -- must run the unmodified model on the cache it returned ("correct, just not
    fused"), keep `n_valid` in step, and never touch a stale static cache.  Real reference model on CPU."""
    import torch

    from chatglm_q_b200.fused_decode import FusedDecodeModel
    from chatglm_q_b200.graph_decode import GraphDecodeModel

    cfg, model = _tiny_ref_int4_model(torch)
    make = (lambda: FusedDecodeModel(model, max_len=8)) if wrapper == "fused" else (lambda: GraphDecodeModel(model, max_len=8))

    def reference(calls):
        kv, outs = None, []
        with torch.no_grad():
            for ids in calls:
                _, lg, kv = model(input_ids=ids, past_key_values=kv)
                outs.append(lg)
        return outs

    def through(w, calls):
        h, outs = None, []
        for ids in calls:
            _, lg, h = w(input_ids=ids, past_key_values=h)
            outs.append(lg)
        return outs

    # (a) batch of two: prefill, two single-token steps, a two-token call
    calls = [torch.tensor([[5, 17, 200], [9, 3, 4]]), torch.tensor([[7], [8]]), torch.tensor([[1], [2]]),
             torch.tensor([[3, 4], [5, 6]])]
    w = make()
    for got, want in zip(through(w, calls), reference(calls)):
        assert torch.equal(got, want)
    assert w.n_valid == 7
    # (b) prompt longer than the window: eager single-token steps, then a two-token call, then one more token
    calls = [torch.arange(1, 13).reshape(1, 12), torch.tensor([[7]]), torch.tensor([[8]]), torch.tensor([[9, 10]]),
             torch.tensor([[11]])]
    w = make()
    for got, want in zip(through(w, calls), reference(calls)):
        assert torch.equal(got, want)
    assert w.n_valid == 17
    # (c) a handle from an earlier session is not accepted as the current cache
    w = make()
    _, _, old = w(input_ids=torch.arange(1, 13).reshape(1, 12), past_key_values=None)
    _, _, new = w(input_ids=torch.arange(20, 32).reshape(1, 12), past_key_values=None)
    _, lg_stale, _ = w(input_ids=torch.tensor([[7]]), past_key_values=old)      # == a fresh one-token prefill
    assert torch.equal(lg_stale, reference([torch.tensor([[7]])])[0])
