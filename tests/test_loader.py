"""Load-time path: chatglm_q_b200.loader against the reference's own save format and loader
contract (chatglm_q/loader.py:90-156) -- CPU only, tiny random int4g32 model built by the reference's factory."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"


def _tiny_model(seed):
    if not (REF / "chatglm_q").exists():
        pytest.skip("baseline/_ref (pip-installed reference) not present")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    pytest.importorskip("safetensors")
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding
    from chatglm_q.int4.quantizer import quantize_int4
    from chatglm_q.loader import create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config

    torch.manual_seed(seed)
    cfg = ChatGLM2Config(hidden_size=256, inner_hidden_size=512, head_hidden_size=64, num_multi_query_groups=2,
                         num_attention_heads=4, num_layers=2, vocab_size=512, max_sequence_length=64)
    model = create_quant_int4_model(cfg, 32, torch.float16)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                q, s = quantize_int4(torch.randn(mod.in_features, mod.out_features) / mod.in_features ** 0.5)
                mod.apply_weights_(q, s.half(), torch.randn(mod.out_features).half() if mod.bias is not None else None)
            elif isinstance(mod, QEmbedding):
                q, s = quantize_int4(torch.randn(cfg.vocab_size, cfg.hidden_size))
                mod.apply_weights_(q, s.half())
        for n, p in model.named_parameters():
            if n.endswith("ln.weight"):
                p.copy_((1 + 0.1 * torch.randn(p.shape)).half())
    return cfg, model


def test_load_state_into_matches_the_reference_loader_contract(tmp_path, capsys):
    from safetensors.torch import save_file

    from chatglm_q_b200 import loader

    cfg, src = _tiny_model(1)
    sd = src.state_dict()
    names = sorted(sd)
    half = len(names) // 2                       # two shard files, like save_model_and_tokenizer(shard=True)
    save_file({k: sd[k] for k in names[:half]}, tmp_path / "model_weights_0.safetensors")
    extra = dict({k: sd[k] for k in names[half:]}, **{"not.in.model": torch.zeros(3)})
    dropped = names[-1]
    extra.pop(dropped)
    save_file(extra, tmp_path / "model_weights_1.safetensors")
    _, dst = _tiny_model(2)
    missing = loader.load_state_into(dst, [tmp_path / "model_weights_0.safetensors", tmp_path / "model_weights_1.safetensors"])
    printed = capsys.readouterr().out
    assert missing == [dropped] and '"not.in.model" is ignored' in printed and "are not initialized" in printed
    got = dst.state_dict()
    for k in names:
        if k != dropped:
            assert torch.equal(got[k], sd[k]), k
            assert got[k].dtype == sd[k].dtype


@pytest.mark.parametrize("world", [2, 4])
def test_load_tp_shards_equals_sharding_the_loaded_model(tmp_path, world):
    from safetensors.torch import save_file

    from chatglm_q_b200 import loader, tp

    cfg, src = _tiny_model(3)
    sd = src.state_dict()
    save_file(dict(sd), tmp_path / "w.safetensors")
    dims = tp.ModelDims(cfg.hidden_size, cfg.inner_hidden_size, cfg.head_hidden_size, cfg.num_multi_query_groups,
                        cfg.num_attention_heads, cfg.vocab_size)
    for rank in range(world):
        plan = tp.plan_block(world, rank, dims)
        got = loader.load_tp_shards([tmp_path / "w.safetensors"], dims, cfg.num_layers, world, rank, "cpu")
        for i, layer in enumerate(src.layers):
            for name, lin, sh in (("attn.qkv_proj", layer.attn.qkv_proj, plan.qkv), ("attn.o_proj", layer.attn.o_proj, plan.o),
                                  ("ffn.w_in", layer.ffn.w_in, plan.w_in), ("ffn.w_out", layer.ffn.w_out, plan.w_out)):
                w, s, b = tp.shard_w4(lin.weight, lin.weight_scale, lin.bias, sh, rank=0)
                assert torch.equal(got[f"layers.{i}.{name}.weight"], w)
                assert torch.equal(got[f"layers.{i}.{name}.weight_scale"], s)
                if b is not None:
                    assert torch.equal(got[f"layers.{i}.{name}.bias"], b)
        w, s, _ = tp.shard_w4(src.lm_head.weight, src.lm_head.weight_scale, None, plan.lm_head, rank=0)
        assert torch.equal(got["lm_head.weight"], w) and torch.equal(got["lm_head.weight_scale"], s)
        assert torch.equal(got["word_embedding.weight"], sd["word_embedding.weight"])       # replicated
