"""Generate tests/golden/decode_tiny.npz by running the REAL reference model (imported from
/root/reference) on CPU in float16: a tiny random int4g32 ChatGLM2, a prompt (prefill) and greedy
decode steps through the UNMODIFIED `ChatGLM2Model.forward` with `past_key_values`.

Run once in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden_decode.py
The fixture pins oracle/decode_oracle.py (tests/test_oracle_golden.py) and is replayed against the
fused CUDA decode step (tests/test_gpu_parity.py).
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding  # noqa: E402
from chatglm_q.int4.quantizer import quantize_int4  # noqa: E402
from chatglm_q.loader import create_quant_int4_model  # noqa: E402
from chatglm_q.model import ChatGLM2Config  # noqa: E402

OUT = Path(__file__).resolve().parent
CFG = dict(hidden_size=256, inner_hidden_size=384, head_hidden_size=64, num_multi_query_groups=2,
           num_attention_heads=4, num_layers=2, vocab_size=256, max_sequence_length=64)
PROMPT = [5, 17, 200, 42, 7]
STEPS = 6


def main():
    torch.manual_seed(7)
    torch.set_num_threads(4)
    cfg = ChatGLM2Config(**CFG)
    model = create_quant_int4_model(cfg, 32, torch.float16)
    model.eval()
    g = {}
    names = {"attn.qkv_proj": "qkv", "attn.o_proj": "o", "ffn.w_in": "win", "ffn.w_out": "wout"}
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, DynamicQuantizeLinear):
                k, n = mod.in_features, mod.out_features
                q, s = quantize_int4(torch.randn(k, n) / k ** 0.5)
                bias = (torch.randn(n) * 0.05).half() if mod.bias is not None else None
                mod.apply_weights_(q, s.half(), bias)
                if name == "lm_head":
                    key = "lm"
                else:
                    _, idx, a, b = name.split(".")
                    key = f"l{idx}_{names[a + '.' + b]}"
                g[key + "_w"], g[key + "_s"] = q.numpy(), s.half().numpy()
                if bias is not None:
                    g[key + "_b"] = bias.numpy()
            elif isinstance(mod, QEmbedding):
                q, s = quantize_int4(torch.randn(cfg.vocab_size, cfg.hidden_size))
                mod.apply_weights_(q, s.half())
                g["emb_w"], g["emb_s"] = q.numpy(), s.half().numpy()
        for i, layer in enumerate(model.layers):
            for ln in ("attn_ln", "ffn_ln"):
                w = getattr(layer, ln).weight
                w.copy_((1.0 + 0.2 * torch.randn_like(w.float())).half())
                g[f"l{i}_{ln}"] = w.detach().numpy()
        model.final_ln.weight.copy_((1.0 + 0.2 * torch.randn(cfg.hidden_size)).half())
        g["final_ln"] = model.final_ln.weight.detach().numpy()
        g["freqs"] = model.freqs_cis_cache.numpy()

        ids = torch.tensor([PROMPT])
        _, logits, kv = model(input_ids=ids)
        g["prompt"] = np.array(PROMPT, dtype=np.int64)
        g["prefill_logits_last"] = logits[0, -1].numpy()
        for i, (k, v) in enumerate(kv):
            g[f"prefill_k{i}"], g[f"prefill_v{i}"] = k[0, :, :, 0].numpy(), v[0, :, :, 0].numpy()
        toks, step_logits = [], []
        tok = int(logits[0, -1].argmax())
        for _ in range(STEPS):
            toks.append(tok)
            _, logits, kv = model(input_ids=torch.tensor([[tok]]), past_key_values=kv)
            step_logits.append(logits[0, -1].numpy())
            tok = int(logits[0, -1].argmax())
        g["step_tokens"] = np.array(toks, dtype=np.int64)
        g["step_logits"] = np.stack(step_logits)
        for i, (k, v) in enumerate(kv):
            g[f"final_k{i}"], g[f"final_v{i}"] = k[0, :, :, 0].numpy(), v[0, :, :, 0].numpy()
    g["cfg"] = np.array([cfg.num_attention_heads, cfg.num_multi_query_groups, cfg.head_hidden_size, cfg.num_layers,
                         cfg.hidden_size, cfg.inner_hidden_size, cfg.vocab_size, cfg.max_sequence_length])
    g["eps"] = np.array(cfg.layernorm_epsilon)
    np.savez_compressed(OUT / "decode_tiny.npz", **g)
    print("wrote", OUT / "decode_tiny.npz", sum(v.nbytes for v in g.values()), "bytes raw")


if __name__ == "__main__":
    main()
