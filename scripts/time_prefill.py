"""Prefill of 2 048 tokens through the UNMODIFIED reference forward with this repo's tcgen05 kernels behind its QLinear
modules (full-size random ChatGLM2-6B int4g32), all-position logits against FusedDecodeModel(last_logits_only=True)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

assert bench.import_reference() is not None, "baseline/_ref missing"
from chatglm_q_b200.fused_decode import FusedDecodeModel  # noqa: E402
from chatglm_q_b200.install import install  # noqa: E402

dev = torch.device("cuda:0")
install("chatglm_q")
cfg, model = bench.build_ref_int4_model(torch, dev)
L = 2048
ids = torch.randint(1000, 60000, (1, L), generator=torch.Generator().manual_seed(0)).to(dev)
last = {}
for flag in (False, True):
    fm = FusedDecodeModel(model, max_len=L + 64, last_logits_only=flag)
    with torch.no_grad():
        for _ in range(2):
            _, lg, kv = fm(input_ids=ids, past_key_values=None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            _, lg, kv = fm(input_ids=ids, past_key_values=None)
        e1.record()
        torch.cuda.synchronize()
    last[flag] = lg[0, -1].clone()
    print(f"prefill {L} tokens, last_logits_only={flag}: {e0.elapsed_time(e1) / 3:.2f} ms  logits {tuple(lg.shape)}  "
          f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")
    del fm, lg, kv
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
print("last-position logits identical:", torch.equal(last[False], last[True]),
      "max |diff|", float((last[False].float() - last[True].float()).abs().max()))
