"""Device time of the fused decode step of the full-size random ChatGLM2-6B int4g32 model, default protocol
against the experimental tile-granular hand-over (CGQ_HANDOVER / FusedDecodeModel(handover=True))."""
import os
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
prompt = torch.tensor([bench.StubTokenizer(32).encode("x")], device=dev)
ref_logits = None
for hand in (False, True):
    fm = FusedDecodeModel(model, max_len=256, handover=hand)
    with torch.no_grad():
        _, lg, kv = fm(input_ids=prompt, past_key_values=None)
        tok = lg[0, -1].argmax().reshape(1, 1)
        outs = []
        for _ in range(8):
            _, lg, kv = fm(input_ids=tok, past_key_values=kv)
            outs.append(lg.clone())
            tok = lg[0, -1].argmax().reshape(1, 1)
    torch.cuda.synchronize()
    if ref_logits is None:
        ref_logits = outs
    same = all(torch.equal(a, b) for a, b in zip(ref_logits, outs))
    st = torch.tensor([96, 96], dtype=torch.int32, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 100
    for i in range(reps + 10):
        if i == 10:
            e0.record()
        fm.state[:2].copy_(st, non_blocking=True)
        fm.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"[mode={os.environ.get('CGQ_HAND_MODE', '0')} sleep={os.environ.get('CGQ_HAND_SLEEP', '20')} mask={os.environ.get('CGQ_HAND_MASK', '7')}] "
          f"fused step handover={hand}: {e0.elapsed_time(e1) * 1e3 / reps:.1f} us/token, logits identical to default: {same}")
    del fm
