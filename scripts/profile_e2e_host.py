"""cProfile of the host side of the headline e2e leg (unmodified ChatGLMDecoder.generate on FusedDecodeModel + the
one-launch sampler): where the ~60 us per token between two graph replays go."""
import cProfile
import pstats
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
assert bench.import_reference() is not None
from chatglm_q.decoder import ChatGLMDecoder  # noqa: E402
from chatglm_q_b200.fused_decode import FusedDecodeModel  # noqa: E402
from chatglm_q_b200.install import install, uninstall  # noqa: E402

cfg, model = bench.build_ref_int4_model(torch, dev)
install("chatglm_q", sampler=True)
try:
    fused = FusedDecodeModel(model, max_len=32 + 256 + 32, alias_logits=True)
    dec = ChatGLMDecoder(cfg, fused, bench.StubTokenizer(32), device=dev, time_log=False)
    torch.manual_seed(0)
    for _ in dec.generate("warm-up", max_generated_tokens=8):
        pass
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    n = 0
    for _ in dec.generate("bench", max_generated_tokens=256):
        n += 1
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(18)
    print("tokens", n)
finally:
    uninstall("chatglm_q")
