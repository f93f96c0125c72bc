"""Device / wall time of the one-launch sampler against the reference's torch sampler (same logits, same GPU)."""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from chatglm_q_b200 import _lib, ops  # noqa: E402

dev = "cuda"
logits = (torch.randn(65024, generator=torch.Generator().manual_seed(0)) * 3).half().to(dev)
lib = _lib.load()
q = torch.empty(100, device=dev).exponential_(1)
tok = torch.empty(1, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
for top_k in (100, 1024):
    qq = torch.empty(top_k, device=dev).exponential_(1)
    for _ in range(5):
        lib.cgq_top_p_sample(logits.data_ptr(), 65024, 0, top_k, 0.8, 1.0, qq.data_ptr(), tok.data_ptr(), None, None, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        lib.cgq_top_p_sample(logits.data_ptr(), 65024, 0, top_k, 0.8, 1.0, qq.data_ptr(), tok.data_ptr(), None, None, stream)
    e1.record()
    torch.cuda.synchronize()
    print(f"cgq_top_p_sample V=65024 top_k={top_k}: {e0.elapsed_time(e1) * 1e3 / 200:.2f} us/launch (back to back)")


def wall(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e6


print(f"ops.top_p_sampling(...).item(): {wall(lambda: ops.top_p_sampling(logits).item()):.1f} us/call wall")
ref = ROOT / "baseline" / "_ref"
if (ref / "chatglm_q").exists():
    sys.path.insert(0, str(ref))
    import chatglm_q.decoder as dec

    print(f"reference top_p_sampling(...).item(): {wall(lambda: dec.top_p_sampling(logits).item()):.1f} us/call wall")
