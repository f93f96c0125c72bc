"""decode_attn_kernel in isolation: a CUDA graph of 200 launches (same qkv, 28 rotating KV caches), per-launch time
against context length.  In the fused step the kernel sits between the qkv and o_proj launches; this is its own cost."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from chatglm_q_b200 import ops  # noqa: E402

dev = "cuda"
NH, NG, DH, L = 32, 2, 128, 28
for max_len, ctx in ((256, 8), (256, 96), (256, 200), (2048, 1000), (4096, 3000)):
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn((NH + 2 * NG) * DH, device=dev, generator=g).half()
    freqs = torch.randn(max_len + 8, DH, device=dev, generator=g).half()
    kc = [torch.randn(max_len, NG, DH, device=dev, generator=g).half() for _ in range(L)]
    vc = [torch.randn(max_len, NG, DH, device=dev, generator=g).half() for _ in range(L)]
    state = torch.tensor([ctx + 1, ctx, 0], dtype=torch.int32, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(L):
            ops.decode_attention(qkv, freqs, kc[i], vc[i], state, NH, NG, DH)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for r in range(8):
                for i in range(L):
                    ops.decode_attention(qkv, freqs, kc[i], vc[i], state, NH, NG, DH)
    torch.cuda.current_stream().wait_stream(side)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(20):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"decode_attn max_len={max_len} ctx={ctx}: {e0.elapsed_time(e1) * 1e3 / (20 * 8 * L):.2f} us per launch (back to back, PDL)")
