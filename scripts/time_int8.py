"""int8 (per-channel) dequant-matmul timings on the ChatGLM2-6B shapes (BASELINE.json config 4): decode M=1 / 8 through
w8_gemv_kernel, prefill M=2048 through the tcgen05 kernel; weights rotated over > 3x L2 so they come from HBM."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from chatglm_q_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
shapes = [("qkv", 4096, 4608, 28), ("o_proj", 4096, 4096, 28), ("w_in", 4096, 27392, 28), ("w_out", 13696, 4096, 28),
          ("lm_head", 4096, 65024, 1)]
for M in (1, 8, 2048):
    tok_us, tok_bytes = 0.0, 0
    for name, K, N, per_token in shapes:
        copies = max(2, int(420e6 // (K * N)) + 1)
        ws = [torch.randint(-127, 128, (N, K), dtype=torch.int8, device=dev, generator=g) for _ in range(copies)]
        sc = (torch.rand(N, device=dev, generator=g) * 0.01 + 0.001).half()
        a = torch.randn(M, K, device=dev, generator=g).half()
        reps = 10 if M > 8 else 30
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up on the capture stream: workspace, tensor maps
            for i in range(copies):
                ops.dynamic_quant_matmul(a, ws[i].t(), sc)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):     # one replay = one launch per weight copy, no host gaps
                for i in range(copies):
                    ops.dynamic_quant_matmul(a, ws[i].t(), sc)
        torch.cuda.current_stream().wait_stream(side)
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * copies)
        del graph
        nbytes = K * N + 2 * N + 2 * M * K + 2 * M * N
        print(f"int8 M={M} {name} K={K} N={N}: {us:.2f} us  {nbytes / us / 1e3:.1f} GB/s  {2 * M * N * K / us / 1e6:.2f} TFLOP/s")
        tok_us += us * per_token
        tok_bytes += nbytes * per_token
        del ws
    print(f"int8 M={M} all linears of one step (graph-replayed launches, sum of the above): {tok_us / 1e3:.3f} ms, "
          f"{tok_bytes / 1e9:.3f} GB, {tok_bytes / tok_us / 1e3:.0f} GB/s, {M * 1e6 / tok_us:.0f} tok/s")
