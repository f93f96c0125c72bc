"""tcgen05 prefill kernel at small M: the same weights every launch (L2-resident) against rotated copies (HBM) --
tells a memory-latency-bound pipeline from a synchronisation-bound one."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from chatglm_q_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
K = 4096
for M in (16, 128):
    for N in (4608, 27392):
        copies = max(2, int(420e6 // (K * N // 2)) + 1)
        ws = [(torch.randint(0, 256, (K // 2, N), dtype=torch.uint8, device=dev, generator=g),
               (torch.rand((K // 32, N), device=dev, generator=g) * 0.02 - 0.01).half()) for _ in range(copies)]
        a = torch.randn(M, K, device=dev, generator=g).half()
        for label, pick in (("HBM (rotated copies)", lambda i: ws[i % copies]), ("L2 (same weights)", lambda i: ws[0])):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(copies):
                    ops.dynamic_quant_matmul_s4(a, *pick(i))
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for i in range(copies):
                        ops.dynamic_quant_matmul_s4(a, *pick(i))
            torch.cuda.current_stream().wait_stream(side)
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(10):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            print(f"M={M} N={N} {label}: {e0.elapsed_time(e1) * 1e3 / (10 * copies):.2f} us")
            del graph
        del ws
