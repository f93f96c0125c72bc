"""A few launches of the tcgen05 prefill kernel at one shape, for ncu (M, N from argv; K = 4096)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from chatglm_q_b200 import ops  # noqa: E402

M, N, K = int(sys.argv[1]), int(sys.argv[2]), 4096
g = torch.Generator(device="cuda").manual_seed(0)
w = torch.randint(0, 256, (K // 2, N), dtype=torch.uint8, device="cuda", generator=g)
s = (torch.rand((K // 32, N), device="cuda", generator=g) * 0.02 - 0.01).half()
a = torch.randn(M, K, device="cuda", generator=g).half()
for _ in range(4):
    ops.dynamic_quant_matmul_s4(a, w, s)
torch.cuda.synchronize()
