"""Stress the M<=8 decode kernel (multi-wave grids) against the CUDA-core kernel; count bad launches."""
import os, sys, torch
sys.path.insert(0, ".")
from chatglm_q_b200 import ops
k = 4096
for n in (65024, 40960, 32768):
    g = torch.Generator(device="cuda").manual_seed(n)
    bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device="cuda", generator=g)
    s = (torch.rand((k // 32, n), device="cuda", generator=g) * 0.02 - 0.01).half()
    for m in (8, 7, 6, 5, 3):
        a = torch.randn((m, k), device="cuda", generator=g).half()
        ref = ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE).float()
        rms = ref.pow(2).mean().sqrt()
        nbad, tiles = 0, set()
        for rep in range(12):
            y = ops.dynamic_quant_matmul_s4(a, bq, s).float()
            bad = ((y - ref).abs() > 1e-2 * ref.abs() + 1e-2 * rms)
            if bad.any():
                nbad += 1
                tiles |= set((bad.nonzero()[:, 1] // 128).tolist())
        print(f"{os.environ.get('TAG','')} N={n} M={m}: {nbad}/12 launches bad, tiles {sorted(tiles)[:10]}")
