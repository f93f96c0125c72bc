import os, sys, torch
sys.path.insert(0, ".")
from chatglm_q_b200 import ops
sys.path.insert(0, "baseline/_ref")
k, n = 4096, 32768
g = torch.Generator(device="cuda").manual_seed(n)
bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device="cuda", generator=g)
s = (torch.rand((k // 32, n), device="cuda", generator=g) * 0.02 - 0.01).half()
w = ops.unpack_int4(bq, s).float()
for m in (8, 5, 4):
    a = torch.randn((m, k), device="cuda", generator=g).half()
    truth = a.float() @ w
    rms = truth.pow(2).mean().sqrt()
    def bad_of(y):
        return ((y.float() - truth).abs() > 1e-2 * truth.abs() + 1e-2 * rms)
    ref = ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
    print(f"M={m}: simple kernel bad={int(bad_of(ref).sum())}")
    for impl, name in ((ops.IMPL_GEMV, "gemv"), (ops.IMPL_GEMV_EXACT, "gemv_exact")):
        for rep in range(3):
            y = ops.dynamic_quant_matmul_s4(a, bq, s, impl=impl)
            b = bad_of(y)
            tiles = sorted(set((b.nonzero()[:, 1] // 128).tolist()))
            per_row = b.sum(1).tolist()
            msg = ""
            if tiles:
                t = tiles[0]
                sl = slice(t * 128, t * 128 + 128)
                ratio = (y.float()[:, sl] / truth[:, sl]).median(dim=1).values.tolist()
                msg = f" first bad tile {t}: median y/truth per row {[round(r, 3) for r in ratio]}"
            print(f"  {name} rep{rep}: bad={int(b.sum())} ntiles={len(tiles)} per_row={per_row}{msg}")
