"""Stress of the M = 8 decode kernel after an L2-warming launch (MODE=simple_each): how the ring-release race of
DESIGN.md §3.1 was reproduced; tests/test_gpu_parity.py::test_decode_kernel_stress is the permanent check."""
import os, sys, torch
sys.path.insert(0, ".")
from chatglm_q_b200 import ops
k = 4096
for n in (40960, 32768):
    g = torch.Generator(device="cuda").manual_seed(n)
    bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device="cuda", generator=g)
    s = (torch.rand((k // 32, n), device="cuda", generator=g) * 0.02 - 0.01).half()
    w = ops.unpack_int4(bq, s).float()
    a = torch.randn((8, k), device="cuda", generator=g).half()
    truth = a.float() @ w
    rms = truth.pow(2).mean().sqrt()
    shown = 0
    mode = os.environ.get("MODE", "")
    if mode == "simple_once":
        ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
    nbad = 0
    for rep in range(30):
        if mode == "simple_each":
            ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
        if mode == "sync_each":
            torch.cuda.synchronize()
        y = ops.dynamic_quant_matmul_s4(a, bq, s).float()
        b = ((y - truth).abs() > 1e-2 * truth.abs() + 1e-2 * rms)
        nbad += int(b.any())
        if b.any() and shown < 4:
            shown += 1
            cols = sorted(set(b.nonzero()[:, 1].tolist()))
            tiles = sorted(set(c // 128 for c in cols))
            for t in tiles[:3]:
                cc = [c % 128 for c in cols if c // 128 == t]
                err = (y - truth)[:, t * 128:(t + 1) * 128]
                # which k-group explains the error? project the error of row 0 on each group's contribution
                contrib = torch.stack([a[0, gg * 32:(gg + 1) * 32].float() @ w[gg * 32:(gg + 1) * 32, t * 128:(t + 1) * 128] for gg in range(128)])
                e0 = err[0]
                cmask = torch.zeros(128, dtype=torch.bool, device="cuda"); cmask[cc] = True
                score = ((contrib[:, cmask] + e0[cmask]).abs().sum(1))   # err == -contrib[g] if group g was dropped
                score2 = ((contrib[:, cmask] - e0[cmask]).abs().sum(1))  # err == +contrib[g] if group g was doubled
                print(f"N={n} rep={rep} tile={t} bad cols(mod128)={cc[:24]}{'...' if len(cc)>24 else ''} n={len(cc)} "
                      f"max|err|/rms={float(err.abs().max()/rms):.3f} best dropped-group={int(score.argmin())} "
                      f"(resid {float(score.min()):.3f} vs |e| {float(e0[cmask].abs().sum()):.3f}) doubled-group={int(score2.argmin())} (resid {float(score2.min()):.3f})")
    print(f"MODE={mode} N={n}: {nbad}/30 bad launches")
