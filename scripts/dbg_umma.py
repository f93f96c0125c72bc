import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from chatglm_q_b200 import ops
torch.manual_seed(0)
K, N = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = 'cuda'
def run(a, w, s, impl):
    return ops.dynamic_quant_matmul_s4(a, w, s, impl=impl).float().cpu().numpy()[0]
# case 1: all weights nibble lo=9 hi=9 (q-8 = 1), scale 1, a = ones -> y = K
for name, lo, hi in (("lo=9,hi=9", 9, 9), ("lo=9,hi=8", 9, 8), ("lo=8,hi=9", 8, 9)):
    w = torch.full((K // 2, N), lo | (hi << 4), dtype=torch.uint8, device=dev)
    s = torch.ones((K // 32, N), dtype=torch.float16, device=dev)
    a = torch.ones((1, K), dtype=torch.float16, device=dev)
    y = run(a, w, s, ops.IMPL_GEMV); y0 = run(a, w, s, ops.IMPL_SIMPLE)
    print(name, "umma:", y[:8], "...", y[60:68], y[120:128], " simple:", y0[:2])
# case 2: column-dependent weights: col n has lo nibble = 8 + (n % 7), hi = 8
n_idx = torch.arange(N, device=dev)
w = ((8 + (n_idx % 7)).to(torch.uint8) | (8 << 4)).repeat(K // 2, 1).contiguous()
y = run(a, w, s, ops.IMPL_GEMV); y0 = run(a, w, s, ops.IMPL_SIMPLE)
print("col pattern umma  :", y[:20]); print("col pattern simple:", y0[:20])
bad = np.nonzero(np.abs(y - y0) > 1e-3 * np.abs(y0).max())[0]; print("bad cols:", bad[:64], len(bad))
# case 3: k-dependent activations: a[k] = (k % 5) - 2, weights q-8 = 1 everywhere
w = torch.full((K // 2, N), 0x99, dtype=torch.uint8, device=dev)
a = ((torch.arange(K, device=dev) % 5) - 2).half().reshape(1, K)
print("k pattern umma/simple:", run(a, w, s, ops.IMPL_GEMV)[:4], run(a, w, s, ops.IMPL_SIMPLE)[:4])
# case 4: random
w = torch.randint(0, 256, (K // 2, N), dtype=torch.uint8, device=dev)
a = torch.randn((1, K), device=dev).half()
y = run(a, w, s, ops.IMPL_GEMV); y0 = run(a, w, s, ops.IMPL_SIMPLE)
print("random umma  :", y[:8]); print("random simple:", y0[:8])
bad = np.nonzero(np.abs(y - y0) > 1e-2 * np.abs(y0).max())[0]; print("bad cols:", bad[:64], len(bad))
