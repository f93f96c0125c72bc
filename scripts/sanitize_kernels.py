"""Small launches of every decode-path kernel that tools/chainbench does not reach, for compute-sanitizer
(scripts/gpu_sanitize.sh): the M = 2..8 exact decode kernel and (CGQ_GEMV_TRICK_MGT1=1) its subnormal-operand
variant on a multi-wave grid, the int8 decode kernel, the sampler, the embedding / unpack kernels.  Each result is
checked against the CUDA-core kernel so that a sanitizer-clean run is also a correct one."""
import sys

import torch

sys.path.insert(0, ".")
from chatglm_q_b200 import ops  # noqa: E402

dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)


def close(y, ref, what):
    err = (y.float() - ref.float()).abs()
    rms = ref.float().pow(2).mean().sqrt()
    bad = err > 1e-2 * ref.float().abs() + 1e-2 * rms
    print(f"{what}: {'ok' if not bad.any() else 'MISMATCH ' + str(int(bad.sum()))}", flush=True)


k = 4096
for n in (65024, 4608):          # one multi-wave Z=1 grid, one cluster-of-8 grid
    bq = torch.randint(0, 256, (k // 2, n), dtype=torch.uint8, device=dev, generator=g)
    s = (torch.rand((k // 32, n), device=dev, generator=g) * 0.02 - 0.01).half()
    for m in (1, 2, 5, 8):
        a = torch.randn((m, k), device=dev, generator=g).half()
        ref = ops.dynamic_quant_matmul_s4(a, bq, s, impl=ops.IMPL_SIMPLE)
        ys = [ops.dynamic_quant_matmul_s4(a, bq, s) for _ in range(3)]
        torch.cuda.synchronize()
        close(ys[0], ref, f"w4 gemv M={m} N={n}")
        assert all(torch.equal(y, ys[0]) for y in ys), f"M={m} N={n}: launches differ"
# int8 decode kernel
n = 4608
q8 = torch.randint(-128, 128, (n, k), dtype=torch.int8, device=dev, generator=g)
s8 = (torch.randn(n, device=dev, generator=g) / 2048).half()
for m in (1, 8):
    a = torch.randn((m, k), device=dev, generator=g).half()
    ref = ops.dynamic_quant_matmul(a, q8.t(), s8, impl=ops.IMPL_SIMPLE)
    y = ops.dynamic_quant_matmul(a, q8.t(), s8)
    torch.cuda.synchronize()
    close(y, ref, f"w8 gemv M={m} N={n}")
# sampler
logits = torch.randn(65024, device=dev, generator=g).half()
torch.manual_seed(0)
t = ops.top_p_sampling(logits, 100, 0.8, 1.0)
torch.cuda.synchronize()
print("sampler token", int(t))
print("done")
