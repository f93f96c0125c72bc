"""Generate tests/golden/backward.npz: grad_A of the REAL reference's `DynamicQuantizeMatMul` (autograd through its
torch path on CPU, int4/qlinear.py:53-64 and int8/qlinear.py:41-52), imported from /root/reference.

    python tests/golden/make_golden_backward.py
Pins oracle.qmatmul_oracle.qmatmul_int4_grad_a / qmatmul_int8_grad_a (tests/test_oracle_golden.py) and is replayed
against cgq_w4a16_grad_a / cgq_w8a16_grad_a on the GPU (tests/test_gpu_parity.py).  fp32 on CPU: the reference's own
tests use fp32 and 1e-4 (tests/test_triton_ops_int4.py:24-37)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from chatglm_q.int4 import qlinear as q4  # noqa: E402
from chatglm_q.int8 import qlinear as q8  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    torch.set_num_threads(4)
    rng = np.random.default_rng(20241017)
    g = {}
    for tag, (m, k, n) in {"a": (5, 128, 48), "b": (10, 256, 208)}.items():
        a = rng.standard_normal((m, k)).astype(np.float32)
        go = rng.standard_normal((m, n)).astype(np.float32)
        bq = rng.integers(0, 256, size=(k // 2, n), dtype=np.uint8)
        bs = (rng.random((k // 32, n)) * 0.04 - 0.02).astype(np.float32)
        at = torch.from_numpy(a).requires_grad_()
        out = q4.dynamic_quant_matmul(at, torch.from_numpy(bq), torch.from_numpy(bs))
        out.backward(torch.from_numpy(go))
        g[f"s4_{tag}_grad_out"], g[f"s4_{tag}_b"], g[f"s4_{tag}_scale"] = go, bq, bs
        g[f"s4_{tag}_grad_a"] = at.grad.numpy()
        w = rng.integers(-128, 128, size=(n, k), dtype=np.int8)
        ws = (rng.standard_normal(n) * 0.02).astype(np.float32)        # signed scales, as tests/test_triton_ops.py
        at = torch.from_numpy(a).requires_grad_()
        out = q8.dynamic_quant_matmul(at, torch.from_numpy(w).t(), torch.from_numpy(ws))
        out.backward(torch.from_numpy(go))
        g[f"s8_{tag}_w"], g[f"s8_{tag}_scale"] = w, ws
        g[f"s8_{tag}_grad_a"] = at.grad.numpy()
    np.savez_compressed(OUT / "backward.npz", **g)
    print({k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
