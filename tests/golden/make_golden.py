"""Generate tests/golden/*.npz by running the REAL reference (imported from /root/reference) on CPU.

Run once in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures pin oracle/qmatmul_oracle.py and oracle/cgq_oracle.c (tests/test_oracle_golden.py) and are
replayed against the CUDA kernels (tests/test_gpu_parity.py).  Inputs come from numpy PCG64 seeds so a
test can regenerate anything too large to store.  bfloat16 tensors are stored as uint16 bit patterns.
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
from chatglm_q.int4 import qlinear as q4  # noqa: E402
from chatglm_q.int4.quantizer import quantize_int4  # noqa: E402
from chatglm_q.int8 import qlinear as q8  # noqa: E402
from chatglm_q.int8.quantizer import quantize_int8  # noqa: E402

OUT = Path(__file__).resolve().parent
TD = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def store(t: torch.Tensor) -> np.ndarray:
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def main():
    torch.set_num_threads(4)
    rng = np.random.default_rng(20240601)
    g = {}

    # --- (1) SURVEY §8(c) known-answer vectors, recomputed by the reference itself
    g["ka1_out"] = q4.unpack_int4(torch.tensor([[0xA3]], dtype=torch.uint8), torch.tensor([[0.5]])).numpy()
    g["ka2_out"] = q4.unpack_int4(torch.tensor([[0x00], [0xFF]], dtype=torch.uint8), torch.tensor([[1.0]])).numpy()
    b3 = np.full((32, 2), 0x88, dtype=np.uint8)
    b3[15] = [50, 68]
    b3[16] = [215, 215]
    s3 = np.array([[1, 2], [10, 20]], dtype=np.float32)
    g["ka3_bytes"], g["ka3_scale"] = b3, s3
    g["ka3_out"] = q4.unpack_int4(torch.from_numpy(b3), torch.from_numpy(s3)).numpy()

    # --- (2) unpack_int4 on random bytes, every activation dtype
    ub = rng.integers(0, 256, size=(64, 48), dtype=np.uint8)
    us = (rng.random((4, 48)) * 0.04 - 0.02).astype(np.float32)
    us[1, :8] = 0.0
    g["unpack_bytes"], g["unpack_scale_f32"] = ub, us
    for name, td in TD.items():
        st = torch.from_numpy(us).to(td)
        g[f"unpack_scale_{name}"] = store(st)
        g[f"unpack_out_{name}"] = store(q4.unpack_int4(torch.from_numpy(ub), st))
    g["unpack_i8"] = (((torch.from_numpy(ub).reshape(64, 1, 48).repeat(1, 2, 1)
                        >> torch.tensor([0, 4], dtype=torch.uint8).reshape(1, 2, 1)) & 0xF).to(torch.int8) - 8
                      ).reshape(128, 48).numpy()

    # --- (3) quantisers
    w4 = (rng.standard_normal((128, 40)) / 8).astype(np.float32)
    w4[32:64, 3] = 0.0  # an all-zero group -> scale clamp 1e-10
    qb, qs = quantize_int4(torch.from_numpy(w4))
    g["q4_in"], g["q4_bytes"], g["q4_scale"] = w4, qb.numpy(), qs.numpy()
    w8 = (rng.standard_normal((24, 96)) / 8).astype(np.float32)
    q8b, q8s = quantize_int8(torch.from_numpy(w8))
    g["q8_in"], g["q8_q"], g["q8_scale"] = w8, q8b.numpy(), q8s.numpy()

    # --- (4) int4 DynamicQuantizeLinear.forward on the reference CPU path, with bias
    M, K, N = 5, 128, 48
    x = rng.standard_normal((M, K)).astype(np.float32)
    wq, ws = quantize_int4(torch.from_numpy((rng.standard_normal((K, N)) / np.sqrt(K)).astype(np.float32)))
    bias = (rng.standard_normal(N) * 0.1).astype(np.float32)
    g["l4_bytes"] = wq.numpy()
    for name, td in TD.items():
        lin = q4.DynamicQuantizeLinear(K, N, bias=True, dtype=td)
        lin.apply_weights_(wq, ws.to(td), torch.from_numpy(bias).to(td))
        xt = torch.from_numpy(x).to(td)
        with torch.no_grad():
            y = lin(xt)
        g[f"l4_x_{name}"], g[f"l4_scale_{name}"] = store(xt), store(lin.weight_scale)
        g[f"l4_bias_{name}"], g[f"l4_y_{name}"] = store(lin.bias), store(y)

    # --- (4b) the reference's own test shape (tests/test_triton_ops_int4.py:11-22), fp32
    a = rng.standard_normal((32, 512)).astype(np.float32)
    b = (rng.standard_normal((512, 256)) / np.sqrt(512)).astype(np.float32)
    bq, bs = quantize_int4(torch.from_numpy(b))
    g["t4_a"], g["t4_bytes"], g["t4_scale"] = a, bq.numpy(), bs.numpy()
    g["t4_y"] = (torch.from_numpy(a) @ q4.unpack_int4(bq, bs)).numpy()

    # --- (5) int8: the reference's own test (tests/test_triton_ops.py:9-17: signed scales), fp32
    A = rng.standard_normal((10, 128)).astype(np.float32)
    B = rng.integers(-127, 127, size=(128, 256), dtype=np.int8)   # [K, N] as in the test
    Bs = (rng.standard_normal(256) / 256).astype(np.float32)
    g["t8_a"], g["t8_b_kn"], g["t8_scale"] = A, B, Bs
    g["t8_y"] = (torch.from_numpy(A) @ (torch.from_numpy(B) * torch.from_numpy(Bs))).numpy()
    # int8 module forward with bias, 16-bit dtypes
    M, K, N = 7, 96, 24
    x8 = rng.standard_normal((M, K)).astype(np.float32)
    b8 = (rng.standard_normal(N) * 0.1).astype(np.float32)
    g["l8_q"] = q8b.numpy()
    for name, td in TD.items():
        lin = q8.DynamicQuantizeLinear(K, N, bias=True, dtype=td)
        lin.apply_weights_(q8b, q8s.to(td), torch.from_numpy(b8).to(td))
        xt = torch.from_numpy(x8).to(td)
        with torch.no_grad():
            y = lin(xt)
        g[f"l8_x_{name}"], g[f"l8_scale_{name}"] = store(xt), store(lin.weight_scale)
        g[f"l8_bias_{name}"], g[f"l8_y_{name}"] = store(lin.bias), store(y)

    # --- (6) QEmbedding (int4 packs along the vocab axis; int8 scales per feature)
    V, D = 64, 40
    ew = (rng.standard_normal((V, D))).astype(np.float32)
    eq, es = quantize_int4(torch.from_numpy(ew))
    ids = torch.from_numpy(rng.integers(0, V, size=(2, 9)).astype(np.int64))
    emb4 = q4.QEmbedding(V, D, dtype=torch.float16)
    emb4.apply_weights_(eq, es.half())
    g["e4_ids"], g["e4_bytes"], g["e4_scale"] = ids.numpy(), eq.numpy(), store(es.half())
    g["e4_y"] = store(emb4(ids))
    e8q = rng.integers(-127, 128, size=(V, D), dtype=np.int8)
    e8s = (rng.random(D) * 0.02).astype(np.float32)
    emb8 = q8.QEmbedding(V, D, dtype=torch.float16)
    emb8.apply_weights_(torch.from_numpy(e8q), torch.from_numpy(e8s).half())
    g["e8_q"], g["e8_scale"], g["e8_y"] = e8q, store(torch.from_numpy(e8s).half()), store(emb8(ids))

    np.savez_compressed(OUT / "reference_vectors.npz", **g)

    # --- (7) BASELINE.json config 1: int8 QLinear forward (128,4096)x(4096,4096), reference CPU path.
    # Inputs are regenerated from the seed by the test; only a sub-sample of the output is stored.
    r1 = np.random.default_rng(1)
    W = (r1.standard_normal((4096, 4096)) / 64).astype(np.float32)
    X = r1.standard_normal((128, 4096)).astype(np.float32)
    q, s = quantize_int8(torch.from_numpy(W))
    lin = q8.DynamicQuantizeLinear(4096, 4096, bias=False, dtype=torch.float32)
    lin.apply_weights_(q, s)
    with torch.no_grad():
        Y = lin(torch.from_numpy(X)).numpy()
    exact = np.array_equal(Y, (torch.from_numpy(X) @ (q.t() * s)).numpy())
    relerr = float(np.linalg.norm(Y - X @ W.T) / np.linalg.norm(X @ W.T))
    np.savez_compressed(OUT / "config1_int8.npz", y_sub=Y[::16, ::64], q_sub=q.numpy()[::64, ::64],
                        s_sub=s.numpy()[::64], exact=np.array(exact), relerr=np.array(relerr),
                        y_checksum=np.array(np.float64(Y.astype(np.float64).sum())))
    print("config1: exact == x @ (q.t()*s):", exact, " rel err vs unquantised:", relerr)
    for f in OUT.glob("*.npz"):
        print(f.name, f.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
