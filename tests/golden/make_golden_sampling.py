"""Generate tests/golden/sampling.npz from the REAL reference sampler (imported from /root/reference).

`chatglm_q.decoder.top_p_sampling` (decoder.py:12-27) returns only the sampled token, so the UNMODIFIED function
is run with torch.multinomial / torch.gather wrapped to record what it hands them: the (probs, indices) pair,
and — by re-seeding — the Exp(1) variates multinomial draws for itself (`empty_like(probs).exponential_(1)`,
checked here: argmax(probs / q) reproduces multinomial's choice on every case).

Run once in the build container (the reference does not exist on the GPU box):
    python tests/golden/make_golden_sampling.py
The fixture pins oracle/sampling_oracle.py (tests/test_oracle_golden.py) and is replayed against the CUDA
sampler (tests/test_gpu_sampling.py).
"""
import sys
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
import chatglm_q.decoder as dec  # noqa: E402

OUT = Path(__file__).resolve().parent

# name, dtype, V, generator of fp32 logits, top_k, top_p, temperature
CASES = [
    ("vocab_f16", torch.float16, 65024, lambda g, v: torch.randn(v, generator=g) * 2.0, 100, 0.8, 1.0),
    ("vocab_bf16", torch.bfloat16, 65024, lambda g, v: torch.randn(v, generator=g) * 3.0, 100, 0.8, 1.0),
    ("peaked_f16", torch.float16, 65024, lambda g, v: torch.randn(v, generator=g) * 6.0, 50, 0.5, 0.7),
    ("ties_f16", torch.float16, 4099, lambda g, v: torch.randint(-3, 4, (v,), generator=g).float() * 0.5, 100, 0.95, 1.0),
    ("small_f16", torch.float16, 37, lambda g, v: torch.randn(v, generator=g), 100, 0.8, 1.3),
    ("flat_bf16", torch.bfloat16, 1000, lambda g, v: torch.zeros(v), 64, 0.9, 1.0),
    ("negative_f16", torch.float16, 2048, lambda g, v: -torch.rand(v, generator=g) * 20.0 - 1.0, 10, 0.99, 2.0),
]


def main():
    out = {}
    names = []
    for i, (name, dt, v, make, top_k, top_p, temp) in enumerate(CASES):
        g = torch.Generator().manual_seed(100 + i)
        logits = make(g, v).to(dt)
        seen = {}
        real_multinomial, real_gather = torch.multinomial, torch.gather

        def multinomial(probs, num_samples=1, **kw):
            seen["probs"] = probs.clone()
            state = torch.get_rng_state()
            seen["q"] = torch.empty_like(probs).exponential_(1)
            torch.set_rng_state(state)
            choice = real_multinomial(probs, num_samples=num_samples, **kw)
            seen["choice"] = choice.clone()
            return choice

        def gather(inp, dim, index, **kw):
            seen["indices"] = inp.clone()
            return real_gather(inp, dim, index, **kw)

        torch.manual_seed(500 + i)
        torch.multinomial, torch.gather = multinomial, gather
        try:
            token = dec.top_p_sampling(logits, top_k, top_p, temp)      # the unmodified reference function
        finally:
            torch.multinomial, torch.gather = real_multinomial, real_gather
        probs, q, idx = seen["probs"], seen["q"], seen["indices"]
        assert int(torch.argmax(probs / q)) == int(seen["choice"]), name   # multinomial == argmax(p / q)
        assert int(idx[int(seen["choice"])]) == int(token), name
        bits = logits.view(torch.int16).numpy().view(np.uint16)
        out[f"{name}_logits_bits"] = bits
        out[f"{name}_params"] = np.array([top_k, top_p, temp], dtype=np.float64)
        out[f"{name}_probs"] = probs.numpy().astype(np.float32)
        out[f"{name}_indices"] = idx.numpy().astype(np.int64)
        out[f"{name}_q"] = q.numpy().astype(np.float32)
        out[f"{name}_token"] = np.array(int(token), dtype=np.int64)
        names.append(f"{name}:{'bfloat16' if dt == torch.bfloat16 else 'float16'}")
        print(name, v, "token", int(token), "kept", int((probs > 0).sum()), "of", probs.numel())
    out["cases"] = np.array(names)
    np.savez_compressed(OUT / "sampling.npz", **out)
    print("wrote", OUT / "sampling.npz", (OUT / "sampling.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
