"""CPU oracle of the token sampler — TEST INFRASTRUCTURE ONLY.

numpy restatement of `chatglm_q.decoder.top_p_sampling` (reference chatglm_q/decoder.py:12-27).  Only `tests/`,
`__graft_entry__.smoke()` and bench.py's CPU legs may import this module; the product path
(chatglm_q_b200/ops.py::top_p_sampling -> cgq_top_p_sample, include/cgq.h) never does.

Parity is pinned: tests/golden/sampling.npz holds the (probs, indices) pair the UNMODIFIED reference function
hands to torch.multinomial / torch.gather on fixed logits (captured by tests/golden/make_golden_sampling.py,
which imports the real reference from /root/reference), the Exp(1) variates torch.multinomial drew and the token
it returned; tests/test_oracle_golden.py checks this module against them.
"""
from __future__ import annotations

import numpy as np


def top_p_distribution(logits: np.ndarray, top_k: int = 100, top_p: float = 0.8, temperature: float = 1.0):
    """decoder.py:14-22 on one row of float32 logits -> (probs [k] fp32 descending, masked, renormalised;
    indices [k] int64).  Ties keep the lower vocabulary index first (a stable descending sort)."""
    x = logits.astype(np.float32) / np.float32(temperature)                      # :14 logits.float() / temperature
    e = np.exp(x - x.max(), dtype=np.float32)
    probs = e / e.sum(dtype=np.float32)                                          # :14 softmax
    order = np.argsort(-probs, kind="stable")[:top_k]                            # :15-17 sort descending, head
    p = probs[order].copy()
    cumsum = np.cumsum(p, dtype=np.float32)                                      # :20
    p[(cumsum - p) > np.float32(top_p)] = 0.0                                    # :21
    p = p / p.sum(dtype=np.float32)                                              # :22
    return p.astype(np.float32), order.astype(np.int64)


def sample_with(probs: np.ndarray, indices: np.ndarray, q: np.ndarray) -> int:
    """decoder.py:25-26 given the Exp(1) variates torch.multinomial draws for one sample:
    multinomial(probs, 1) == argmax(probs / q) (first maximum), then gather."""
    return int(indices[int(np.argmax(probs.astype(np.float32) / q.astype(np.float32)))])
