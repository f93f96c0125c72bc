"""CPU model of the selection algorithm of csrc/sampling.cu (not the kernel: the ARITHMETIC of its design).

The kernel finds the top_k head of the descending sort without sorting: order-preserving 16-bit keys, a threshold
from the k-th largest of the 1 024 per-thread maxima (fast path), an exact two-digit radix select over all keys when
more than 1 024 elements reach that threshold (slow path), ties broken by the lower vocabulary id, ranking by
counting.  This file restates those steps in numpy with the kernel's ownership map (warp-contiguous segments,
two adjacent elements per lane and step) and checks them against oracle/sampling_oracle.py — so a change of the
design is caught here, without a GPU; the kernel itself is checked against the same oracle in test_gpu_sampling.py.
"""
import numpy as np
import pytest

from oracle import qmatmul_oracle as orc
from oracle import sampling_oracle as so

THREADS, WARPS, MAX_TOPK = 1024, 32, 1024


def keys_of(bits: np.ndarray) -> np.ndarray:
    b = bits.astype(np.uint32)
    return np.where(b & 0x8000, (~b) & 0xFFFF, b | 0x8000).astype(np.uint32)


def select_bin(hist: np.ndarray, k: int):
    """bin (from the top) in which the running count reaches k, and the count above it (select_bin in the kernel)"""
    above = 0
    for b in range(255, -1, -1):
        if above < k <= above + hist[b]:
            return b, above
        above += int(hist[b])
    raise AssertionError("fewer than k elements")


def radix_select(keys: np.ndarray, k: int):
    """threshold key t with count(key > t) < k <= count(key >= t), and count(key > t)"""
    b1, above1 = select_bin(np.bincount(keys >> 8, minlength=256), k)
    inside = keys[(keys >> 8) == b1]
    b2, above2 = select_bin(np.bincount(inside & 0xFF, minlength=256), k - above1)
    return (b1 << 8) | b2, above1 + above2


def kernel_model(bits: np.ndarray, top_k: int):
    """-> (ids of the top-k head in the kernel's order, which path was taken)"""
    V = len(bits)
    k = min(top_k, V)
    unit = 64 * WARPS
    vpad = (V + unit - 1) // unit * unit
    keys = np.zeros(vpad, np.uint32)
    keys[:V] = keys_of(bits)
    valid = np.arange(vpad) < V
    seg = vpad // WARPS
    # ownership: element e of warp w, step s, lane l, half h is w*seg + s*64 + 2*l + h
    e = np.arange(vpad)
    owner = (e // seg) * 32 + ((e % seg) % 64) // 2
    tmax = np.zeros(THREADS, np.uint32)
    np.maximum.at(tmax, owner[valid], keys[valid])
    tau, _ = radix_select(tmax, k)                       # k <= 1024 thread maxima always exist (invalid = key 0)
    cand = np.nonzero(valid & (keys >= tau))[0]
    path = "fast"
    if len(cand) > MAX_TOPK:
        path = "slow"
        t, c_gt = radix_select(keys[:V], k)
        gt = np.nonzero(valid & (keys > t))[0]
        ties = np.nonzero(valid & (keys == t))[0][: k - c_gt]          # first `need` ties in index order
        assert len(gt) == c_gt
        cand = np.concatenate([gt, ties])
    assert len(cand) >= k
    comp = (keys[cand].astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - cand.astype(np.uint64))
    rank = (comp[None, :] > comp[:, None]).sum(axis=1)                 # ranking by counting
    head = np.empty(k, np.int64)
    keep = rank < k
    head[rank[keep]] = cand[keep]
    return head, path


def _bits(x: np.ndarray, dtype: str) -> np.ndarray:
    x = orc.round_to(x.astype(np.float32), dtype)
    if dtype == "float16":
        return x.astype(np.float16).view(np.uint16)
    return (x.view(np.uint32) >> 16).astype(np.uint16)


def _f32(bits: np.ndarray, dtype: str) -> np.ndarray:
    return orc.bf16_from_bits(bits) if dtype == "bfloat16" else bits.view(np.float16).astype(np.float32)


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
@pytest.mark.parametrize("v,top_k,want_path", [(65024, 100, "fast"), (65024, 1, "fast"), (65024, 1024, "slow"),
                                               (65023, 50, "fast"), (2049, 100, "fast"), (9, 100, "fast"),
                                               (1, 5, "fast"), (70000, 40, "fast")])
def test_selection_design_matches_oracle(dtype, v, top_k, want_path):
    rng = np.random.default_rng(v * 3 + top_k)
    bits = _bits(rng.standard_normal(v) * 2.5, dtype)
    head, path = kernel_model(bits, top_k)
    _, want = so.top_p_distribution(_f32(bits, dtype), top_k, 1.0, 1.0)
    assert np.array_equal(head, want)
    assert path == want_path


def test_selection_design_ties_and_masked_vocabulary():
    zeros = np.zeros(65024, np.float32)
    head, path = kernel_model(_bits(zeros, "float16"), 100)
    assert path == "slow" and np.array_equal(head, np.arange(100))     # every element ties: lowest ids
    x = zeros.copy()
    x[40000] = 1.0
    x[5:20] = -1.0
    head, _ = kernel_model(_bits(x, "float16"), 100)
    assert np.array_equal(head, np.concatenate([[40000], np.arange(5), np.arange(20, 114)]))
    few = np.random.default_rng(1).integers(-3, 4, 4099).astype(np.float32) * 0.5   # 7 distinct values
    head, path = kernel_model(_bits(few, "float16"), 100)
    _, want = so.top_p_distribution(few, 100, 1.0, 1.0)
    assert np.array_equal(head, want) and path == "fast"
    masked = np.full(65024, -np.inf, np.float32)
    masked[[7, 99, 64000]] = [0.5, 2.0, 1.0]
    head, _ = kernel_model(_bits(masked, "float16"), 3)
    assert head.tolist() == [99, 64000, 7]


def test_order_preserving_keys():
    for dtype in ("float16", "bfloat16"):
        bits = np.arange(65536, dtype=np.uint32).astype(np.uint16)
        vals = _f32(bits, dtype)
        ok = np.isfinite(vals) | np.isinf(vals)
        ok &= ~np.isnan(vals)
        k, v = keys_of(bits)[ok].astype(np.int64), vals[ok].astype(np.float64)
        order = np.argsort(k, kind="stable")
        assert (np.diff(v[order]) >= 0).all()            # keys order like the numbers (-0 below +0, equal values)
