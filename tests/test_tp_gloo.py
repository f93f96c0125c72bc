"""CPU tests of the tensor-parallel shard plan (chatglm_q_b200/tp.py): plan arithmetic for
world 1/2/4/8 and the N>1 data path under `gloo`, world_size 2 — every rank computes its shard of a block's
linears with the oracle, row-parallel partials are summed with tp.all_reduce_sum, the column-parallel lm_head
is re-assembled with tp.gather_columns, and the result must match the single-rank oracle."""
import os
import socket

import numpy as np
import pytest
import torch

from chatglm_q_b200 import tp
from oracle import qmatmul_oracle as orc
from util import assert_parity


def test_plan_matches_survey_numbers():
    for world, qkv_n, win_n in ((1, 4608, 27392), (2, 2304, 13696), (4, 1280, 6848), (8, 768, None)):
        plans = [tp.plan_block(world, r) for r in range(world)]
        assert all(p.qkv.n_out(4608) == qkv_n for p in plans)
        if win_n is not None:
            assert all(p.w_in.n_out(27392) == win_n for p in plans)
        assert sum(p.o.k_in(4096) for p in plans) == 4096 * (1 if world > 1 else 1)
        assert sum(p.w_out.k_in(13696) for p in plans) == 13696
        assert sum(p.lm_head.n_out(65024) for p in plans) == 65024
        assert plans[0].allreduces_per_block == (0 if world == 1 else 2)
        for p in plans:                                  # row splits sit on group boundaries
            for sh in (p.o, p.w_out):
                if sh.krows is not None:
                    assert sh.krows[0] % 32 == 0 and sh.krows[1] % 32 == 0
    groups = [p.w_out.k_in(13696) // 32 for p in (tp.plan_block(8, r) for r in range(8))]
    assert groups == [54, 54, 54, 54, 53, 53, 53, 53]    # 428 groups do not divide by 8
    # every head of a rank sits in one KV group; K/V columns replicated on the ranks that need them
    p5 = tp.plan_block(8, 5)
    assert p5.heads == (20, 24) and p5.kv_groups == (1, 2)
    assert p5.qkv.cols == ((2560, 3072), (4096 + 128, 4096 + 256), (4352 + 128, 4352 + 256))


DIMS = tp.ModelDims(hidden_size=256, inner_hidden_size=384, head_hidden_size=64, num_multi_query_groups=2,
                    num_attention_heads=4, vocab_size=256)


def _weights():
    rng = np.random.default_rng(11)
    H, I, V = DIMS.hidden_size, DIMS.inner_hidden_size, DIMS.vocab_size
    qkv_n = DIMS.head_hidden_size * (DIMS.num_attention_heads + 2 * DIMS.num_multi_query_groups)
    shapes = {"qkv": (H, qkv_n), "o": (H, H), "w_in": (H, 2 * I), "w_out": (I, H), "lm_head": (H, V)}
    w = {}
    for name, (k, n) in shapes.items():
        bq, s = orc.quantize_int4((rng.standard_normal((k, n)) / np.sqrt(k)).astype(np.float32))
        bias = orc.round_to(rng.standard_normal(n) * 0.05, "float16") if name == "qkv" else None
        w[name] = (bq, orc.round_to(s, "float16"), bias)
    x = orc.round_to(rng.standard_normal((3, H)), "float16")
    return w, x


def _block(w, x, plan, rank, reduce, gather):
    """The five linears of a block on one rank (attention / activation replaced by fixed slices, as in
    bench.py's TokenStep): returns (x_out, logits)."""
    def lin(name, a, sh):
        bq, s, b = tp.shard_w4(torch.from_numpy(w[name][0]), torch.from_numpy(w[name][1]),
                               None if w[name][2] is None else torch.from_numpy(w[name][2]), sh, rank)
        y = orc.qmatmul_int4(a, bq.numpy(), s.numpy(), None if b is None else b.numpy(), "float16")
        return torch.from_numpy(y)

    H, I = DIMS.hidden_size, DIMS.inner_hidden_size
    qkv = lin("qkv", x, plan.qkv)                                       # column-parallel: rank-local heads first
    o = reduce(lin("o", qkv[:, :plan.o.k_in(H)].numpy(), plan.o))       # row-parallel + all-reduce
    hin = lin("w_in", orc.round_to(o.numpy(), "float16"), plan.w_in)
    half = plan.w_in.n_out(2 * I) // 2
    x2 = reduce(lin("w_out", hin[:, :half].numpy(), plan.w_out))
    logits = gather(lin("lm_head", orc.round_to(x2.numpy(), "float16"), plan.lm_head), plan.lm_head, DIMS.vocab_size)
    return x2, logits


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        w, x = _weights()
        plan = tp.plan_block(world, rank, DIMS)
        x2, logits = _block(w, x, plan, rank, tp.all_reduce_sum, tp.gather_columns)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), x2=x2.numpy(), logits=logits.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_block_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    w, x = _weights()
    plans = [tp.plan_block(2, r, DIMS) for r in range(2)]

    # the exchange done by hand in one process: o and x2 partials summed, logits concatenated
    def lin(name, a, sh, r):
        bq, s, b = tp.shard_w4(torch.from_numpy(w[name][0]), torch.from_numpy(w[name][1]),
                               None if w[name][2] is None else torch.from_numpy(w[name][2]), sh, r)
        return orc.qmatmul_int4(a, bq.numpy(), s.numpy(), None if b is None else b.numpy(), "float16")
    H, I = DIMS.hidden_size, DIMS.inner_hidden_size
    o = sum(lin("o", lin("qkv", x, plans[r].qkv, r)[:, :plans[r].o.k_in(H)], plans[r].o, r) for r in range(2))
    o = orc.round_to(o, "float16")
    x2 = orc.round_to(sum(lin("w_out", lin("w_in", o, plans[r].w_in, r)[:, :plans[r].w_in.n_out(2 * I) // 2],
                              plans[r].w_out, r) for r in range(2)), "float16")
    logits = np.concatenate([lin("lm_head", x2, plans[r].lm_head, r) for r in range(2)], axis=-1)
    # ... and against the UNSHARDED linears where the math is shard-independent: o_proj over all heads
    full_o = orc.qmatmul_int4(orc.qmatmul_int4(x, *w["qkv"], "float16")[:, :H], w["o"][0], w["o"][1], None, "float16")
    assert_parity(o, full_o, "row-parallel o_proj == unsharded o_proj")
    full_logits = orc.qmatmul_int4(x2, w["lm_head"][0], w["lm_head"][1], None, "float16")
    assert np.array_equal(logits, full_logits), "column-parallel lm_head must equal the unsharded columns exactly"
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert_parity(got["x2"], x2, f"rank {r}: all-reduced w_out output")
        assert got["logits"].shape == (3, DIMS.vocab_size)
        assert_parity(got["logits"], logits, f"rank {r}: gathered logits")
