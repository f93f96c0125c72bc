"""Tensor-parallel shard plan for the quantised linears of a ChatGLM2 block.

The reference has no distributed code at all; this module is the new multi-GPU
surface of the path and its oracle is the single-GPU result.  The split is Megatron-style and is
done ON THE PACKED TENSORS (no re-quantisation, no re-packing):

  column-parallel  qkv_proj, w_in, lm_head : slice columns of weight / weight_scale / bias, no exchange
  row-parallel     o_proj, w_out           : slice k-rows on a 32-row group boundary (so packed
                                             byte pairs and scale rows split cleanly), partial
                                             products are summed by ONE all-reduce per linear

Exact reference math needs TWO all-reduces per transformer block (after o_proj and after w_out:
`ffn_ln` consumes the completed residual, chatglm_q/model.py:243-245).

  qkv (model.py:139-146): columns = Q (32 heads x 128) | K (2 groups x 128) | V (2 groups x 128);
      head h attends with KV group h // 16.  Rank r owns heads [r*32/T, (r+1)*32/T), all of which
      sit in ONE KV group for T in {1 (both groups), 2, 4, 8}; that group's K/V columns are
      replicated on the T/2 ranks that need them.
  w_in (model.py:200): columns = [h | gate]; a rank takes the SAME slice of both halves so that
      silu(h) * gate stays rank-local.  inner = 13696 = 428 groups of 32, which 8 does not divide:
      the first (428 % T) ranks take one extra group (54/53 groups at T=8).
  w_out: k-rows = the rank's w_in slice.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
from torch import Tensor

GROUP = 32


@dataclass(frozen=True)
class ModelDims:
    """The handful of ChatGLM2Config fields (chatglm_q/model.py:9-22) the plan depends on."""
    hidden_size: int = 4096
    inner_hidden_size: int = 13696
    head_hidden_size: int = 128
    num_multi_query_groups: int = 2
    num_attention_heads: int = 32
    vocab_size: int = 65024


@dataclass(frozen=True)
class Shard:
    """One rank's part of one linear.  `cols`: half-open column ranges to concatenate (None = all);
    `krows`: half-open k range (None = all); `reduce`: partial sums need an all-reduce."""
    cols: tuple[tuple[int, int], ...] | None = None
    krows: tuple[int, int] | None = None
    reduce: bool = False

    def n_out(self, n_full: int) -> int:
        return n_full if self.cols is None else sum(b - a for a, b in self.cols)

    def k_in(self, k_full: int) -> int:
        return k_full if self.krows is None else self.krows[1] - self.krows[0]


def split_groups(n_groups: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, as-even-as-possible split of `n_groups` quantisation groups over `world` ranks
    (first n_groups % world ranks get one more).  Returns half-open GROUP ranges."""
    base, extra = divmod(n_groups, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


@dataclass(frozen=True)
class BlockPlan:
    world: int
    rank: int
    dims: ModelDims
    qkv: Shard = field(default=Shard())
    o: Shard = field(default=Shard())
    w_in: Shard = field(default=Shard())
    w_out: Shard = field(default=Shard())
    lm_head: Shard = field(default=Shard())
    heads: tuple[int, int] = (0, 32)       # this rank's attention heads
    kv_groups: tuple[int, int] = (0, 2)    # this rank's KV groups

    @property
    def allreduces_per_block(self) -> int:
        return 0 if self.world == 1 else 2


def plan_block(world: int, rank: int, dims: ModelDims = ModelDims()) -> BlockPlan:
    d, nh, ng = dims.head_hidden_size, dims.num_attention_heads, dims.num_multi_query_groups
    H, I, V = dims.hidden_size, dims.inner_hidden_size, dims.vocab_size
    assert 0 <= rank < world
    if world == 1:
        return BlockPlan(1, 0, dims, heads=(0, nh), kv_groups=(0, ng))
    assert nh % world == 0, f"{nh} heads do not split over {world} ranks"
    hpr = nh // world                       # heads per rank
    hpg = nh // ng                          # heads per KV group
    assert hpg % hpr == 0 or hpr % hpg == 0, "a rank's heads must align with KV groups"
    h0, h1 = rank * hpr, (rank + 1) * hpr
    g0, g1 = h0 // hpg, (h1 - 1) // hpg + 1  # groups touched (one group when hpr <= hpg)
    q_cols = (h0 * d, h1 * d)
    k_cols = (nh * d + g0 * d, nh * d + g1 * d)
    v_cols = (nh * d + ng * d + g0 * d, nh * d + ng * d + g1 * d)
    assert I % GROUP == 0 and (nh * d) % (GROUP * world) == 0
    gi0, gi1 = split_groups(I // GROUP, world)[rank]
    i0, i1 = gi0 * GROUP, gi1 * GROUP
    assert V % world == 0
    v0, v1 = rank * (V // world), (rank + 1) * (V // world)
    return BlockPlan(
        world, rank, dims,
        qkv=Shard(cols=(q_cols, k_cols, v_cols)),
        o=Shard(krows=(h0 * d, h1 * d), reduce=True),
        w_in=Shard(cols=((i0, i1), (I + i0, I + i1))),
        w_out=Shard(krows=(i0, i1), reduce=True),
        lm_head=Shard(cols=((v0, v1),)),
        heads=(h0, h1), kv_groups=(g0, g1))


def _cat_cols(t: Tensor, cols, dim: int) -> Tensor:
    if cols is None:
        return t
    return torch.cat([t.narrow(dim, a, b - a) for a, b in cols], dim=dim).contiguous()


def shard_w4(weight: Tensor, scale: Tensor, bias: Tensor | None, sh: Shard, rank: int = 0):
    """Slice an int4g32 linear's buffers (weight u8 [K/2, N], scale [K/32, N], bias [N] —
    chatglm_q/int4/qlinear.py:83-88).  The bias of a row-parallel linear stays on rank 0 only."""
    if sh.krows is not None:
        k0, k1 = sh.krows
        assert k0 % GROUP == 0 and k1 % GROUP == 0, "row split must sit on a group boundary"
        weight = weight[k0 // 2:k1 // 2]
        scale = scale[k0 // GROUP:k1 // GROUP]
        if bias is not None and rank != 0:
            bias = None
    weight, scale = _cat_cols(weight, sh.cols, 1), _cat_cols(scale, sh.cols, 1)
    if bias is not None:
        bias = _cat_cols(bias, sh.cols, 0)
    return weight.contiguous(), scale.contiguous(), bias


def shard_w8(weight: Tensor, scale: Tensor, bias: Tensor | None, sh: Shard, rank: int = 0):
    """Slice an int8 linear's buffers (weight i8 [N, K], scale [N], bias [N] —
    chatglm_q/int8/qlinear.py:82-87)."""
    if sh.krows is not None:
        k0, k1 = sh.krows
        weight = weight[:, k0:k1]
        if bias is not None and rank != 0:
            bias = None
    weight, scale = _cat_cols(weight, sh.cols, 0), _cat_cols(scale, sh.cols, 0)
    if bias is not None:
        bias = _cat_cols(bias, sh.cols, 0)
    return weight.contiguous(), scale.contiguous(), bias


def shard_input(x: Tensor, sh: Shard) -> Tensor:
    """The activation columns a row-parallel linear consumes on this rank."""
    if sh.krows is None:
        return x
    return x[..., sh.krows[0]:sh.krows[1]]


def all_reduce_sum(partial: Tensor, group=None) -> Tensor:
    """The one exchange step of a row-parallel linear: sum of the ranks' partial products
    (NCCL over NVLink/NVSwitch on the GPUs, gloo in the CPU tests).  In place, returns `partial`."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def gather_columns(part: Tensor, sh: Shard, n_full: int, group=None) -> Tensor:
    """Re-assemble a column-parallel output (lm_head logits) on every rank."""
    import torch.distributed as dist

    if sh.cols is None or not (dist.is_available() and dist.is_initialized()):
        return part
    world = dist.get_world_size(group)
    if world == 1:
        return part
    parts = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(parts, part.contiguous(), group=group)
    return torch.cat(parts, dim=-1)


# ---------------------------------------------------------------------- peer-visible device memory (NVLink P2P)
class PeerMemory:
    """`nbytes` of zeroed device memory on every rank of `group`, each rank's block mapped into every other rank's
    address space (cgq_ipc_*: cudaMalloc + legacy CUDA IPC handles, exchanged here over torch.distributed).
    `ptrs[r]` is rank r's block as seen from THIS process (ptrs[rank] = the local allocation): the decode kernels
    store their partial sums / logits straight into peers' blocks over NVLink (include/cgq.h, cgq_tp_ctx).
    torch is plumbing only: the handle exchange and a tensor view of the local block."""

    def __init__(self, nbytes: int, group=None):
        import ctypes

        import torch.distributed as dist

        from . import _lib

        lib = _lib.load()
        self._lib, self.group = lib, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.nbytes = int(nbytes)
        self.device = torch.device("cuda", torch.cuda.current_device())
        local = ctypes.c_void_p(0)
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.cgq_ipc_alloc(self.nbytes, ctypes.byref(local), handle))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(local.value)
                continue
            p = ctypes.c_void_p(0)
            _lib.check(lib.cgq_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        self._local = local.value
        dist.barrier(group=group)          # every rank has mapped every block before anyone stores into one

    def tensor(self, dtype: torch.dtype, numel: int, byte_offset: int = 0) -> Tensor:
        """A torch view of the LOCAL block (through __cuda_array_interface__; the memory stays owned by this object)."""
        typestr = {torch.float16: "<f2", torch.int32: "<i4", torch.uint8: "|u1", torch.float32: "<f4",
                   torch.int64: "<i8"}[dtype]
        assert byte_offset + numel * torch.empty((), dtype=dtype).element_size() <= self.nbytes

        class _Raw:
            __cuda_array_interface__ = {"shape": (numel,), "typestr": typestr, "version": 2,
                                        "data": (self._local + byte_offset, False)}

        t = torch.as_tensor(_Raw(), device=self.device)
        t._cgq_owner = self                # keep the allocation alive as long as the view is
        return t

    def close(self):
        if getattr(self, "_local", None):
            for p in self._opened:
                self._lib.cgq_ipc_close(p)
            self._lib.cgq_ipc_free(self._local)
            self._local, self._opened = None, []


class TpExchange:
    """The device-side plumbing of one rank's tensor-parallel decode step: LL receive buffers for the row-parallel
    linears, a peer-visible logits row for the broadcast lm_head, barrier flags, the error word.  Hands out the
    `cgq_tp_ctx` hints for `cgq_tp_next` (include/cgq.h)."""

    def __init__(self, hidden: int, vocab: int, step_counter: Tensor, group=None):
        from . import _lib

        self.max_n = int(hidden)
        # layout of every rank's block: [LL words: 2 slots x world x hidden x 8 B][logits: vocab x 2 B][flags: 8 x 4 B][err 4 B]
        import torch.distributed as dist

        world = dist.get_world_size(group)
        self.ll_bytes = 2 * world * self.max_n * 8
        self.logit_off = self.ll_bytes
        self.flag_off = (self.logit_off + vocab * 2 + 255) // 256 * 256
        self.err_off = self.flag_off + 64
        self.mem = PeerMemory(self.err_off + 64, group)
        self.world, self.rank = self.mem.world, self.mem.rank
        self.step = step_counter                      # int32 device tensor, element 0 = token counter
        self.logits = self.mem.tensor(torch.float16, vocab, self.logit_off)
        self.err = self.mem.tensor(torch.int32, 1, self.err_off)
        self._lib, self._TpCtx = _lib, _lib.TpCtx

    def _ctx(self, reduce: bool, bcast_offset: int | None):
        c = self._TpCtx()
        c.world, c.rank, c.max_n = self.world, self.rank, self.max_n
        c.out_offset = 0 if bcast_offset is None else int(bcast_offset)
        for r in range(self.world):
            c.recv[r] = self.mem.ptrs[r] if reduce else None
            c.out[r] = (self.mem.ptrs[r] + self.logit_off) if bcast_offset is not None else None
        c.step = self.step.data_ptr()
        c.err = self.mem.ptrs[self.rank] + self.err_off
        return c

    def next_reduce(self, idx: int) -> None:
        """The NEXT cgq_w4a16_gemv_fused launch of this thread is a row-parallel linear: exchange + sum its partials."""
        import ctypes

        self._lib.check(self._lib.load().cgq_tp_next(ctypes.byref(self._ctx(True, None)), idx))

    def next_broadcast(self, column_offset: int) -> None:
        """The NEXT launch stores its N columns into every rank's logits row at `column_offset`."""
        import ctypes

        self._lib.check(self._lib.load().cgq_tp_next(ctypes.byref(self._ctx(False, column_offset)), 0))

    def barrier(self, stream: int) -> None:
        import ctypes

        arr = (ctypes.c_void_p * 8)(*[(self.mem.ptrs[r] + self.flag_off) for r in range(self.world)], *([None] * (8 - self.world)))
        self._lib.check(self._lib.load().cgq_tp_barrier(arr, self.world, self.rank, self.step.data_ptr(),
                                                       self.mem.ptrs[self.rank] + self.err_off, stream))

    def error(self) -> int:
        """0, or the epoch of an exchange whose peer words never arrived (synchronises)."""
        return int(self.err.item())
