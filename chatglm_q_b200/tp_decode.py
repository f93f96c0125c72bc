"""Tensor-parallel fused decode step: one process per GPU, the batch-1 token step of the int4g32 model sharded over
`world` GPUs of one node (the reference itself is single-GPU, its oracle here is the one-GPU result).

    column-parallel   qkv_proj (this rank's heads + their KV group), w_in (same slice of h and gate), lm_head
    row-parallel      o_proj, w_out: fp32 partial sums exchanged INSIDE the decode kernel's epilogue -- 8-byte
                      {value, epoch} words stored straight into every peer's receive buffer over NVLink, summed in
                      rank order (include/cgq.h: cgq_tp_ctx, cgq_tp_next).  No NCCL call on the token path; every rank
                      holds the bit-identical hidden state.
    lm_head           every rank stores its vocabulary slice into every rank's logits row (peer stores), then one
                      cross-GPU barrier kernel: all ranks sample from identical logits, so `ChatGLMDecoder.generate`
                      runs unmodified on every rank (same seed => same tokens, no token broadcast).

`TPFusedDecodeModel(model, group=...)` wraps the SAME full model object on every rank (prefill runs on it,
replicated); the decode step reads per-rank contiguous shards cut from its module buffers (tp.shard_w4: slices of
the packed tensors, no re-quantisation).  A deployment that cannot hold the full model per GPU would load the shards
directly (tp.shard_w4 on the checkpoint tensors) -- the step only needs the shards.
"""
from __future__ import annotations

import torch
from torch import Tensor

from . import _lib, tp
from ._lib import PRO_NONE, PRO_RMSNORM, PRO_SILU_GATE
from .fused_decode import FusedDecodeModel, _FusedCache


class _Lin:
    """One rank's shard of an int4g32 linear (contiguous packed weight / scales / bias)."""

    def __init__(self, weight: Tensor, scale: Tensor, bias: Tensor | None):
        self.weight, self.weight_scale, self.bias = weight, scale, bias


class TPFusedDecodeModel(FusedDecodeModel):
    def __init__(self, model: torch.nn.Module, max_len: int = 1024, group=None, **kw):
        import torch.distributed as dist

        kw.pop("one_launch", None)
        super().__init__(model, max_len=max_len, one_launch=False, **kw)
        assert dist.is_available() and dist.is_initialized(), "TPFusedDecodeModel needs an initialised process group"
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        cfg = self.cfg
        self.dims = tp.ModelDims(cfg.hidden_size, cfg.inner_hidden_size, cfg.head_hidden_size,
                                 cfg.num_multi_query_groups, cfg.num_attention_heads, cfg.vocab_size)
        self.plan = tp.plan_block(self.world, self.rank, self.dims)
        if self.world > 1 and self.plan.kv_groups[1] - self.plan.kv_groups[0] != 1:
            raise TypeError("tensor-parallel decode expects every rank's heads to sit in one KV group")
        if self.dtype != torch.float16 and self.world > 1:
            raise TypeError("tensor-parallel decode is built for float16")
        self.ex = None
        self._shards = None

    # ---------------------------------------------------------------- shards + static buffers
    def _shard(self, lin, sh: tp.Shard) -> _Lin:
        bias = getattr(lin, "bias", None)
        w, s, b = tp.shard_w4(lin.weight, lin.weight_scale, bias, sh, rank=0)     # (bias kept: added once, after the sum)
        return _Lin(w, s, b)

    def _setup(self, device: torch.device):
        old = self.__dict__.get("state")
        counter = int(old[2]) if (old is not None and old.numel() >= 3) else 0
        super()._setup(device)
        if self.world == 1:
            return
        cfg, dt, pl = self.cfg, self.dtype, self.plan
        H, DH = cfg.hidden_size, cfg.head_hidden_size
        z = lambda *shape, dtype=dt: torch.zeros(shape, device=device, dtype=dtype)  # noqa: E731
        nh = pl.heads[1] - pl.heads[0]
        self.state = z(4, dtype=torch.int32)              # [0] cached tokens, [1] this step's value, [2] token counter
        self.state[2] = counter                           # exchange epochs never repeat, whatever is re-allocated
        self.qkv = z(pl.qkv.n_out(DH * (cfg.num_attention_heads + 2 * cfg.num_multi_query_groups)))
        self.ao = z(nh * DH)
        self.u = z(pl.w_in.n_out(2 * cfg.inner_hidden_size))
        # this rank's KV group only: [1, max_len, 1, 1, DH] per layer
        self.kv = tuple((z(1, self.max_len, 1, 1, DH), z(1, self.max_len, 1, 1, DH)) for _ in range(cfg.num_layers))
        if self._shards is None:
            m = self.model
            self._shards = [dict(qkv=self._shard(l.attn.qkv_proj, pl.qkv), o=self._shard(l.attn.o_proj, pl.o),
                                 w_in=self._shard(l.ffn.w_in, pl.w_in), w_out=self._shard(l.ffn.w_out, pl.w_out))
                            for l in m.layers]
            self._head = self._shard(m.lm_head, pl.lm_head)
        if self.ex is None:
            self.ex = tp.TpExchange(H, cfg.vocab_size, self.state[2:], self.group)
        else:
            self.ex.step = self.state[2:]
        self.logits = self.ex.logits.view(1, 1, cfg.vocab_size)    # peer-visible: every rank's lm_head slice lands here
        self.n_local_heads = nh

    # ---------------------------------------------------------------- the step, as C-ABI calls
    def _launch_step(self):
        if self.world == 1:
            return super()._launch_step()
        lib = _lib.load()
        cfg, m, pl, ex = self.cfg, self.model, self.plan, self.ex
        stream = torch.cuda.current_stream(self.device).cuda_stream
        emb = m.word_embedding
        _lib.check(lib.cgq_decode_begin_w4(
            self.ids.data_ptr(), emb.weight.data_ptr(), emb.weight_scale.data_ptr(), self.x.data_ptr(),
            emb.weight.shape[0] * 2, emb.weight.shape[1], 32, self.code, self.state.data_ptr(), stream))
        idx = 0
        for li, (layer, sh, (kc, vc)) in enumerate(zip(m.layers, self._shards, self.kv)):
            self._gemv(lib, stream, sh["qkv"], self.x, self.qkv, PRO_RMSNORM, layer.attn_ln)
            if li + 1 < len(self.kv):
                lib.cgq_attention_next_kv(self.kv[li + 1][0].data_ptr(), self.kv[li + 1][1].data_ptr())
            _lib.check(lib.cgq_decode_attention(
                self.qkv.data_ptr(), self.freqs.data_ptr(), kc.data_ptr(), vc.data_ptr(), self.ao.data_ptr(),
                self.state.data_ptr(), self.n_local_heads, 1, cfg.head_hidden_size, self.max_len, self.code, stream))
            ex.next_reduce(idx)
            self._gemv(lib, stream, sh["o"], self.ao, self.x, resid=self.x)
            self._gemv(lib, stream, sh["w_in"], self.x, self.u, PRO_RMSNORM, layer.ffn_ln)
            ex.next_reduce(idx + 1)
            self._gemv(lib, stream, sh["w_out"], self.u, self.x, PRO_SILU_GATE, resid=self.x)
            idx += 2
        assert idx < 127, "too many exchanges per token for the 7-bit exchange index"
        v0 = pl.lm_head.cols[0][0]
        ex.next_broadcast(v0)
        self._gemv(lib, stream, self._head, self.x, self.logits, PRO_RMSNORM, m.final_ln)
        ex.barrier(stream)

    def launches_per_step(self) -> int:
        return super().launches_per_step() + (1 if self.world > 1 else 0)

    # ---------------------------------------------------------------- cache import / export (this rank's KV group)
    def _import_kv(self, kv, device):
        if self.world == 1:
            return super()._import_kv(kv, device)
        n = kv[0][0].shape[1]
        self._eager_kv = None
        self.n_valid = n
        if n + 1 > self.max_len or kv[0][0].shape[0] != 1:
            raise RuntimeError(f"tensor-parallel decode: {n} cached tokens (batch {kv[0][0].shape[0]}) do not fit the static "
                               f"window of {self.max_len}; construct TPFusedDecodeModel with a larger max_len")
        if not self._ready or self.device != device:
            self._setup(device)
        g0 = self.plan.kv_groups[0]
        for (ks, vs), (k, v) in zip(self.kv, kv):
            ks[:, :n].copy_(k[:, :, g0:g0 + 1])
            vs[:, :n].copy_(v[:, :, g0:g0 + 1])
        st = torch.tensor([n, n], dtype=torch.int32)
        self.state[:2].copy_(st, non_blocking=False)

    def _export_kv(self):
        if self.world == 1:
            return super()._export_kv()
        raise RuntimeError("tensor-parallel decode keeps only this rank's KV group: continue with single-token calls, "
                           "or start a new prefill")

    @torch.no_grad()
    def __call__(self, input_ids: Tensor = None, past_key_values=None, **kwargs):
        if self.world > 1 and past_key_values is not None and isinstance(past_key_values, _FusedCache) \
                and input_ids is not None and input_ids.shape[1] == 1 and self.n_valid + 1 > self.max_len:
            raise RuntimeError(f"tensor-parallel decode: static window of {self.max_len} tokens exhausted")
        return super().__call__(input_ids=input_ids, past_key_values=past_key_values, **kwargs)
