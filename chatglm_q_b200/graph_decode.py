"""CUDA-graph decode step behind `ChatGLM2Model.forward`'s call signature.

`ChatGLMDecoder.generate` (chatglm_q/decoder.py:76-97) calls `model(input_ids=..., past_key_values=...)`
once per token; the unmodified model then issues ~1 100 eager kernels and grows the KV cache with
`torch.cat` (chatglm_q/model.py:151-155), so a decode step costs ~17 ms of host time however fast the
dequant-matmuls are.  `GraphDecodeModel` wraps the SAME model object and makes the per-token call a
single CUDA-graph replay WITHOUT touching the reference code:

  * the KV cache is a fixed window of `max_len - 1` slots, valid tokens right-aligned, the unused
    slots on the left masked out through the model's own `attention_mask` argument
    (model.py:293-304: pad slots get -1e10 before the fp32 softmax, i.e. exactly zero weight);
    position ids follow from the mask's cumsum (model.py:296-297), so they stay correct;
  * one step = the reference forward on static shapes: `cat([cache, new])` has `max_len` rows, the new
    cache is rows 1..max_len of it (the oldest pad slot falls off), copied back into the static buffers;
  * the mask shifts left by one and the step is captured once, replayed per token; the only per-token
    host work is the 8-byte token-id copy and the sampling the decoder does itself.

The 113 dequant-matmuls inside the graph are this repo's kernels (chatglm_q_b200.install), launched with
programmatic dependent launch edges.  Prefill (more than one new token, or no cache) runs the model as is.
"""
from __future__ import annotations

import torch
from torch import Tensor


class _GraphCache:
    """Opaque `past_key_values` handle returned to the decoder (the state lives in the wrapper)."""

    def __init__(self, owner: "GraphDecodeModel"):
        self.owner = owner
        self.epoch = owner._epoch


class GraphDecodeModel:
    def __init__(self, model: torch.nn.Module, max_len: int = 1024):
        self.model = model
        self.max_len = int(max_len)
        self.graph: torch.cuda.CUDAGraph | None = None
        self.n_valid = 0
        self.batch = 1             # rows of the static window (a batch is captured too, on CUDA)
        self._eager_kv = None      # set when the window is exhausted (or a CPU batch): plain reference path
        self._epoch = 0

    # nn.Module-ish surface the decoder / loader touch
    def __getattr__(self, name):
        return getattr(self.model, name)

    def to(self, *a, **k):
        self.model.to(*a, **k)
        return self

    # ---------------------------------------------------------------- the captured step
    def _step(self):
        _, logits, kv = self.model(input_ids=self.ids, attention_mask=self.mask, past_key_values=self.kv)
        for (ks, vs), (k, v) in zip(self.kv, kv):
            ks.copy_(k[:, 1:])
            vs.copy_(v[:, 1:])
        self.mask.copy_(torch.cat([self.mask[:, 1:], self.mask[:, -1:]], dim=1))   # one more valid slot
        self.logits.copy_(logits)

    def _capture(self):
        dev = self.ids.device
        side = torch.cuda.Stream(device=dev)
        snap = ([(k.clone(), v.clone()) for k, v in self.kv], self.mask.clone())
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                       # warm-up: allocator pools, tensor maps, cuBLAS handles
                self._step()
            side.synchronize()
            self._restore(snap)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self._restore(snap)                          # the capture pass does not execute, but stay safe

    def _restore(self, snap):
        for (k, v), (k0, v0) in zip(self.kv, snap[0]):
            k.copy_(k0)
            v.copy_(v0)
        self.mask.copy_(snap[1])

    # ---------------------------------------------------------------- model(...) as the decoder calls it
    @torch.no_grad()
    def __call__(self, input_ids: Tensor = None, past_key_values=None, **kwargs):
        if kwargs or input_ids is None:
            return self.model(input_ids=input_ids, past_key_values=past_key_values, **kwargs)
        fresh = (past_key_values is None or not isinstance(past_key_values, _GraphCache)
                 or past_key_values.owner is not self or past_key_values.epoch != self._epoch)
        if fresh or input_ids.shape[1] != 1 or input_ids.shape[0] != self.batch:
            # prefill through the unmodified model, then move its cache into the static window
            if fresh:
                self._epoch += 1
            loss, logits, kv = self.model(input_ids=input_ids,
                                          past_key_values=None if fresh else self._export_kv())
            self._import_kv(kv, input_ids.device, logits)
            return loss, logits, _GraphCache(self)
        if self._eager_kv is not None or self.n_valid + 1 > self.max_len - 1:
            # window exhausted: continue on the reference path (correct, not graph-accelerated)
            if self._eager_kv is None:
                self._eager_kv = self._export_kv()
            loss, logits, self._eager_kv = self.model(input_ids=input_ids, past_key_values=self._eager_kv)
            self.n_valid = self._eager_kv[0][0].shape[1]
            return loss, logits, past_key_values
        self.ids.copy_(input_ids, non_blocking=True)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self.n_valid += 1
        return None, self.logits.clone(), past_key_values      # a fresh tensor per step, like the reference

    # ---------------------------------------------------------------- cache import / export
    def _import_kv(self, kv, device, logits):
        n = kv[0][0].shape[1]
        L = self.max_len - 1
        self._eager_kv = None
        B = kv[0][0].shape[0]
        if n > L or (B != 1 and not kv[0][0].is_cuda):   # longer than the window (or a batch without a GPU to capture on)
            self._eager_kv = kv
            self.n_valid = n
            self.batch = 1
            return
        first = not hasattr(self, "kv") or self.kv[0][0].device != kv[0][0].device \
            or self.kv[0][0].dtype != kv[0][0].dtype or self.kv[0][0].shape[0] != B
        if first:
            self.kv = tuple((torch.zeros((B, L, *k.shape[2:]), device=k.device, dtype=k.dtype),
                             torch.zeros((B, L, *v.shape[2:]), device=v.device, dtype=v.dtype)) for k, v in kv)
            self.mask = torch.zeros((B, self.max_len), dtype=torch.long, device=device)
            self.ids = torch.zeros((B, 1), dtype=torch.long, device=device)
            self.logits = torch.zeros((B, 1, logits.shape[-1]), device=device, dtype=logits.dtype)
            self.graph = None
        self.batch = B
        for (ks, vs), (k, v) in zip(self.kv, kv):
            ks.zero_()
            vs.zero_()
            ks[:, L - n:] = k
            vs[:, L - n:] = v
        self.mask.zero_()
        self.mask[:, self.max_len - (n + 1):] = 1      # n cached tokens + the slot of the next new token
        self.n_valid = n

    def _export_kv(self):
        if self._eager_kv is not None:
            return self._eager_kv
        n, L = self.n_valid, self.max_len - 1
        return tuple((k[:, L - n:].clone(), v[:, L - n:].clone()) for k, v in self.kv)
