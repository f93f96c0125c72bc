"""Fused batch-1 decode step behind `ChatGLM2Model.forward`'s call signature.

`ChatGLMDecoder.generate` (chatglm_q/decoder.py:76-97) calls `model(input_ids=..., past_key_values=...)`
once per token.  `FusedDecodeModel` wraps the SAME int4 model object (its module buffers are used in
place, nothing is re-packed or copied) and answers a one-token call with ONE CUDA-graph replay of
5 x n_layers + 2 launches of this repo's kernels through the C-ABI (include/cgq.h):

    cgq_decode_begin_w4                       word_embedding row of the token, position bookkeeping
    per layer (chatglm_q/model.py:230-246):
      cgq_w4a16_gemv_fused  RMSNORM, bias     attn_ln + qkv_proj            (:231, :139)
      cgq_decode_attention                    RoPE, KV append, attention    (:140-174)
      cgq_w4a16_gemv_fused  resid             o_proj, x = x + h             (:175, :243)
      cgq_w4a16_gemv_fused  RMSNORM           ffn_ln + w_in                 (:244, :200)
      cgq_w4a16_gemv_fused  SILU_GATE, resid  silu(h)*gate, w_out, x + h    (:201, :246)
    cgq_w4a16_gemv_fused  RMSNORM             final_ln + lm_head            (:381-382)

All launches are chained with programmatic dependent launch: each dequant-matmul streams its weights
from HBM while the kernel before it is still running.  The KV cache is a static `[max_len, groups, d]`
buffer per layer in the reference's own `past_key_values` layout; the position lives on the device so
the graph is static.  Prefill (more than one token, or no cache) runs the unmodified model (whose
linears are this repo's tcgen05 kernels after `install()`); when the static window is exhausted the
wrapper falls back to the unmodified model with an exported cache (correct, just not fused).

torch is used for device memory and the graph capture only.  There is no CPU / eager fallback for
the fused step itself: a model this path cannot take (int8, fp32, other head sizes) raises TypeError
at construction — use `GraphDecodeModel` for those.
"""
from __future__ import annotations

import os

import torch
from torch import Tensor

from . import _lib
from ._lib import PRO_NONE, PRO_RMSNORM, PRO_SILU_GATE

_DTYPE_CODE = {torch.float16: 0, torch.bfloat16: 1}


class _FusedCache:
    """Opaque `past_key_values` handle returned to the decoder (the state lives in the wrapper).  Stamped with the
    wrapper's epoch: a handle from an earlier `generate()` (another prefill has replaced the cache since) is treated
    like no cache at all instead of silently continuing on the newer session's rows."""

    def __init__(self, owner: "FusedDecodeModel"):
        self.owner = owner
        self.epoch = owner._epoch


def _is_w4_linear(m) -> bool:
    w, s = getattr(m, "weight", None), getattr(m, "weight_scale", None)
    return (isinstance(w, Tensor) and isinstance(s, Tensor) and w.dtype == torch.uint8 and w.dim() == 2
            and s.dim() == 2 and s.shape[1] == w.shape[1] and w.shape[0] * 2 == s.shape[0] * 32)


def _is_w8_linear(m) -> bool:
    w, s = getattr(m, "weight", None), getattr(m, "weight_scale", None)
    return (isinstance(w, Tensor) and isinstance(s, Tensor) and w.dtype == torch.int8 and w.dim() == 2
            and s.dim() == 1 and s.shape[0] == w.shape[0])


class FusedDecodeModel:
    def __init__(self, model: torch.nn.Module, max_len: int = 1024, handover: bool | None = None,
                 speculate: bool = False, last_logits_only: bool = False, alias_logits: bool | None = None,
                 one_launch: bool | None = None):
        cfg = model.config
        # The fused step writes its logits into ONE static buffer.  The reference returns a fresh tensor per call, so
        # by default the wrapper hands out a copy (130 KB); `alias_logits=True` returns the static buffer itself (it is
        # overwritten by the next step -- fine for ChatGLMDecoder.generate, which samples at once).  The speculative
        # sampler recognises "the logits of the step in flight" by that buffer, so speculate=True implies aliasing.
        self.alias_logits = bool(speculate) if alias_logits is None else bool(alias_logits)
        if speculate and not self.alias_logits:
            raise ValueError("FusedDecodeModel(speculate=True) needs alias_logits=True")
        self._epoch = 0
        # one_launch=True (or CGQ_ONE_LAUNCH=1): the whole token as ONE persistent cooperative kernel (cgq_step_*,
        # csrc/decode_mk.cu) when the model fits it (fp16, every N a multiple of 32, K <= 13824).  Correct and
        # sanitizer-clean, but measured SLOWER than the PDL-chained 5 x layers + 2 launches (1.47 vs 1.19 ms per token on
        # B200: ~2 us per grid barrier + ~2.5 us of serial prologue per phase, profiles/r02_step_program_timeline.txt),
        # so it is opt-in.
        self.one_launch = (bool(int(os.environ.get("CGQ_ONE_LAUNCH", "0") or 0)) if one_launch is None
                           else bool(one_launch))
        self.one_launch_refused = ""
        self._step_handle = None
        # last_logits_only=True: a prefill (several tokens) computes lm_head for the LAST position only and returns
        # logits [1, 1, V] -- all ChatGLMDecoder.generate reads is logits[0, -1] (decoder.py:85); at 2 048 tokens that
        # is 1.09 TFLOP and a 266 MB logits tensor less.  Off by default: the reference returns
        # every position.
        self.last_logits_only = bool(last_logits_only)
        # speculate=True: `self.sampler()` (bound to the decoder's top_p_sampling) starts the NEXT step from the
        # device-resident token right after the sampling kernel, so the decoder's host round trip (.item(), Python,
        # the H2D copy of the token id) overlaps the step instead of idling the GPU.  Exact: the next call's token is
        # checked on the host and a mismatch takes the step back.  Needs CPU `input_ids` (ChatGLMDecoder(device=None)).
        self.speculate = bool(speculate)
        self._spec = False
        self._spec_token = None
        self._host_ids = False
        # EXPERIMENTAL (cgq_handover_next): tile-granular hand-over between consecutive
        # dequant-matmuls instead of griddepcontrol.wait; off unless asked for (argument or CGQ_HANDOVER=1)
        self.handover = bool(int(os.environ.get("CGQ_HANDOVER", "0") or 0)) if handover is None else bool(handover)
        # which hand-overs take the counters: 1 o_proj->w_in, 2 w_in->w_out, 4 w_out->next qkv / lm_head
        self.hand_mask = int(os.environ.get("CGQ_HAND_MASK", "7") or 7)
        self.model = model
        self.cfg = cfg
        self.max_len = int(min(max_len, cfg.max_sequence_length - 1))   # row max_len of freqs_cis_cache is read
        self.graph: torch.cuda.CUDAGraph | None = None
        self.n_valid = 0
        self._eager_kv = None
        self._ready = False
        self.hints = int(os.environ.get("CGQ_PF_MB", "0") or 0) > 0
        lins = [model.lm_head]
        for layer in model.layers:
            lins += [layer.attn.qkv_proj, layer.attn.o_proj, layer.ffn.w_in, layer.ffn.w_out]
        emb = model.word_embedding
        ew = getattr(emb, "weight", None)
        if all(_is_w4_linear(m) for m in lins):
            self.kind = "w4"           # int4g32: uint8 [K/2, N] weights, [K/32, N] scales (int4/qlinear.py:75-108)
            if not (isinstance(ew, Tensor) and ew.dtype == torch.uint8 and hasattr(emb, "weight_scale")):
                raise TypeError("FusedDecodeModel needs the int4 QEmbedding word_embedding with the int4g32 model")
        elif all(_is_w8_linear(m) for m in lins):
            self.kind = "w8"           # int8 per channel: int8 [N, K] weights, [N] scales (int8/qlinear.py:77-107)
            if not (isinstance(ew, Tensor) and ew.dtype == torch.int8 and hasattr(emb, "weight_scale")
                    and emb.weight_scale.dim() == 1):
                raise TypeError("FusedDecodeModel needs the int8 QEmbedding word_embedding with the int8 model")
        else:
            raise TypeError("FusedDecodeModel needs the int4g32 model (uint8 [K/2,N] weights, [K/32,N] scales) or the "
                            "int8 model (int8 [N,K] weights, [N] scales), not a mix")
        if self.kind != "w4":
            self.one_launch = False
        dt = model.lm_head.weight_scale.dtype
        if dt not in _DTYPE_CODE:
            raise TypeError(f"FusedDecodeModel computes in float16 / bfloat16, model is {dt}")
        if cfg.head_hidden_size not in (64, 128):
            raise TypeError("FusedDecodeModel attention kernel is built for head sizes 64 and 128")
        self.dtype = dt
        self.code = _DTYPE_CODE[dt]

    # nn.Module-ish surface the decoder / loader touch
    def __getattr__(self, name):
        return getattr(self.model, name)

    def to(self, *a, **k):
        self.model.to(*a, **k)
        self._ready = False
        return self

    # ---------------------------------------------------------------- static buffers
    def _setup(self, device: torch.device):
        cfg, dt = self.cfg, self.dtype
        H, DH, NH, NG = cfg.hidden_size, cfg.head_hidden_size, cfg.num_attention_heads, cfg.num_multi_query_groups
        z = lambda *shape, dtype=dt: torch.zeros(shape, device=device, dtype=dtype)  # noqa: E731
        self.ids = z(1, 1, dtype=torch.long)
        self.state = z(4, dtype=torch.int32)      # [0] cached tokens, [1] this step's value, [2] token counter
        self.x = z(H)
        self.qkv = z(DH * (NH + 2 * NG))
        self.ao = z(DH * NH)
        self.u = z(2 * cfg.inner_hidden_size)
        self.logits = z(1, 1, cfg.vocab_size)
        self.tok_dev = z(1, dtype=torch.long)
        on_gpu = torch.device(device).type == "cuda"
        self.tok_host = torch.zeros(1, dtype=torch.long).pin_memory() if on_gpu else torch.zeros(1, dtype=torch.long)
        self.tok_event = torch.cuda.Event() if on_gpu else None
        self._spec = False
        # hand-over counters: one 128-byte line per producing launch (o_proj, w_in, w_out of every layer)
        self.ctr = z(3 * cfg.num_layers, 32, dtype=torch.int32)
        # the reference's past_key_values layout (n_batch, n_past, n_groups, 1, d_head), model.py:347-349
        self.kv = tuple((z(1, self.max_len, NG, 1, DH), z(1, self.max_len, NG, 1, DH)) for _ in range(cfg.num_layers))
        self.freqs = self.model.freqs_cis_cache
        assert self.freqs.dtype == dt and self.freqs.is_contiguous() and self.freqs.shape[1] == DH
        self.device = device
        self.graph = None
        self._drop_step()                 # the step program holds the old buffers' addresses
        if dt != torch.float16:
            self.one_launch = False
        self._ready = True

    # ---------------------------------------------------------------- the step, as C-ABI calls
    def _gemv(self, lib, stream, lin, a, out, prologue=PRO_NONE, norm=None, resid=None, nxt=None,
              wait=None, signal=None):
        k2, n = lin.weight.shape
        if self.kind == "w4" and self.handover and (wait is not None or signal is not None):
            # wait = (counter row, producing linear): poll until all of ITS output tiles are announced
            _lib.check(lib.cgq_handover_next(
                None if wait is None else self.ctr[wait[0]].data_ptr(),
                0 if wait is None else lib.cgq_w4_gemv_tiles(wait[1].weight.shape[1]),
                None if signal is None else self.ctr[signal].data_ptr()))
        if self.kind == "w4" and nxt is not None and self.hints:   # experimental L2 prefetch of the NEXT linear's weights (CGQ_PF_MB)
            _lib.check(lib.cgq_prefetch_next_w4(nxt.weight.data_ptr(), nxt.weight_scale.data_ptr(),
                                                nxt.weight.shape[1], nxt.weight.shape[0] * 2))
        bias = lin.bias if getattr(lin, "bias", None) is not None else None
        if self.kind == "w8":
            n8, k8 = lin.weight.shape
            _lib.check(lib.cgq_w8a16_gemv_fused(
                a.data_ptr(), lin.weight.data_ptr(), lin.weight_scale.data_ptr(),
                None if bias is None else bias.data_ptr(), None if resid is None else resid.data_ptr(), out.data_ptr(),
                n8, k8, self.code, prologue, None if norm is None else norm.weight.data_ptr(),
                float(norm.eps) if norm is not None else 0.0, stream))
            return
        _lib.check(lib.cgq_w4a16_gemv_fused(
            a.data_ptr(), lin.weight.data_ptr(), lin.weight_scale.data_ptr(),
            None if bias is None else bias.data_ptr(), None if resid is None else resid.data_ptr(),
            out.data_ptr(), n, 2 * k2, 32, self.code, prologue,
            None if norm is None else norm.weight.data_ptr(), float(norm.eps) if norm is not None else 0.0,
            stream))

    def _build_step(self):
        """The whole token step as a `cgq_step_op` array (include/cgq.h) -> one persistent cooperative kernel."""
        import ctypes

        from ._lib import STEP_ATTENTION, STEP_EMBED, STEP_LINEAR, StepOp

        lib = _lib.load()
        cfg, m = self.cfg, self.model
        ops = []

        def lin(l, a, out, pro=PRO_NONE, norm=None, resid=None, epi=0):
            k2, n = l.weight.shape
            bias = getattr(l, "bias", None)
            ops.append(StepOp(kind=STEP_LINEAR, Wq=l.weight.data_ptr(), scale=l.weight_scale.data_ptr(),
                              bias=None if bias is None else bias.data_ptr(), A=a.data_ptr(), C=out.data_ptr(),
                              resid=None if resid is None else resid.data_ptr(),
                              norm_w=None if norm is None else norm.weight.data_ptr(), N=n, K=2 * k2, prologue=pro,
                              eps=float(norm.eps) if norm is not None else 0.0, epilogue=epi))

        emb = m.word_embedding
        ops.append(StepOp(kind=STEP_EMBED, Wq=emb.weight.data_ptr(), scale=emb.weight_scale.data_ptr(),
                          C=self.x.data_ptr(), N=emb.weight.shape[1], V=emb.weight.shape[0] * 2, ids=self.ids.data_ptr()))
        for layer, (kc, vc) in zip(m.layers, self.kv):
            lin(layer.attn.qkv_proj, self.x, self.qkv, PRO_RMSNORM, layer.attn_ln)
            ops.append(StepOp(kind=STEP_ATTENTION, A=self.qkv.data_ptr(), C=self.ao.data_ptr(), freqs=self.freqs.data_ptr(),
                              kcache=kc.data_ptr(), vcache=vc.data_ptr(), n_head=cfg.num_attention_heads,
                              n_groups=cfg.num_multi_query_groups, d_head=cfg.head_hidden_size, max_len=self.max_len))
            lin(layer.attn.o_proj, self.ao, self.x, resid=self.x)
            # silu(h) * gate is applied once, in w_in's epilogue (u holds inner_hidden_size values)
            lin(layer.ffn.w_in, self.x, self.u, PRO_RMSNORM, layer.ffn_ln, epi=_lib.EPI_SILU_PAIR)
            lin(layer.ffn.w_out, self.u, self.x, PRO_NONE, resid=self.x)
        lin(m.lm_head, self.x, self.logits, PRO_RMSNORM, m.final_ln)
        arr = (StepOp * len(ops))(*ops)
        handle = ctypes.c_uint64(0)
        with torch.cuda.device(self.device):
            rc = lib.cgq_step_create(arr, len(ops), self.code, self.state.data_ptr(), ctypes.byref(handle))
        if rc != 0:
            return None, (lib.cgq_last_error() or b"").decode()
        return handle.value, ""

    def _drop_step(self):
        handle = self.__dict__.get("_step_handle")       # (never through __getattr__: it forwards to the model)
        if handle is not None:
            try:
                _lib.load().cgq_step_destroy(handle)
            except Exception:  # noqa: BLE001 -- interpreter shutdown
                pass
            self.__dict__["_step_handle"] = None

    def __del__(self):
        self._drop_step()

    def _launch_step(self):
        lib = _lib.load()
        cfg, m = self.cfg, self.model
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if self.one_launch and self._step_handle is None:
            self._step_handle, why = self._build_step()
            if self._step_handle is None:       # a shape the step program does not take: the launch-per-op chain
                self.one_launch = False
                self.one_launch_refused = why
        if self.one_launch:
            _lib.check(lib.cgq_step_run(self._step_handle, stream))
            return
        emb = m.word_embedding
        if self.kind == "w8":
            _lib.check(lib.cgq_decode_begin_w8(
                self.ids.data_ptr(), emb.weight.data_ptr(), emb.weight_scale.data_ptr(), self.x.data_ptr(),
                emb.weight.shape[0], emb.weight.shape[1], self.code, self.state.data_ptr(), stream))
        else:
            _lib.check(lib.cgq_decode_begin_w4(
                self.ids.data_ptr(), emb.weight.data_ptr(), emb.weight_scale.data_ptr(), self.x.data_ptr(),
                emb.weight.shape[0] * 2, emb.weight.shape[1], 32, self.code, self.state.data_ptr(), stream))
        firsts = [layer.attn.qkv_proj for layer in m.layers][1:] + [m.lm_head]
        if self.handover:
            self.ctr.zero_()
        prev_out = None          # (counter row, linear) of the previous layer's w_out
        for i, (layer, (kc, vc), nxt) in enumerate(zip(m.layers, self.kv, firsts)):
            # hand-over schedule: o_proj -> w_in -> w_out -> next qkv / lm_head by tile counters; qkv -> attention
            # -> o_proj stay on griddepcontrol.wait (hazards on x / qkv / ao / u: DESIGN.md §6.1a)
            self._gemv(lib, stream, layer.attn.qkv_proj, self.x, self.qkv, PRO_RMSNORM, layer.attn_ln,
                       nxt=layer.attn.o_proj, wait=prev_out if self.hand_mask & 4 else None)
            if i + 1 < len(self.kv):     # the next layer's cache rows are requested into L2 one layer ahead
                lib.cgq_attention_next_kv(self.kv[i + 1][0].data_ptr(), self.kv[i + 1][1].data_ptr())
            _lib.check(lib.cgq_decode_attention(
                self.qkv.data_ptr(), self.freqs.data_ptr(), kc.data_ptr(), vc.data_ptr(), self.ao.data_ptr(),
                self.state.data_ptr(), cfg.num_attention_heads, cfg.num_multi_query_groups,
                cfg.head_hidden_size, self.max_len, self.code, stream))
            hm = self.hand_mask
            self._gemv(lib, stream, layer.attn.o_proj, self.ao, self.x, resid=self.x, nxt=layer.ffn.w_in,
                       signal=3 * i if hm & 1 else None)
            self._gemv(lib, stream, layer.ffn.w_in, self.x, self.u, PRO_RMSNORM, layer.ffn_ln, nxt=layer.ffn.w_out,
                       wait=(3 * i, layer.attn.o_proj) if hm & 1 else None, signal=3 * i + 1 if hm & 2 else None)
            self._gemv(lib, stream, layer.ffn.w_out, self.u, self.x, PRO_SILU_GATE, resid=self.x, nxt=nxt,
                       wait=(3 * i + 1, layer.ffn.w_in) if hm & 2 else None, signal=3 * i + 2 if hm & 4 else None)
            prev_out = (3 * i + 2, layer.ffn.w_out)
        self._gemv(lib, stream, m.lm_head, self.x, self.logits, PRO_RMSNORM, m.final_ln,
                   nxt=m.layers[0].attn.qkv_proj,     # the next token's first linear
                   wait=prev_out if self.hand_mask & 4 else None)

    def launches_per_step(self) -> int:
        return 1 if (self.one_launch and self._step_handle is not None) else 5 * self.cfg.num_layers + 2

    def _capture(self):
        dev = self.device
        side = torch.cuda.Stream(device=dev)
        state0 = self.state.clone()
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.cuda.device(dev):
            self._launch_step()                       # warm-up: tensor maps, function attributes
            side.synchronize()
            self.state[:2].copy_(state0[:2])          # the warm-up step consumed a position; take it back (the token
                                                      # counter state[2] keeps counting: exchange epochs never repeat)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                self._launch_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.state[:2].copy_(state0[:2])

    # ---------------------------------------------------------------- model(...) as the decoder calls it
    @torch.no_grad()
    def __call__(self, input_ids: Tensor = None, past_key_values=None, **kwargs):
        if kwargs or input_ids is None:
            return self.model(input_ids=input_ids, past_key_values=past_key_values, **kwargs)
        on_host = input_ids.device.type == "cpu"
        dev = self.model.lm_head.weight.device
        fresh = (past_key_values is None or not isinstance(past_key_values, _FusedCache)
                 or past_key_values.owner is not self or past_key_values.epoch != self._epoch)
        if fresh or input_ids.shape[1] != 1 or input_ids.shape[0] != 1:
            # prefill, several new tokens or a batch: the unmodified model, on the cache wherever it lives
            self._drop_speculation()
            if fresh:
                self._epoch += 1
            with self._last_position_head():
                loss, logits, kv = self.model(input_ids=input_ids.to(dev),
                                              past_key_values=None if fresh else self._export_kv())
            self._import_kv(kv, dev)
            return loss, logits, _FusedCache(self)
        if self._spec:
            # a step for the token the sampler produced is already in flight (self.sampler())
            self._spec = False
            if on_host and int(input_ids[0, 0]) == self._spec_token:
                self.n_valid += 1
                self._host_ids = True
                return None, self._out_logits(), past_key_values
            self._drop_speculation(rewind=True)      # another token: take the step back, run the right one
        if self._eager_kv is not None or self.n_valid + 1 > self.max_len:
            # static window exhausted (or a cache the static buffers cannot hold): the unmodified model from here on
            if self._eager_kv is None:
                self._eager_kv = self._export_kv()
            loss, logits, self._eager_kv = self.model(input_ids=input_ids.to(dev), past_key_values=self._eager_kv)
            self.n_valid = self._eager_kv[0][0].shape[1]
            return loss, logits, past_key_values
        self._host_ids = on_host
        self.ids.copy_(input_ids, non_blocking=True)
        if self.graph is None:
            self._capture()
        self.graph.replay()
        self.n_valid += 1
        return None, self._out_logits(), past_key_values

    def _out_logits(self):
        return self.logits if self.alias_logits else self.logits.clone()

    def _last_position_head(self):
        """Context manager: while active, `model.lm_head` only sees the last position of its input (the unmodified
        ChatGLM2Model.forward applies it to every position, model.py:381-382)."""
        import contextlib

        if not self.last_logits_only or not isinstance(self.model, torch.nn.Module):
            return contextlib.nullcontext()
        model, head = self.model, self.model.lm_head

        class _LastRow(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.inner = head

            def forward(self, x):
                return self.inner(x[:, -1:, :])

        @contextlib.contextmanager
        def swap():
            model.lm_head = _LastRow()
            try:
                yield
            finally:
                model.lm_head = head

        return swap()

    # ---------------------------------------------------------------- sampler bound to this model (speculation)
    def _drop_speculation(self, rewind: bool = False):
        """Forget an in-flight speculative step.  Its KV row (slot n_valid) is overwritten by whatever runs next;
        the device-side position is put back when the fused step continues from here (`rewind`) — a prefill
        re-imports the cache and sets it anyway."""
        if self._spec or rewind:
            self._spec = False
            if rewind:
                self.state[:2].copy_(torch.tensor([self.n_valid, self.n_valid], dtype=torch.int32), non_blocking=False)

    def _can_speculate(self, logits: Tensor) -> bool:
        return (self.speculate and self._ready and self.graph is not None and self._host_ids
                and self._eager_kv is None and logits.dim() == 1 and logits.data_ptr() == self.logits.data_ptr()
                and self.n_valid + 1 <= self.max_len)

    def sampler(self):
        """A `top_p_sampling(logits, top_k, top_p, temperature)` (chatglm_q/decoder.py:12-27) bound to this model, for
        `install(package, sampler=model.sampler())`.  On the logits of a fused step it runs cgq_top_p_sample with the
        token left on the device, copies it to pinned host memory, starts the next step from the device copy and only
        then waits for the host copy: the GPU does not idle across the decoder's host round trip.  Returns a CPU
        0-dim int64 tensor in that case (the decoder calls `.item()` on it); anything else goes to ops.top_p_sampling.
        Same Exp(1) draw as ops.top_p_sampling / torch.multinomial: the same seed gives the same tokens."""
        from . import ops

        def top_p_sampling(logits: Tensor, top_k=100, top_p=0.8, temperature=1.0):
            if self._spec and logits.data_ptr() == self.logits.data_ptr():
                raise RuntimeError("FusedDecodeModel.sampler(): the logits of this step were already sampled and the "
                                   "next step has been started from that token (one sample per step)")
            if not self._can_speculate(logits):
                return ops.top_p_sampling(logits, top_k, top_p, temperature)
            assert temperature > 0 and top_p >= 0 and top_k >= 1
            lib = _lib.load()
            V = logits.shape[-1]
            k = min(int(top_k), V)
            with torch.cuda.device(self.device):
                q = torch.empty((1, k), dtype=torch.float32, device=self.device).exponential_(1)
                stream = torch.cuda.current_stream().cuda_stream
                _lib.check(lib.cgq_top_p_sample(logits.data_ptr(), V, self.code, int(top_k), float(top_p),
                                                float(temperature), q.data_ptr(), self.tok_dev.data_ptr(), None, None,
                                                stream))
                self.tok_host.copy_(self.tok_dev, non_blocking=True)
                self.tok_event.record()
                self.ids.copy_(self.tok_dev.view(1, 1), non_blocking=True)
                self.graph.replay()
                self._spec = True
                self.tok_event.synchronize()
            self._spec_token = int(self.tok_host[0])
            return torch.tensor(self._spec_token, dtype=torch.long)

        return top_p_sampling

    # ---------------------------------------------------------------- cache import / export
    def _import_kv(self, kv, device):
        n = kv[0][0].shape[1]
        self._eager_kv = None
        self.n_valid = n
        if n + 1 > self.max_len or kv[0][0].shape[0] != 1:
            self._eager_kv = kv
            return
        if not self._ready or self.device != device:
            self._setup(device)
        for (ks, vs), (k, v) in zip(self.kv, kv):
            ks[:, :n].copy_(k)
            vs[:, :n].copy_(v)
        self.state[:2].copy_(torch.tensor([n, n], dtype=torch.int32), non_blocking=False)

    def _export_kv(self):
        if self._eager_kv is not None:      # the cache lives outside the static buffers (batch > 1, window exhausted)
            return self._eager_kv
        n = self.n_valid
        return tuple((k[:, :n].clone(), v[:, :n].clone()) for k, v in self.kv)


def accelerate(model: torch.nn.Module, max_len: int = 1024):
    """The fastest decode wrapper this repo has for `model`: the fused step for the int4g32 model,
    the CUDA-graphed reference forward (graph_decode.GraphDecodeModel) for everything else."""
    try:
        return FusedDecodeModel(model, max_len=max_len)
    except TypeError:
        from .graph_decode import GraphDecodeModel

        return GraphDecodeModel(model, max_len=max_len)
