#!/usr/bin/env python
"""bench.py — ChatGLM2-6B int4g32 batch-1 decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's own CPU path

A "step" is ONE decode token's pass over the hot path: the 113 int4g32 dequant-matmuls of
ChatGLM2-6B at M=1 (28 x {qkv_proj 4096->4608 +bias, o_proj 4096->4096, w_in 4096->27392,
w_out 13696->4096} + lm_head 4096->65024), chained through the C-ABI (`cgq_w4a16_gemm`), weights
resident in HBM, captured once in a CUDA graph (the 113 launches are launch-bound from Python) and
replayed.  One step streams 3.36 GB of distinct weights (27x the 126 MB L2), so every launch reads
its weights from HBM; no explicit L2 flush is needed and none is done.

  value      tokens/s of that device-resident step (whole job; N>1 = tensor-parallel, see below)
  roofline   the dominant (only) kernel, w4_gemv_kernel: algorithmic bytes of the step / step time
             (= algorithmic bytes per launch / average launch duration, CUDA events on the launch
             stream) against MEASURED_PEAKS.json's HBM copy bandwidth
  e2e        the same metric through the reference-facing API: the UNMODIFIED reference
             `ChatGLMDecoder.generate` (baseline/_ref) on a random-init ChatGLM2-6B int4g32 model
             with this repo's kernels installed behind its QLinear modules; every step copies the
             token id host->device and reads the sampled token back (`.item()`), tok/s computed
             exactly as the reference's `gen` figure (chatglm_q/decoder.py:99-105)
  cpu_baseline  the reference's torch CPU path of the same step on this box's host cores
  microbench    BASELINE.json configs[1] shapes (seq x 4096) x (4096 x N), per-shape GB/s / TFLOP/s

N > 1 (torchrun, one rank per GPU): the token step is tensor-parallel (chatglm_q_b200/tp.py):
column-split qkv / w_in / lm_head, row-split o_proj / w_out with an NCCL all-reduce after each
(2 per block), logits all-gathered.  Total work is fixed => "scaling": "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ChatGLM2-6B int4g32 decode tok/s (bs=1)"
UNIT = "tok/s"
H, INNER, VOCAB, LAYERS = 4096, 13696, 65024, 28
QKV_N = 4096 + 2 * 256


# ----------------------------------------------------------------------------- shapes / bytes
def w4_bytes(m: int, k: int, n: int, bias: bool) -> int:
    """Algorithmic bytes of one int4g32 dequant-matmul (SURVEY §8d / DESIGN.md)."""
    return k * n // 2 + (k // 32) * n * 2 + m * k * 2 + m * n * 2 + (n * 2 if bias else 0)


def token_linears(world: int = 1, rank: int = 0):
    """(name, K, N, bias) of one rank's linears for one block, and its lm_head."""
    from chatglm_q_b200 import tp

    plan = tp.plan_block(world, rank)
    block = [
        ("qkv_proj", H, plan.qkv.n_out(QKV_N), True),
        ("o_proj", plan.o.k_in(H), H, False),
        ("w_in", H, plan.w_in.n_out(2 * INNER), False),
        ("w_out", plan.w_out.k_in(INNER), H, False),
    ]
    head = ("lm_head", H, plan.lm_head.n_out(VOCAB), False)
    return plan, block, head


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


def csrc_sha16() -> str:
    """Hash of the CUDA sources the library is built from: ncu summaries under profiles/ are stamped with it, so a
    number measured on another build is never reported as this build's."""
    import hashlib

    h = hashlib.sha256()
    for f in sorted((ROOT / "chatglm_q_b200" / "csrc").glob("*.cu*")):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def ncu_traffic_per_launch():
    """(dram__bytes_read.sum + dram__bytes_write.sum per launch of the decode kernel, source) from the committed ncu
    launch list of this same step (profiles/r02_token_summary.json, scripts/gpu_profile2.sh).  Only a summary stamped
    with THIS build's source hash counts; anything else is reported as null with the reason."""
    p = ROOT / "profiles" / "r02_token_summary.json"
    try:
        d = json.loads(p.read_text())
        if d.get("csrc_sha16") != csrc_sha16():
            return None, f"profiles/{p.name} was taken on another build (csrc {d.get('csrc_sha16')} != {csrc_sha16()})"
        return round(float(d["dram_bytes"]) / int(d["launches"])), f"profiles/{p.name} (ncu launch list, same csrc hash)"
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return None, "no ncu launch list committed for this build"


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- synthetic weights
def make_w4(torch, k: int, n: int, bias: bool, device, gen, k_full: int = None):
    """Random packed nibbles in 1..15 (what quantize_int4 emits: zero-mean q-8 in -7..7, so a chain
    of linears neither drifts nor overflows) + group scales sized for unit gain.  A row-parallel shard
    (k rows of a k_full-row weight) keeps the full linear's scale, so that the all-reduced sum has unit gain."""
    lo = torch.randint(1, 16, (k // 2, n), dtype=torch.uint8, device=device, generator=gen)
    hi = torch.randint(1, 16, (k // 2, n), dtype=torch.uint8, device=device, generator=gen)
    w = lo | (hi << 4)
    del lo, hi
    s = (torch.rand((k // 32, n), device=device, generator=gen) * 0.5 + 0.75) * (1.0 / (4.4 * (k_full or k) ** 0.5))
    b = (torch.randn(n, device=device, generator=gen) * 0.02).half() if bias else None
    return w, s.half(), b


class TokenStep:
    """The 113 (per-rank) dequant-matmuls of one decode token, chained on device buffers."""

    def __init__(self, torch, device, world: int, rank: int, m: int = 1):
        from chatglm_q_b200 import ops

        self.torch, self.ops, self.world, self.m = torch, ops, world, m
        self.hints = int(os.environ.get("CGQ_PF_MB", "0") or 0) > 0    # experimental L2 prefetch hints: off
        self.plan, block, head = token_linears(world, rank)
        gen = torch.Generator(device=device).manual_seed(1234 + rank)
        k_full = {"o_proj": H, "w_out": INNER}          # row-parallel linears: the shard keeps the full-K scale
        self.layers = [[(name, *make_w4(torch, k, n, bias, device, gen, k_full.get(name))) for name, k, n, bias in block]
                       for _ in range(LAYERS)]
        self.head = make_w4(torch, head[1], head[2], False, device, gen)
        # the hidden state is replicated: the same on every rank
        self.x = torch.randn((m, H), device=device, generator=torch.Generator(device=device).manual_seed(99)).half()
        self.bytes = LAYERS * sum(w4_bytes(m, k, n, b) for _, k, n, b in block) + w4_bytes(m, head[1], head[2], False)
        self.launches = LAYERS * 4 + 1
        self.kq = block[1][1]      # o_proj K on this rank
        self.ki = block[3][1]      # w_out K on this rank
        self.logits = None

    def run(self):
        ops, x = self.ops, self.x
        dist = None
        if self.world > 1:
            import torch.distributed as dist
        hint = ops.prefetch_next_s4 if self.hints else (lambda *a: None)
        wl, sl, _ = self.head
        firsts = [(l[0][1], l[0][2]) for l in self.layers[1:]] + [(wl, sl)]
        for ((_, wq, sq, bq), (_, wo, so, _), (_, wi, si, _), (_, wu, su, _)), nxt in zip(self.layers, firsts):
            hint(wo, so)       # each launch also streams the NEXT linear's weights into L2 (cgq_prefetch_next_w4)
            qkv = ops.dynamic_quant_matmul_s4(x, wq, sq, bias=bq)
            hint(wi, si)
            o = ops.dynamic_quant_matmul_s4(qkv[:, :self.kq], wo, so)     # stand-in for attention out
            if dist is not None:
                dist.all_reduce(o)
            hint(wu, su)
            hin = ops.dynamic_quant_matmul_s4(o, wi, si)
            hint(*nxt)
            x = ops.dynamic_quant_matmul_s4(hin[:, :self.ki], wu, su)     # stand-in for silu(h)*gate
            if dist is not None:
                dist.all_reduce(x)
        hint(self.layers[0][0][1], self.layers[0][0][2])                  # the next token's first linear
        logits = ops.dynamic_quant_matmul_s4(x, wl, sl)
        if dist is not None:
            parts = [self.torch.empty_like(logits) for _ in range(self.world)]
            dist.all_gather(parts, logits)
            logits = parts[0]
        self.logits = logits
        return logits


class TPTokenStep:
    """N > 1: the same 113 dequant-matmuls, tensor-parallel (chatglm_q_b200/tp.py): every rank generates the SAME full
    weights (same seed) and cuts its shards from them -- column slices of qkv / w_in / lm_head, k-row slices of
    o_proj / w_out.  The row-parallel linears exchange their fp32 partial sums INSIDE the decode kernel's epilogue
    over NVLink peer memory (cgq_tp_next, include/cgq.h): no NCCL call on the token path.  lm_head stores its
    vocabulary slice into every rank's logits row, one cross-GPU barrier kernel ends the step.  The full weights stay
    on the rank for the parity gate (TP logits vs the unsharded chain on the same inputs)."""

    def __init__(self, torch, device, world: int, rank: int):
        from chatglm_q_b200 import ops, tp

        self.torch, self.ops, self.world, self.rank = torch, ops, world, rank
        self.plan, block, head = token_linears(world, rank)
        gen = torch.Generator(device=device).manual_seed(1234)          # identical on every rank
        full_block = [("qkv_proj", H, QKV_N, True), ("o_proj", H, H, False), ("w_in", H, 2 * INNER, False),
                      ("w_out", INNER, H, False)]
        shards = (self.plan.qkv, self.plan.o, self.plan.w_in, self.plan.w_out)
        self.full, self.layers = [], []
        for _ in range(LAYERS):
            fl, sl = [], []
            for (name, k, n, bias), sh in zip(full_block, shards):
                w, s, b = make_w4(torch, k, n, bias, device, gen)
                fl.append((w, s, b))
                sl.append(tp.shard_w4(w, s, b, sh, rank=0))
            self.full.append(fl)
            self.layers.append(sl)
        self.full_head = make_w4(torch, H, VOCAB, False, device, gen)
        self.head = tp.shard_w4(*self.full_head, self.plan.lm_head, rank=0)
        self.x0 = torch.randn((1, H), device=device, generator=torch.Generator(device=device).manual_seed(99)).half()
        self.bytes = LAYERS * sum(w4_bytes(1, k, n, b) for _, k, n, b in block) + w4_bytes(1, head[1], head[2], False)
        self.launches = LAYERS * 4 + 1
        self.kq, self.ki = block[1][1], block[3][1]          # o_proj / w_out K on this rank
        z = lambda n: torch.zeros(n, device=device, dtype=torch.float16)  # noqa: E731
        self.state = torch.zeros(4, dtype=torch.int32, device=device)     # [2] = token counter (exchange epochs)
        self.x, self.qkv, self.o, self.hin = z(H), z(block[0][2]), z(H), z(block[2][2])
        self.ex = tp.TpExchange(H, VOCAB, self.state[2:], None)
        self.logits = self.ex.logits
        self.v0 = self.plan.lm_head.cols[0][0]
        self.one = torch.ones(1, dtype=torch.int32, device=device)

    def run(self):
        ops, ex = self.ops, self.ex
        self.state[2:3].add_(self.one)                          # what cgq_decode_begin_w4 does in the fused step
        self.x.copy_(self.x0[0])
        idx = 0
        for (wq, sq, bq), (wo, so, _), (wi, si, _), (wu, su, _) in self.layers:
            ops.gemv_fused_s4(self.x, wq, sq, bias=bq, out=self.qkv)
            ex.next_reduce(idx)
            ops.gemv_fused_s4(self.qkv[:self.kq], wo, so, out=self.o)          # stand-in for attention out: this rank's q columns
            ops.gemv_fused_s4(self.o, wi, si, out=self.hin)
            ex.next_reduce(idx + 1)
            ops.gemv_fused_s4(self.hin[:self.ki], wu, su, out=self.x)          # stand-in for silu(h)*gate: this rank's h columns
            idx += 2
        ex.next_broadcast(self.v0)
        ops.gemv_fused_s4(self.x, *self.head[:2], out=self.logits)
        ex.barrier(self.torch.cuda.current_stream().cuda_stream)
        return self.logits

    def run_unsharded(self):
        """The same chain on the full weights with the single-GPU launches: the parity gate's reference."""
        ops = self.ops
        x = self.x0
        for (wq, sq, bq), (wo, so, _), (wi, si, _), (wu, su, _) in self.full:
            qkv = ops.dynamic_quant_matmul_s4(x, wq, sq, bias=bq)
            o = ops.dynamic_quant_matmul_s4(qkv[:, :H], wo, so)
            hin = ops.dynamic_quant_matmul_s4(o, wi, si)
            x = ops.dynamic_quant_matmul_s4(hin[:, :INNER], wu, su)
        return ops.dynamic_quant_matmul_s4(x, *self.full_head[:2])[0]

    def parity_gate(self, dist):
        """TP logits against the unsharded chain on every rank (bar: |d| <= 2e-2 |ref| + 2e-2 rms -- two
        implementations of a 28-block chain with different summation orders), identical rows on all ranks, no lost
        peer word."""
        torch = self.torch
        ref = self.run_unsharded().float()
        got = self.run().float().clone()
        torch.cuda.synchronize()
        rms = ref.pow(2).mean().sqrt()
        ratio = ((got - ref).abs() / (2e-2 * ref.abs() + 2e-2 * rms)).max()
        t = torch.stack([ratio, got.double().sum().float(), -got.double().sum().float()])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        err = self.ex.error()
        return {"worst_ratio": round(float(t[0]), 4), "bar": "2e-2 * (|ref| + rms(ref)), 28-block chain, max over ranks",
                "ok": bool(t[0] <= 1.0) and err == 0 and bool(torch.isfinite(got).all()),
                "rows_identical_on_all_ranks": bool(t[1] == -t[2]), "lost_peer_words": err,
                "reference": "unsharded chain of the same weights through the single-GPU launches, on every rank"}


# ----------------------------------------------------------------------------- microbench (configs[1])
def microbench(torch, device, peaks, seqs=(1, 128, 2048), ns=(4608, 13696, 27392), k=4096, iters=10):
    """BASELINE.json configs[1]: (seq x 4096) x (4096 x N) int4g32 dequant-matmul.  Per shape a CUDA graph
    of one launch per weight copy (>= 400 MB of distinct weights per rotation, 3x the 126 MB L2) is
    replayed `iters` times between CUDA events: per-launch device time without Python launch overhead."""
    from chatglm_q_b200 import ops

    out = []
    gen = torch.Generator(device=device).manual_seed(7)
    stream = torch.cuda.Stream(device=device)
    # GPU-vs-GPU baseline: the reference's own Triton kernel (chatglm_q/int4/triton_ops.py:90-139, unmodified, from
    # baseline/_ref) on the same inputs, timed the same way
    tri, tri_err = None, None
    try:
        if import_reference() is None:
            raise ImportError("baseline/_ref missing")
        from chatglm_q.int4.triton_ops import dynamic_quant_matmul_s4 as tri
    except Exception as e:  # noqa: BLE001 -- a Triton that cannot import is a finding to report, not a crash
        tri_err = f"{type(e).__name__}: {e}"[:300]

    def timed(fn, use):
        with torch.cuda.stream(stream), torch.no_grad():
            for i in range(use):
                fn(i)
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for i in range(use):
                    fn(i)
            for _ in range(3):
                g.replay()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(iters):
                g.replay()
            e1.record(stream)
            stream.synchronize()
        del g
        return e0.elapsed_time(e1) * 1e3 / (iters * use)

    for n in ns:
        per = k * n // 2 + (k // 32) * n * 2
        copies = max(2, -(-400_000_000 // per))
        ws = [make_w4(torch, k, n, False, device, gen) for _ in range(copies)]
        for m in seqs:
            use = copies if m <= 8 else min(copies, 4)   # M > 8 is tensor-bound: L2 residency is irrelevant
            a = torch.randn((m, k), device=device, generator=gen).half()
            us = timed(lambda i: ops.dynamic_quant_matmul_s4(a, ws[i][0], ws[i][1]), use)
            by, fl = w4_bytes(m, k, n, False), 2.0 * m * n * k
            row = {"M": m, "K": k, "N": n, "us": round(us, 2),
                   "GBps": round(by / us / 1e3, 1), "TFLOPs": round(fl / us / 1e6, 2),
                   "hbm_frac": round(by / us / 1e3 / peaks["hbm_gbs"], 3),
                   "tensor_frac": round(fl / us / 1e6 / peaks["bf16_tflops"], 4),
                   "kernel": ("w4_gemv_kernel (IMMA.16832 on base-256 digits of the activation, TMA ring, cluster DSMEM reduce)"
                              if m == 1 else "w4_gemv_kernel (mma.sync f16, TMA ring, cluster DSMEM reduce)") if m <= 8
                             else "wq_gemm_tc_kernel (tcgen05 + TMEM)"}
            if tri is not None:
                try:
                    tus = timed(lambda i: tri(a, ws[i][0], ws[i][1], allow_tf32=False), use)
                    row["triton_us"] = round(tus, 2)
                    row["speedup_vs_reference_triton"] = round(tus / us, 2)
                except Exception as e:  # noqa: BLE001
                    tri_err = f"{type(e).__name__}: {e}"[:300]
                    tri = None
            if tri is None:
                row["triton_us"] = None
                row["triton_error"] = tri_err
            out.append(row)
            del a
        del ws
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- reference import
def import_reference():
    """The unmodified reference package from baseline/_ref (pip --target install of /root/reference)."""
    ref = ROOT / "baseline" / "_ref"
    if not (ref / "chatglm_q").exists():
        return None
    if str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    import chatglm_q  # noqa: F401
    return chatglm_q


class StubTokenizer:
    """Duck-typed tokenizer (no sentencepiece.model exists offline); the eos id is unreachable so
    exactly max_generated_tokens are produced (chatglm_q/decoder.py:90-91)."""

    def __init__(self, prompt_len: int):
        self.prompt_len = prompt_len

    def __getitem__(self, key):
        return -1

    def encode(self, text):
        return [64790, 64792] + [1000 + 7 * i for i in range(self.prompt_len - 2)]

    def decode(self, ids):
        return "x" * len(ids)


def build_ref_int4_model(torch, device, seed=0):
    """Random-init ChatGLM2-6B int4g32 model built by the reference's own factory
    (chatglm_q/loader.py:53-66) directly on the GPU, filled with synthetic quantised weights."""
    from chatglm_q.loader import ChatGLMLoadConfig, create_quant_int4_model
    from chatglm_q.model import ChatGLM2Config
    from chatglm_q.int4.qlinear import DynamicQuantizeLinear, QEmbedding

    cfg = ChatGLM2Config()
    with torch.device(device):
        model = create_quant_int4_model(cfg, 32, torch.float16)
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                k, n = mod.in_features, mod.out_features
                w, s, b = make_w4(torch, k, n, mod.bias is not None, device, gen)
                mod.apply_weights_(w, s, b)
            elif isinstance(mod, QEmbedding):
                v, d = mod.num_embeddings, mod.embedding_dim
                mod.weight.copy_(torch.randint(0, 256, (v // 2, d), dtype=torch.uint8, device=device, generator=gen))
                mod.weight_scale.copy_((torch.rand((v // 32, d), device=device, generator=gen) * 0.25 + 0.05).half())
    model.eval()
    return ChatGLMLoadConfig(model_config=cfg, quant_type="int4g32", torch_dtype="float16"), model


def e2e_decode(torch, device, gen_tokens=128, prompt_len=32):
    """Reference ChatGLMDecoder.generate, unmodified, with this repo's kernels installed behind its QLinear
    modules.  Headline `value`: the model object wrapped in chatglm_q_b200.FusedDecodeModel (one CUDA-graph
    replay of the fused decode step per token); `graphed_reference_forward` is the unmodified forward under a
    CUDA graph, `eager` the plain unwrapped model."""
    if import_reference() is None:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "baseline/_ref missing: reference decoder not importable"}
    from chatglm_q.decoder import ChatGLMDecoder
    from chatglm_q_b200.graph_decode import GraphDecodeModel
    from chatglm_q_b200.install import install
    import chatglm_q.decoder as decmod

    cfg, model = build_ref_int4_model(torch, device)
    real_perf = time.perf_counter

    def run(decoder):
        """the reference times each step itself (decoder.py:80-87) but only prints the figure; record the
        same intervals by wrapping the clock it reads"""
        times = []

        class _Clock:
            @staticmethod
            def perf_counter():
                t = real_perf()
                times.append(t)
                return t

            def __getattr__(self, k):
                return getattr(time, k)

        torch.manual_seed(0)
        for _ in decoder.generate("warm-up", max_generated_tokens=8):   # allocator / tensor maps / graph capture
            pass
        torch.cuda.synchronize()
        decmod.time = _Clock()
        try:
            for _ in decoder.generate("bench", max_generated_tokens=gen_tokens):
                pass
        finally:
            decmod.time = time
        steps = [b - a for a, b in zip(times[0::2], times[1::2])]
        rest = steps[1:]
        return round(len(rest) / sum(rest), 2), round(steps[0], 4), len(steps)

    from chatglm_q_b200.fused_decode import FusedDecodeModel

    tok = StubTokenizer(prompt_len)
    # GPU-vs-GPU baseline first, before anything of this repo is bound: the UNMODIFIED reference model + decoder on
    # its own Triton kernels (chatglm_q/int4/triton_ops.py), eager -- what a user of the reference gets on this B200
    ref_triton = {"value": None}
    try:
        import chatglm_q.int4.qlinear as _q4

        assert _q4.KERNEL_IMPL == "triton", f"reference kernel impl is {_q4.KERNEL_IMPL!r}"
        v, pf, _ = run(ChatGLMDecoder(cfg, model, tok, device=device, time_log=False))
        ref_triton = {"value": v, "prefill_s": pf, "unit": UNIT,
                      "how": "unmodified reference model + decoder on the reference's own Triton kernels (no install()), "
                             "eager, same prompt / tokens / sampler"}
    except Exception as e:  # noqa: BLE001 -- Triton failing to compile the reference kernels is a finding
        ref_triton = {"value": None, "error": f"{type(e).__name__}: {e}"[:400]}
    install("chatglm_q")
    eager, eager_prefill, _ = run(ChatGLMDecoder(cfg, model, tok, device=device, time_log=False))
    graphed, graphed_prefill, _ = run(ChatGLMDecoder(cfg, GraphDecodeModel(model, max_len=prompt_len + gen_tokens + 32),
                                                     tok, device=device, time_log=False))
    # (alias_logits=True: ChatGLMDecoder.generate samples each step's logits at once, so the static logits row is
    # handed out as is -- the default hands out a copy per step, as the reference returns a fresh tensor)
    fused_model = FusedDecodeModel(model, max_len=prompt_len + gen_tokens + 32, alias_logits=True)
    fused_ref_sampler, _, _ = run(ChatGLMDecoder(cfg, fused_model, tok, device=device, time_log=False))
    # the sampler is a module global of the reference too (decoder.py:12, resolved at :85): rebind it to the
    # one-launch cgq_top_p_sample (same signature, same token for the same seed) -- the headline configuration
    from chatglm_q_b200.install import uninstall
    # headline: every step copies the token id host -> device and reads the sampled token back (8 bytes each way)
    install("chatglm_q", sampler=True)
    try:
        fused, prefill_s, n_tok = run(ChatGLMDecoder(cfg, fused_model, tok, device=device, time_log=False))
    finally:
        uninstall("chatglm_q")
        install("chatglm_q")
    # reported beside it: the sampler bound to the model starts the NEXT step from the device-resident token, so the
    # decoder's host round trip overlaps the step (exact: the id the decoder hands back is checked on the host, a
    # mismatch takes the step back).  The decoder gets device=None and passes CPU ids; on a hit nothing is copied H2D.
    del fused_model
    fused_model = FusedDecodeModel(model, max_len=prompt_len + gen_tokens + 32, speculate=True)
    install("chatglm_q", sampler=fused_model.sampler())
    try:
        fused_spec, _, _ = run(ChatGLMDecoder(cfg, fused_model, tok, device=None, time_log=False))
    finally:
        uninstall("chatglm_q")
        install("chatglm_q")
    # device time of the fused step alone (graph replays between CUDA events; the KV window is rewound so
    # every replay attends over the same context length as the middle of the generation)
    dev_us = None
    if fused_model.graph is not None:
        ctx = prompt_len + gen_tokens // 2
        st = torch.tensor([ctx, ctx], dtype=torch.int32, device=device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        for i in range(reps + 5):
            if i == 5:
                e0.record()
            fused_model.state[:2].copy_(st, non_blocking=True)
            fused_model.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        dev_us = round(e0.elapsed_time(e1) * 1e3 / reps, 1)
    launches = fused_model.launches_per_step()
    del model, fused_model
    torch.cuda.empty_cache()
    return {"value": fused, "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
            "how": f"reference ChatGLMDecoder.generate (unmodified) driving chatglm_q_b200.FusedDecodeModel wrapped around "
                   f"the unmodified int4g32 ChatGLM2Model (its module buffers in place): one CUDA-graph replay of {launches} "
                   f"C-ABI launches per token (fused RMSNorm/SiLU-gate/residual dequant-matmuls + RoPE/KV/attention kernel, "
                   f"PDL-chained); prompt {prompt_len} tok, {n_tok} tok generated, 'gen' tok/s = tokens after the first / "
                   f"their summed wall time (each step: H2D token id, graph replay, top-p sampling by cgq_top_p_sample "
                   f"bound to the decoder's top_p_sampling global -- 2 launches instead of the reference's ~15 torch kernels "
                   f"and multinomial's host sync --, .item() D2H)",
            "prefill_s": prefill_s, "tokens": n_tok,
            "fused_step_reference_sampler": {"value": fused_ref_sampler,
                                             "how": "same fused step, the reference's own torch top_p_sampling"},
            "fused_step_speculative_next_step": {"value": fused_spec, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8,
                                                 "how": "FusedDecodeModel(speculate=True).sampler(): the next step's graph "
                                                        "replay is issued from the device-resident token right after the "
                                                        "sampling kernel and verified against the CPU id the decoder "
                                                        "(device=None) hands back; the GPU does not idle across the host "
                                                        "round trip"},
            "fused_step_device_us": dev_us, "fused_step_launches": launches,
            "reference_triton": ref_triton,
            "graphed_reference_forward": {"value": graphed, "prefill_s": graphed_prefill,
                                          "how": "same decoder, unmodified model forward captured in one CUDA graph "
                                                 "(chatglm_q_b200.GraphDecodeModel), ~1100 kernels per token"},
            "eager": {"value": eager, "prefill_s": eager_prefill,
                      "how": "same decoder, model NOT wrapped: ~1100 eager kernels + torch.cat KV growth per token"}}


# ----------------------------------------------------------------------------- BASELINE config 4: the int8 model
def build_ref_int8_model(torch, device, seed=0):
    """Random-init ChatGLM2-6B int8 model built by the reference's own factory (chatglm_q/loader.py:41-50)."""
    from chatglm_q.loader import ChatGLMLoadConfig, create_quant_int8_model
    from chatglm_q.model import ChatGLM2Config
    from chatglm_q.int8.qlinear import DynamicQuantizeLinear, QEmbedding

    cfg = ChatGLM2Config()
    with torch.device(device):
        model = create_quant_int8_model(cfg, torch.float16)
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, DynamicQuantizeLinear):
                n, k = mod.out_features, mod.in_features
                wq = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=device, generator=gen)
                sc = (torch.rand(n, device=device, generator=gen) * 0.5 + 0.75) / (73.0 * k ** 0.5)
                b = (torch.randn(n, device=device, generator=gen) * 0.02).half() if mod.bias is not None else None
                mod.apply_weights_(wq, sc.half(), b)
            elif isinstance(mod, QEmbedding):
                mod.weight.copy_(torch.randint(-127, 128, mod.weight.shape, dtype=torch.int8, device=device, generator=gen))
                mod.weight_scale.copy_((torch.rand(mod.weight_scale.shape, device=device, generator=gen) * 0.01 + 0.005).half())
    model.eval()
    return ChatGLMLoadConfig(model_config=cfg, quant_type="int8", torch_dtype="float16"), model


def int8_config4(torch, device, peaks, gen_tokens=64, prompt_len=32):
    """BASELINE.json configs[3]: ChatGLM2-6B int8 -- decode bs=1 (device chain of the 113 linears + e2e through the
    unmodified decoder on the fused step), decode bs=8 (the 113 linears at M=8; `generate` is batch-1 only,
    chatglm_q/decoder.py:70), prefill of 2 048 tokens through the unmodified forward (tcgen05 kernels)."""
    from chatglm_q_b200 import ops

    out = {}
    gen = torch.Generator(device=device).manual_seed(11)
    shapes = [("qkv", H, QKV_N, True), ("o", H, H, False), ("w_in", H, 2 * INNER, False), ("w_out", INNER, H, False)]

    def w8(k, n, bias):
        w = torch.randint(-127, 128, (n, k), dtype=torch.int8, device=device, generator=gen)
        sc = ((torch.rand(n, device=device, generator=gen) * 0.5 + 0.75) / (73.0 * k ** 0.5)).half()
        b = (torch.randn(n, device=device, generator=gen) * 0.02).half() if bias else None
        return w, sc, b

    layers = [[w8(k, n, b) for _, k, n, b in shapes] for _ in range(LAYERS)]
    head = w8(H, VOCAB, False)
    nbytes = lambda m: LAYERS * sum(k * n + 2 * n + 2 * m * k + 2 * m * n + (2 * n if b else 0) for _, k, n, b in shapes) \
        + H * VOCAB + 2 * VOCAB + 2 * m * H + 2 * m * VOCAB  # noqa: E731
    stream = torch.cuda.Stream(device=device)
    for m in (1, 8):
        x0 = torch.randn((m, H), device=device, generator=gen).half()

        def chain():
            x = x0
            for (wq, sq, bq), (wo, so, _), (wi, si, _), (wu, su, _) in layers:
                qkv = ops.dynamic_quant_matmul(x, wq.t(), sq, bias=bq)
                o = ops.dynamic_quant_matmul(qkv[:, :H], wo.t(), so)
                hin = ops.dynamic_quant_matmul(o, wi.t(), si)
                x = ops.dynamic_quant_matmul(hin[:, :INNER], wu.t(), su)
            return ops.dynamic_quant_matmul(x, head[0].t(), head[1])

        with torch.cuda.stream(stream), torch.no_grad():
            chain()
            stream.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                lg = chain()
            for _ in range(3):
                g.replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(20):
                g.replay()
            e1.record(stream)
            stream.synchronize()
        ms = e0.elapsed_time(e1) / 20
        assert torch.isfinite(lg.float()).all()
        out[f"decode_bs{m}_linears"] = {"ms_per_step": round(ms, 4), "tok_s": round(m * 1e3 / ms, 1),
                                        "GBps": round(nbytes(m) / ms / 1e6, 1),
                                        "hbm_frac": round(nbytes(m) / ms / 1e6 / peaks["hbm_gbs"], 4),
                                        "how": f"113 int8 dequant-matmuls at M={m} (cgq_w8a16_gemm), one CUDA graph"}
        del g
    del layers, head
    torch.cuda.empty_cache()
    if import_reference() is None:
        return out
    from chatglm_q.decoder import ChatGLMDecoder
    import chatglm_q.decoder as decmod
    from chatglm_q_b200.fused_decode import FusedDecodeModel
    from chatglm_q_b200.install import install, uninstall

    install("chatglm_q", sampler=True)
    try:
        cfg, model = build_ref_int8_model(torch, device)
        fused = FusedDecodeModel(model, max_len=prompt_len + gen_tokens + 32, alias_logits=True)
        times, real_perf = [], time.perf_counter

        class _Clock:
            @staticmethod
            def perf_counter():
                t = real_perf()
                times.append(t)
                return t

            def __getattr__(self, k):
                return getattr(time, k)

        dec = ChatGLMDecoder(cfg, fused, StubTokenizer(prompt_len), device=device, time_log=False)
        torch.manual_seed(0)
        for _ in dec.generate("warm-up", max_generated_tokens=8):
            pass
        torch.cuda.synchronize()
        decmod.time = _Clock()
        try:
            for _ in dec.generate("bench", max_generated_tokens=gen_tokens):
                pass
        finally:
            decmod.time = time
        steps = [b - a for a, b in zip(times[0::2], times[1::2])]
        out["decode_bs1_e2e"] = {"tok_s": round(len(steps[1:]) / sum(steps[1:]), 2), "tokens": len(steps),
                                 "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                                 "how": "unmodified ChatGLMDecoder.generate on FusedDecodeModel(int8 model): one CUDA-graph "
                                        "replay of the fused step (cgq_w8a16_gemv_fused + attention) per token"}
        # prefill of 2 048 tokens through the unmodified forward (its int8 QLinear modules run the tcgen05 kernels)
        ids = torch.randint(1000, 60000, (1, 2048), generator=torch.Generator().manual_seed(0)).to(device)
        with torch.no_grad():
            fused(input_ids=ids, past_key_values=None)
            torch.cuda.synchronize()
            runs = []
            for _ in range(4):     # host-launch bound (eager reference glue): report the best and the spread
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fused(input_ids=ids, past_key_values=None)
                e1.record()
                torch.cuda.synchronize()
                runs.append(e0.elapsed_time(e1))
        out["prefill_2048"] = {"ms": round(min(runs), 2), "ms_runs": [round(r, 2) for r in runs],
                               "how": "2 048-token prompt through the unmodified reference forward, int8 tcgen05 kernels behind "
                                      "its QLinear modules (eager glue of the reference included); best of 4 runs"}
        # decode at bs = 8 end to end: `generate` is batch-1 only (decoder.py:70), so the reference's own
        # model.forward(input_ids [8, L], past_key_values) is driven greedily -- unmodified eager forward (its torch.cat
        # KV growth, its attention), the int8 linears on cgq_w8a16_gemm at M = 8 (w8_gemv_mx_kernel)
        from chatglm_q_b200.graph_decode import GraphDecodeModel
        bs, steps8 = 8, 24
        pin_in = torch.empty((bs, 1), dtype=torch.int64).pin_memory()
        pin_out = torch.empty((bs,), dtype=torch.int64).pin_memory()
        prompt8 = torch.randint(1000, 60000, (bs, prompt_len), generator=torch.Generator().manual_seed(1)).to(device)
        for name, m8, how in (
                ("decode_bs8_e2e", GraphDecodeModel(model, max_len=prompt_len + steps8 + 16),
                 "reference model.forward at batch 8 captured in ONE CUDA graph per step (chatglm_q_b200.GraphDecodeModel: "
                 "static 8-row KV window, the reference's own attention / norms inside the graph), int8 linears on "
                 "cgq_w8a16_gemm at M = 8 (w8_gemv_mx_kernel)"),
                ("decode_bs8_e2e_eager", model,
                 "unmodified reference model.forward at batch 8, eager (its torch.cat KV growth), int8 linears on "
                 "cgq_w8a16_gemm at M = 8")):
            with torch.no_grad():
                _, lg, kv = m8(input_ids=prompt8, past_key_values=None)
                tok = lg[:, -1].argmax(-1)
                lat = []
                for i in range(steps8 + 4):
                    t0 = time.perf_counter()
                    pin_in[:, 0] = pin_out if i else tok.cpu()
                    ids8 = pin_in.to(device, non_blocking=True)                  # H2D: this step's 8 token ids
                    _, lg, kv = m8(input_ids=ids8, past_key_values=kv)
                    pin_out.copy_(lg[:, -1].argmax(-1), non_blocking=False)      # D2H: the 8 sampled tokens (greedy)
                    lat.append(time.perf_counter() - t0)
            lat = lat[4:]
            out[name] = {"tok_s": round(bs * len(lat) / sum(lat), 1), "ms_per_step": round(1e3 * sum(lat) / len(lat), 3),
                         "h2d_bytes_per_step": 64, "d2h_bytes_per_step": 64, "steps": len(lat), "how": how}
        del model, fused
    finally:
        uninstall("chatglm_q")
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_reference_step(threads: int):
    """The step on host cores through the reference's own CPU path (`A.matmul(unpack_int4(B, scale))`,
    chatglm_q/int4/qlinear.py:47-50), fp16 like the GPU arm.  `step(blocks)` runs `blocks` transformer blocks' four
    linears (the 28 blocks of the model reuse ONE block's weights: 3.4 GB of packed weights x the fp16 unpack
    temporaries do not need 28 copies to be timed; the 57 MB of a block exceed the host caches) plus lm_head and
    returns (seconds for the blocks, seconds for lm_head)."""
    import torch

    torch.set_num_threads(threads)
    kind = "reference"
    if import_reference() is not None:
        from chatglm_q.int4.qlinear import DynamicQuantizeLinear

        def make(k, n, bias):
            g = torch.Generator().manual_seed(k + n)
            lin = DynamicQuantizeLinear(k, n, bias=bias, dtype=torch.float16)
            w, s, b = make_w4(torch, k, n, bias, "cpu", g)
            lin.apply_weights_(w, s, b)
            return lin
    else:   # reference not installed: time the oracle's C port of the same algorithm
        kind = "port"
        import numpy as np
        from oracle import c_oracle

        c_oracle.set_threads(threads)

        def make(k, n, bias):
            rng = np.random.default_rng(k + n)
            wq = rng.integers(0, 256, size=(k // 2, n), dtype=np.uint8)
            s = (rng.random((k // 32, n), dtype=np.float32) * 0.01).astype(np.float16).astype(np.float32)
            return lambda x: torch.from_numpy(c_oracle.w4a16_gemm(x.float().numpy(), wq, s, None, "float16"))

    _, block, head = token_linears(1, 0)
    mods = [make(k, n, b) for _, k, n, b in block]
    lm = make(head[1], head[2], False)
    x = torch.randn(1, H).half()

    def step(blocks: int = LAYERS):
        with torch.no_grad():
            t0 = time.perf_counter()
            y = x
            for _ in range(blocks):
                qkv = mods[0](y)
                o = mods[1](qkv[:, :H].contiguous())
                hin = mods[2](o)
                y = mods[3](hin[:, :INNER].contiguous())
                y = y / y.float().abs().max().clamp_min(1e-3).half()      # keep the synthetic chain finite
            t1 = time.perf_counter()
            lm(y)
            t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    return step, kind


WORKLOAD = ("ChatGLM2-6B int4g32 bs=1 decode token: 113 QLinear calls at M=1 "
            "(28 x qkv/o/w_in/w_out + lm_head), random-init packed weights")


def bench_config(world: int, graph: bool = True) -> dict:
    """The workload description BOTH arms print (the driver compares them)."""
    return {"workload": WORKLOAD,
            "parallelism": "single GPU" if world == 1 else f"tp{world} (column/row split, 2 reductions per block)",
            "l2": "inputs larger than L2: 3.36 GB of distinct weights per step vs 126 MB L2"}


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores: every step is ONE FULL decode
    token (28 blocks + lm_head), so `ms_per_step` is the measured wall time of a step and value = 1000 / ms_per_step
    -- nothing is extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    step, kind = cpu_reference_step(threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step(2)                                   # warm-up: two blocks + lm_head (allocator, thread pool)
    walls, blk, head = [], 0.0, 0.0
    for _ in range(args.steps):
        tb, th = step(LAYERS)
        walls.append(tb + th)
        blk += tb
        head += th
    per_tok = sum(walls) / len(walls)
    value = 1.0 / per_tok
    sample = (f"every step = one full token: 28 x (qkv,o,w_in,w_out) + lm_head at M=1 fp16 through the reference CPU path "
              f"(chatglm_q/int4/qlinear.py:47-50); {args.steps} steps, {sum(walls):.1f} s of CPU work "
              f"(blocks {blk:.1f} s, lm_head {head:.1f} s)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_tok * 1e3, 2),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic", "config": bench_config(args.gpus),
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- own arm
def run_own_arm(args):
    import torch

    if os.environ.get("CGQ_BENCH_DEBUG"):       # where is every rank after N seconds? (hang triage)
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["CGQ_BENCH_DEBUG"]), exit=True)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists for this path)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=device)
    peaks = load_peaks()

    step = TokenStep(torch, device, world, rank) if world == 1 else TPTokenStep(torch, device, world, rank)
    stream = torch.cuda.Stream(device=device)
    tp_parity = None
    with torch.cuda.stream(stream), torch.no_grad():
        if world > 1:               # parity gate BEFORE anything is timed
            tp_parity = step.parity_gate(dist)
        for _ in range(2):          # tensor-map cache, workspace
            step.run()
        stream.synchronize()
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                step.run()
        run = graph.replay if graph is not None else step.run
        for _ in range(max(3, args.warmup)):
            run()
        stream.synchronize()

        def barrier():
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.3)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            run()
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = ms / args.steps
    value = 1e3 / ms_per_step
    assert torch.isfinite(step.logits.float()).all(), "non-finite logits from the token step"
    if world > 1:
        assert step.ex.error() == 0, f"a peer's exchange word never arrived (epoch {step.ex.error()})"

    total_bytes = step.bytes   # this rank's algorithmic bytes per step (all ranks stream concurrently)
    achieved = total_bytes / (ms_per_step * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "w4_gemv_kernel", "achieved": round(achieved, 1),
                "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved / peaks["hbm_gbs"], 4),
                "frac_of_8TBps_nominal": round(achieved / 8000.0, 4), "peak_source": peaks["source"],
                "traffic": ncu_traffic_per_launch()[0] if world == 1 else None,
                "traffic_source": ncu_traffic_per_launch()[1] if world == 1 else "N>1: not captured",
                "algorithmic_bytes_per_launch": round(total_bytes / step.launches),
                "algorithmic_bytes_per_step": total_bytes, "launches_per_step": step.launches,
                "avg_launch_us": round(ms_per_step * 1e3 / step.launches, 3),
                "note": "per rank; includes the inter-kernel gaps of the graph-replayed step (at N>1: and the in-kernel "
                        "NVLink exchange of the row-parallel linears + the end-of-step barrier kernel)"}
    del graph, step
    torch.cuda.empty_cache()

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": bench_config(world),
            "launch": "no CUDA graph" if args.no_graph else "CUDA graph replay of the C-ABI launches",
            "tp_parity": tp_parity,
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(roofline["launches_per_step"] * args.steps),
        }
    if world == 1:
        line["e2e"] = (e2e_decode(torch, device, args.gen_tokens) if not args.no_e2e else
                       {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "--no-e2e"})
        if not args.no_micro:
            line["microbench"] = microbench(torch, device, peaks)
        if not args.no_int8:
            line["int8"] = int8_config4(torch, device, peaks)
        if not args.no_cpu:
            threads = os.cpu_count() or 1
            cstep, kind = cpu_reference_step(threads)
            cstep(1)
            nb, tb, th, reps = 4, 0.0, 0.0, 0
            while tb + th < 12.0 and reps < 20:      # bounded sample: 4 of the 28 blocks + lm_head per repetition
                b, h_ = cstep(nb)
                tb += b
                th += h_
                reps += 1
            per = LAYERS * (tb / (reps * nb)) + th / reps
            line["cpu_baseline"] = {"value": round(1.0 / per, 4), "unit": UNIT, "cores": threads, "kind": kind,
                                    "measured_block_ms": round(tb / (reps * nb) * 1e3, 2),
                                    "measured_lm_head_ms": round(th / reps * 1e3, 2), "layers": LAYERS,
                                    "sample": f"{reps} x ({nb} of 28 blocks + lm_head) at M=1 fp16 through the reference "
                                              f"CPU path; token time DERIVED = 28 x measured block + lm_head; "
                                              f"{tb + th:.1f} s of CPU work (`--impl reference` times full tokens)"}
    else:
        # e2e at N>1: the same TP step fed from pinned host memory and read back every step.  A watchdog makes
        # sure a stuck collective can never hang the bench: the line is then printed without the e2e number.
        def bail():
            if rank == 0:
                line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": H * 2, "d2h_bytes_per_step": 8,
                               "note": "TP e2e leg did not finish within 420 s and was abandoned"}
                print(json.dumps(line), flush=True)
            os._exit(0)

        dog = threading.Timer(420.0, bail)
        dog.daemon = True
        dog.start()
        line_e2e = tp_e2e(torch, device, world, rank, args)
        dog.cancel()
        if rank == 0:
            line["e2e"] = line_e2e
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        # NCCL's communicator teardown can block forever after CUDA graphs that captured collectives
        # (seen on 2 x B200: both ranks stuck in destroy_process_group after the line was printed): leave
        # through a barrier and a hard exit instead -- the result is already out.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def tp_e2e(torch, device, world, rank, args):
    """N>1 end to end: the UNMODIFIED reference `ChatGLMDecoder.generate` on EVERY rank, driving
    chatglm_q_b200.TPFusedDecodeModel (the fused decode step sharded over the ranks, in-kernel NVLink exchange); every
    rank samples from the identical all-gathered logits with the same seed, so no token is ever broadcast.  Each step
    copies the token id host->device and reads the sampled token back.  Before timing: greedy parity gate of the TP
    step against the single-GPU fused step on the same model (logits within the 2e-2 chain bar, same tokens)."""
    import torch.distributed as dist

    if import_reference() is None:
        return {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "baseline/_ref missing: reference decoder not importable"}
    from chatglm_q.decoder import ChatGLMDecoder
    import chatglm_q.decoder as decmod
    from chatglm_q_b200.fused_decode import FusedDecodeModel
    from chatglm_q_b200.install import install, uninstall
    from chatglm_q_b200.tp_decode import TPFusedDecodeModel

    prompt_len, gen_tokens = 32, args.gen_tokens
    install("chatglm_q", sampler=True)
    try:
        cfg, model = build_ref_int4_model(torch, device)           # same seed on every rank: identical full models
        max_len = prompt_len + gen_tokens + 32
        tpm = TPFusedDecodeModel(model, max_len=max_len)
        one = FusedDecodeModel(model, max_len=max_len)
        # ---- parity gate: prefill + 6 greedy steps through both wrappers
        prompt = torch.tensor([StubTokenizer(prompt_len).encode("x")], device=device)
        worst, same_tok = 0.0, True
        with torch.no_grad():
            _, lg1, h1 = one(input_ids=prompt, past_key_values=None)
            _, lgt, ht = tpm(input_ids=prompt, past_key_values=None)
            tok = lg1[0, -1].argmax().reshape(1, 1)
            for _ in range(6):
                _, lg1, h1 = one(input_ids=tok, past_key_values=h1)
                _, lgt, ht = tpm(input_ids=tok, past_key_values=ht)
                a, b = lg1[0, -1].float(), lgt[0, -1].float()
                rms = a.pow(2).mean().sqrt()
                worst = max(worst, float(((a - b).abs() / (2e-2 * a.abs() + 2e-2 * rms)).max()))
                top2 = a.topk(2).values
                if float(top2[0] - top2[1]) > 2e-2 * float(a.abs().max()):
                    same_tok = same_tok and int(a.argmax()) == int(b.argmax())
                tok = a.argmax().reshape(1, 1)
        t = torch.tensor([worst, 0.0 if same_tok else 1.0], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gate = {"worst_ratio": round(float(t[0]), 4), "same_greedy_tokens": bool(t[1] == 0), "lost_peer_words": tpm.ex.error(),
                "ok": bool(t[0] <= 1.0 and t[1] == 0) and tpm.ex.error() == 0,
                "how": "prefill + 6 greedy steps: TPFusedDecodeModel vs the single-GPU FusedDecodeModel on the same model, "
                       "|d| <= 2e-2 |ref| + 2e-2 rms, every rank"}
        del one
        # ---- timed generation, tok/s exactly as the reference's `gen` figure
        real_perf = time.perf_counter
        times = []

        class _Clock:
            @staticmethod
            def perf_counter():
                tt = real_perf()
                times.append(tt)
                return tt

            def __getattr__(self, k):
                return getattr(time, k)

        dec = ChatGLMDecoder(cfg, tpm, StubTokenizer(prompt_len), device=device, time_log=False)
        torch.manual_seed(0)
        for _ in dec.generate("warm-up", max_generated_tokens=8):
            pass
        dist.barrier()
        torch.cuda.synchronize()
        decmod.time = _Clock()
        try:
            for _ in dec.generate("bench", max_generated_tokens=gen_tokens):
                pass
        finally:
            decmod.time = time
        steps = [b - a for a, b in zip(times[0::2], times[1::2])]
        rest = steps[1:]
        tt = torch.tensor([sum(rest)], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # device time of the TP fused step alone (graph replays, KV window rewound)
        ctx = prompt_len + gen_tokens // 2
        st = torch.tensor([ctx, ctx], dtype=torch.int32, device=device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        dist.barrier()
        for i in range(reps + 5):
            if i == 5:
                e0.record()
            tpm.state[:2].copy_(st, non_blocking=True)
            tpm.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        dev = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device=device)
        dist.all_reduce(dev, op=dist.ReduceOp.MAX)
        return {"value": round(len(rest) / float(tt.item()), 2), "unit": UNIT, "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 8,
                "tokens": len(steps), "parity_gate": gate, "fused_step_device_us": round(float(dev.item()), 1),
                "fused_step_launches": tpm.launches_per_step(), "lost_peer_words": tpm.ex.error(),
                "how": f"reference ChatGLMDecoder.generate (unmodified) on every rank driving TPFusedDecodeModel (tp{world}: "
                       f"column-parallel qkv / w_in / lm_head, row-parallel o_proj / w_out with the partial sums exchanged "
                       f"inside the decode kernel over NVLink peer memory, logits all-gathered by peer stores + one barrier "
                       f"kernel); one CUDA-graph replay per token; same seed on every rank, no token broadcast; "
                       f"max over ranks of the summed step wall time"}
    finally:
        uninstall("chatglm_q")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None,
                    help="timed steps (default: 500 token steps = ~0.5 s so that the clocks sampler sees the region; "
                         "20 for --impl reference, whose step is ~0.25 s of CPU work)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--gen-tokens", type=int, default=128)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-micro", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-int8", action="store_true", help="skip the int8 model (BASELINE configs[3]) sub-dict")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 500
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
