"""Load-time path of the dequant-matmul kernels: checkpoint tensors -> device buffers.

The reference (`chatglm_q/loader.py:90-107`) builds the model on the CPU, reads every tensor with
`safe_open(...).get_tensor(k)` and copies it into the module buffer by key.  The on-disk layout of the quantised
tensors IS the kernels' layout (uint8 [K/2, N] + scales [K/32, N] for int4g32, int8 [N, K] + scales [N] for int8), so
nothing is re-packed here either:

  * `load_state_into(model, files, device)` -- the reference's by-key `copy_` contract (same keys, same "ignored" /
    "not initialized" reporting, floating-point tensors cast to the buffer's dtype), but each tensor goes
    file -> pinned staging -> its device buffer one at a time: no second full copy of the model ever exists on the host
    or the device;
  * `load_tp_shards(files, dims, world, rank, device)` -- the tensor-parallel decode step only needs this rank's
    shards (chatglm_q_b200/tp.py): column slices of qkv / w_in / lm_head, k-row slices of o_proj / w_out.  They are
    cut with safetensors' lazy slicing, so a rank reads only its own rows of the row-parallel weights and never
    materialises the other ranks' columns on the device.

torch and safetensors are plumbing here (file IO, device memory); the product is the kernels these buffers feed.
"""
from __future__ import annotations

from pathlib import Path
from typing import Iterable

import torch
from torch import Tensor

from . import tp


def _open(path):
    from safetensors import safe_open

    return safe_open(str(path), framework="pt")


def load_state_into(model: torch.nn.Module, files: Iterable[str | Path], device=None, verbose: bool = True) -> list[str]:
    """By-key copy of the checkpoint `files` into `model.state_dict()`'s tensors (chatglm_q/loader.py:90-104), moving
    the model to `device` first if given.  Returns the keys of the model that no file initialised."""
    if device is not None:
        model.to(device)
    state = dict(**model.state_dict())
    pinned = None
    for file in files:
        with _open(file) as f:
            for k in f.keys():
                if k not in state:
                    if verbose:
                        print(f'"{k}" is ignored')
                    continue
                dst = state.pop(k)
                v = f.get_tensor(k)
                if dst.is_floating_point():
                    v = v.type_as(dst) if v.device == dst.device else v.to(dst.dtype)
                assert v.shape == dst.shape, f'"{k}": checkpoint {tuple(v.shape)} vs model {tuple(dst.shape)}'
                if dst.device.type == "cuda":
                    # one pinned staging buffer, reused: the H2D copy of tensor i overlaps the file read of i + 1
                    n = v.numel() * v.element_size()
                    if pinned is None or pinned.numel() < n:
                        pinned = torch.empty(max(n, 64 << 20), dtype=torch.uint8).pin_memory()
                    torch.cuda.current_stream(dst.device).synchronize()
                    stage = pinned[:n].view(v.dtype).view(v.shape)
                    stage.copy_(v)
                    dst.copy_(stage, non_blocking=True)
                else:
                    dst.copy_(v)
    if any(t.device.type == "cuda" for t in model.state_dict().values()):
        torch.cuda.synchronize()
    missing = list(state.keys())
    if missing and verbose:
        print(f'model weights "{", ".join(missing)}" are not initialized')
    return missing


def _slice(f, key: str, rows: tuple[int, int] | None, cols: tuple[tuple[int, int], ...] | None) -> Tensor:
    """Rows [r0, r1) and the concatenated column ranges of tensor `key`, read lazily (only those rows leave the file)."""
    sl = f.get_slice(key)
    part = sl[rows[0]:rows[1]] if rows is not None else sl[:]
    if cols is not None:
        part = torch.cat([part[..., a:b] for a, b in cols], dim=-1)
    return part.contiguous()


def load_tp_shards(files: Iterable[str | Path], dims: tp.ModelDims, n_layers: int, world: int, rank: int, device,
                   dtype: torch.dtype = torch.float16) -> dict:
    """This rank's shards of an int4g32 checkpoint for the tensor-parallel decode step, keyed like the reference's
    state_dict (`layers.{i}.attn.qkv_proj.weight` ...): the same tensors `tp.shard_w4` cuts from a loaded model.
    Replicated tensors (norm weights, embedding, rotary table) are returned whole."""
    plan = tp.plan_block(world, rank, dims)
    by_suffix = {"attn.qkv_proj": plan.qkv, "attn.o_proj": plan.o, "ffn.w_in": plan.w_in, "ffn.w_out": plan.w_out}
    out: dict[str, Tensor] = {}

    def shard_of(key: str) -> tuple[tp.Shard | None, str]:
        for suf, sh in by_suffix.items():
            for leaf in ("weight", "weight_scale", "bias"):
                if key.endswith(f"{suf}.{leaf}"):
                    return sh, leaf
        for leaf in ("weight", "weight_scale", "bias"):
            if key == f"lm_head.{leaf}":
                return plan.lm_head, leaf
        return None, ""

    for file in files:
        with _open(file) as f:
            for k in f.keys():
                sh, leaf = shard_of(k)
                if sh is None:
                    v = f.get_tensor(k)
                else:
                    rows = None
                    if sh.krows is not None:
                        k0, k1 = sh.krows
                        assert k0 % tp.GROUP == 0 and k1 % tp.GROUP == 0
                        rows = {"weight": (k0 // 2, k1 // 2), "weight_scale": (k0 // tp.GROUP, k1 // tp.GROUP)}.get(leaf)
                    if leaf == "bias":
                        v = f.get_tensor(k)
                        if sh.cols is not None:
                            v = torch.cat([v[a:b] for a, b in sh.cols])
                    else:
                        v = _slice(f, k, rows, sh.cols)
                if v.is_floating_point():
                    v = v.to(dtype)
                out[k] = v.to(device, non_blocking=True)
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()
    assert n_layers <= 0 or any(k.startswith(f"layers.{n_layers - 1}.") for k in out), "checkpoint has fewer layers"
    return out
