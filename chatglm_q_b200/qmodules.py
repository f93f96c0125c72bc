"""Module-level seam (S2): nn.Modules that can stand in for the reference's
`DynamicQuantizeLinear` / `QEmbedding` when `loader.create_quant_int{4,8}_model` assigns
`modeling.Linear` / `modeling.Embedding` (chatglm_q/loader.py:41-66).

What the reference loader and model rely on, and what is therefore kept identical:
  * constructor keywords `(in_features, out_features, bias=..., device=..., dtype=...)`
    (call sites chatglm_q/model.py:111-112,194-195,253,262);
  * BUFFERS named `weight`, `weight_scale`, `bias` with the reference shapes / dtypes
    (int4: u8 [K/2, N], [K/32, N], [N] — int4/qlinear.py:83-88; int8: i8 [N, K], [N], [N] —
    int8/qlinear.py:82-87), so `state_dict()` keys and the loader's by-key `copy_` still work;
  * `apply_weights_` as used by the quantiser scripts.
Forward runs on the sm_100a kernels with the bias add fused (same two roundings as the
reference's separate `out += bias`).  `dynamic_quant_matmul_int4/8` are differentiable in the activation
(grad_A on `cgq_w4a16_grad_a` / `cgq_w8a16_grad_a`); the Linear modules' fused-bias forward is inference-only.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from . import ops

GROUP = 32  # the only group size the reference model ever builds


class _QuantModule(nn.Module):
    """Shared plumbing: quantised buffers are filled in place, never re-initialised."""

    _fields: tuple[str, ...] = ()

    @torch.no_grad()
    def apply_weights_(self, q_weight: Tensor, scale: Tensor, bias: Tensor = None):
        self.weight.copy_(q_weight)
        self.weight_scale.copy_(scale)
        if bias is not None:
            self.bias.copy_(bias)

    def reset_parameters(self):  # buffers come from a checkpoint
        pass

    def extra_repr(self) -> str:
        return ", ".join(f"{f}={getattr(self, f)}" for f in self._fields)


class _QLinear(_QuantModule):
    def _alloc_bias(self, bias: bool, n: int, device, dtype):
        self.register_buffer("bias", torch.empty(n, device=device, dtype=dtype) if bias else None)

    @property
    def has_bias(self) -> bool:
        return self.bias is not None


class W4Linear(_QLinear):
    """int4g32 linear; stands in for chatglm_q.int4.qlinear.DynamicQuantizeLinear (:75-108)."""

    _fields = ("in_features", "out_features", "group_size", "has_bias")

    def __init__(self, in_features: int, out_features: int, bias=True, group_size=GROUP,
                 device=None, dtype=None):
        super().__init__()
        assert in_features % group_size == 0, f"{in_features=}, {group_size=}"
        assert group_size == GROUP, f"W4Linear is built for the reference model's group size {GROUP}, got {group_size}"
        self.in_features, self.out_features = in_features, out_features
        self.group_size, self.groups = group_size, in_features // group_size
        self.register_buffer(
            "weight", torch.empty((in_features // 2, out_features), device=device, dtype=torch.uint8))
        self.register_buffer(
            "weight_scale", torch.empty((self.groups, out_features), device=device, dtype=dtype))
        self._alloc_bias(bias, out_features, device, dtype)

    def forward(self, input: Tensor) -> Tensor:
        return ops.dynamic_quant_matmul_s4(input, self.weight, self.weight_scale, bias=self.bias)


class W8Linear(_QLinear):
    """int8 per-channel linear; stands in for chatglm_q.int8.qlinear.DynamicQuantizeLinear (:77-107)."""

    _fields = ("in_features", "out_features", "has_bias")

    def __init__(self, in_features: int, out_features: int, bias: bool = True, device=None, dtype=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.register_buffer(
            "weight", torch.empty((out_features, in_features), device=device, dtype=torch.int8))
        self.register_buffer("weight_scale", torch.empty(out_features, device=device, dtype=dtype))
        self._alloc_bias(bias, out_features, device, dtype)

    def forward(self, input: Tensor) -> Tensor:
        # the kernel consumes the [N, K] buffer itself; `.t()` keeps the reference call shape
        return ops.dynamic_quant_matmul(input, self.weight.t(), self.weight_scale, bias=self.bias)


class W4Embedding(_QuantModule):
    """Vocab-axis-packed int4 embedding; stands in for int4 QEmbedding (int4/qlinear.py:111-142)."""

    _fields = ("num_embeddings", "embedding_dim", "group_size")

    def __init__(self, num_embeddings: int, embedding_dim: int, group_size=GROUP, device=None,
                 dtype=None):
        super().__init__()
        assert num_embeddings % group_size == 0, f"{num_embeddings=}, {group_size=}"
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.group_size, self.groups = group_size, num_embeddings // group_size
        self.register_buffer(
            "weight", torch.empty((num_embeddings // 2, embedding_dim), device=device, dtype=torch.uint8))
        self.register_buffer(
            "weight_scale", torch.empty((self.groups, embedding_dim), device=device, dtype=dtype))

    def forward(self, input: Tensor) -> Tensor:
        return ops.embedding_s4(input, self.weight, self.weight_scale, self.group_size)


class W8Embedding(_QuantModule):
    """int8 embedding with per-feature scale; stands in for int8 QEmbedding (int8/qlinear.py:110-132)."""

    _fields = ("num_embeddings", "embedding_dim")

    def __init__(self, num_embeddings: int, embedding_dim: int, device=None, dtype=None):
        super().__init__()
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.register_buffer(
            "weight", torch.empty((num_embeddings, embedding_dim), device=device, dtype=torch.int8))
        self.register_buffer("weight_scale", torch.empty(embedding_dim, device=device, dtype=dtype))

    def forward(self, input: Tensor) -> Tensor:
        return ops.embedding_s8(input, self.weight, self.weight_scale)


class _QMatMulS4(torch.autograd.Function):
    """`DynamicQuantizeMatMul` of the int4 model (int4/qlinear.py:36-64): forward and grad_A on this repo's kernels."""

    @staticmethod
    def forward(ctx, A: Tensor, B: Tensor, b_scale: Tensor):
        ctx.save_for_backward(B, b_scale)
        return ops.dynamic_quant_matmul_s4(A, B, b_scale)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        B, b_scale = ctx.saved_tensors
        grad_A = ops.dynamic_quant_matmul_transposed_s4(grad_out, B, b_scale) if ctx.needs_input_grad[0] else None
        return grad_A, None, None


class _QMatMulS8(torch.autograd.Function):
    """`DynamicQuantizeMatMul` of the int8 model (int8/qlinear.py:24-52)."""

    @staticmethod
    def forward(ctx, A: Tensor, B: Tensor, b_scale: Tensor):
        ctx.save_for_backward(B, b_scale)
        return ops.dynamic_quant_matmul(A, B, b_scale)

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        B, b_scale = ctx.saved_tensors
        grad_A = ops.dynamic_quant_matmul_transposed(grad_out, B, b_scale) if ctx.needs_input_grad[0] else None
        return grad_A, None, None


def dynamic_quant_matmul_int4(A: Tensor, B: Tensor, b_scale: Tensor) -> Tensor:
    """chatglm_q.int4.qlinear.dynamic_quant_matmul (int4/qlinear.py:71-72): differentiable in A."""
    if A.requires_grad and torch.is_grad_enabled():
        return _QMatMulS4.apply(A, B, b_scale)
    return ops.dynamic_quant_matmul_s4(A, B, b_scale)


def dynamic_quant_matmul_int8(A: Tensor, B: Tensor, b_scale: Tensor) -> Tensor:
    """chatglm_q.int8.qlinear.dynamic_quant_matmul (int8/qlinear.py:73-74): differentiable in A."""
    if A.requires_grad and torch.is_grad_enabled():
        return _QMatMulS8.apply(A, B, b_scale)
    return ops.dynamic_quant_matmul(A, B, b_scale)
