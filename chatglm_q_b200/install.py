"""Bind the sm_100a kernels behind an UNMODIFIED chatglm_q package (seam S1).

`chatglm_q/int4/qlinear.py:7-17` and `chatglm_q/int8/qlinear.py:6-16` import their kernel entry
points into module globals and look them up at call time (`int4/qlinear.py:47-48`), so rebinding
those globals swaps the kernel for every existing module instance with zero reference edits.
"""
from __future__ import annotations

import importlib

from . import ops

_saved: dict[str, dict[str, object]] = {}


def install(package: str = "chatglm_q", sampler=False) -> None:
    """Rebind `<package>.int4.qlinear` / `<package>.int8.qlinear` kernel globals to chatglm_q_b200.

    `sampler=True` also rebinds `<package>.decoder.top_p_sampling` (decoder.py:12-27), which
    `ChatGLMDecoder.generate` resolves by global name at call time (decoder.py:85), to the one-launch
    `ops.top_p_sampling` (same signature, same token for the same torch seed; CUDA fp16 / bf16 logits only).
    `sampler=<callable>` binds that callable instead (e.g. `FusedDecodeModel.sampler()`).

    What this library does not build stays with the reference: fp32 (TF32) activations and int4 group sizes other
    than 32 go to the saved original kernels, CPU / fp32 logits and top_k > 1024 to the saved original sampler
    (`ops._delegates`), so a configuration that works in the reference keeps working after install()."""
    q4 = importlib.import_module(f"{package}.int4.qlinear")
    q8 = importlib.import_module(f"{package}.int8.qlinear")
    for mod, impl, impl_t, slot in ((q4, ops.dynamic_quant_matmul_s4, ops.dynamic_quant_matmul_transposed_s4, "s4"),
                                    (q8, ops.dynamic_quant_matmul, ops.dynamic_quant_matmul_transposed, "s8")):
        if mod.__name__ not in _saved:
            _saved[mod.__name__] = {
                k: getattr(mod, k, None)
                for k in ("_dynamic_quant_matmul_impl", "_dynamic_quant_matmul_transposed_impl", "check_input", "KERNEL_IMPL")
            }
        orig = _saved[mod.__name__]["_dynamic_quant_matmul_impl"]
        orig_t = _saved[mod.__name__]["_dynamic_quant_matmul_transposed_impl"]
        ops._delegates[slot] = orig if callable(orig) else None
        ops._delegates[slot + "t"] = orig_t if callable(orig_t) else None
        mod._dynamic_quant_matmul_impl = impl
        mod._dynamic_quant_matmul_transposed_impl = impl_t      # DynamicQuantizeMatMul.backward (int4/qlinear.py:53-64)
        mod.check_input = ops.check_input
        mod.KERNEL_IMPL = "cgq_b200"
    if sampler:
        dec = importlib.import_module(f"{package}.decoder")
        if dec.__name__ not in _saved:
            _saved[dec.__name__] = {"top_p_sampling": dec.top_p_sampling}
        orig = _saved[dec.__name__]["top_p_sampling"]
        ops._delegates["sampler"] = orig if callable(orig) else None
        dec.top_p_sampling = sampler if callable(sampler) else ops.top_p_sampling


def uninstall(package: str = "chatglm_q") -> None:
    ops._delegates.update({"s4": None, "s8": None, "s4t": None, "s8t": None, "sampler": None})
    for name in (f"{package}.int4.qlinear", f"{package}.int8.qlinear", f"{package}.decoder"):
        saved = _saved.pop(name, None)
        if saved is None:
            continue
        mod = importlib.import_module(name)
        for k, v in saved.items():
            setattr(mod, k, v)
