"""ctypes binding of the C-ABI library (include/cgq.h).  Fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libcgq.so"

# every symbol include/cgq.h declares (tests check the library exports all of them)
SYMBOLS = {
    "cgq_version": (c_int, []),
    "cgq_last_error": (c_char_p, []),
    "cgq_workspace_bytes": (c_size_t, []),
    "cgq_w4a16_gemm": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                               c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cgq_w4a16_gemm_ex": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p,
                                  c_int]),
    "cgq_w8a16_gemm": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                               c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cgq_w8a16_gemm_ex": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_int]),
    "cgq_debug_trace": (None, [c_void_p]),
    "cgq_set_decode_arith": (c_int, [c_int]),
    "cgq_simple_fallback_count": (ctypes.c_ulonglong, []),
    "cgq_forbid_simple": (c_int, [c_int]),
    "cgq_attention_next_kv": (None, [c_void_p, c_void_p]),
    "cgq_w4a16_gemv_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p]),
    "cgq_w8a16_gemv_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_void_p, c_float, c_void_p]),
    "cgq_decode_begin_w8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cgq_handover_next": (c_int, [c_void_p, ctypes.c_uint32, c_void_p]),
    "cgq_w4_gemv_tiles": (c_int, [c_int]),
    "cgq_program_create": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "cgq_program_run": (c_int, [ctypes.c_uint64, c_void_p]),
    "cgq_program_status": (c_int, [ctypes.c_uint64, c_void_p, c_void_p]),
    "cgq_program_destroy": (c_int, [ctypes.c_uint64]),
    "cgq_step_create": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "cgq_step_run": (c_int, [ctypes.c_uint64, c_void_p]),
    "cgq_step_status": (c_int, [ctypes.c_uint64, c_void_p, c_void_p, c_void_p]),
    "cgq_step_destroy": (c_int, [ctypes.c_uint64]),
    "cgq_w4a16_grad_a": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "cgq_w8a16_grad_a": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "cgq_tp_next": (c_int, [c_void_p, ctypes.c_uint32]),
    "cgq_tp_barrier": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cgq_ipc_alloc": (c_int, [c_size_t, c_void_p, c_void_p]),
    "cgq_ipc_open": (c_int, [c_void_p, c_void_p]),
    "cgq_ipc_close": (c_int, [c_void_p]),
    "cgq_ipc_free": (c_int, [c_void_p]),
    "cgq_prefetch_next_w4": (c_int, [c_void_p, c_void_p, c_int, c_int]),
    "cgq_decode_begin_w4": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                    c_int, c_void_p, c_void_p]),
    "cgq_decode_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p]),
    "cgq_top_p_sample": (c_int, [c_void_p, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "cgq_w4_unpack_i8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "cgq_w4_dequant": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cgq_w4_embedding": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_void_p]),
    "cgq_w8_embedding": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_void_p]),
}

IMPL_AUTO, IMPL_SIMPLE, IMPL_GEMV, IMPL_GEMV_EXACT, IMPL_TC, IMPL_GEMV_UMMA = 0, 1, 2, 3, 4, 5
IMPL_GEMV_SUBNORMAL, IMPL_GEMV_IMMA = 6, 7
ARITH_IMMA, ARITH_EXACT, ARITH_SUBNORMAL = 0, 1, 2      # cgq_set_decode_arith
PRO_NONE, PRO_RMSNORM, PRO_SILU_GATE = 0, 1, 2

_lib = None


class LinearOp(ctypes.Structure):
    """`cgq_linear_op` of include/cgq.h (one batch-1 int4g32 linear of a persistent decode program)."""
    _fields_ = [("Wq", c_void_p), ("scale", c_void_p), ("bias", c_void_p), ("A", c_void_p), ("C", c_void_p),
                ("resid", c_void_p), ("norm_w", c_void_p), ("N", c_int), ("K", c_int), ("prologue", c_int),
                ("eps", c_float)]


STEP_LINEAR, STEP_ATTENTION, STEP_EMBED = 0, 1, 2
EPI_NONE, EPI_SILU_PAIR = 0, 1


class StepOp(ctypes.Structure):
    """`cgq_step_op` of include/cgq.h (one phase of the one-launch decode step)."""
    _fields_ = [("kind", c_int), ("Wq", c_void_p), ("scale", c_void_p), ("bias", c_void_p), ("A", c_void_p),
                ("C", c_void_p), ("resid", c_void_p), ("norm_w", c_void_p), ("N", c_int), ("K", c_int),
                ("prologue", c_int), ("eps", c_float), ("epilogue", c_int), ("freqs", c_void_p), ("kcache", c_void_p), ("vcache", c_void_p),
                ("n_head", c_int), ("n_groups", c_int), ("d_head", c_int), ("max_len", c_int), ("ids", c_void_p),
                ("V", c_int)]


class TpCtx(ctypes.Structure):
    """`cgq_tp_ctx` of include/cgq.h (tensor-parallel exchange of the fused decode step)."""
    _fields_ = [("world", c_int), ("rank", c_int), ("max_n", c_int), ("out_offset", c_int), ("recv", c_void_p * 8),
                ("step", c_void_p), ("err", c_void_p), ("out", c_void_p * 8)]


class CgqError(RuntimeError):
    """A C-ABI call returned a non-zero status (message from cgq_last_error())."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m chatglm_q_b200.build` "
            "(there is no CPU / PyTorch fallback for this path)")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().cgq_last_error()
        raise CgqError(f"cgq status {status}: {msg.decode() if msg else ''}")
