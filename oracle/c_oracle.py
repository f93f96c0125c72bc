"""ORACLE — test infrastructure.  ctypes loader for the plain-C restatement (cgq_oracle.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
_DT = {"float16": 0, "bfloat16": 1, "float32": 2}
_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", str(HERE), "all"], check=True)


def _cpu_has_v3() -> bool:
    try:
        flags = Path("/proc/cpuinfo").read_text()
    except OSError:
        return False
    return all(f" {f}" in flags for f in ("avx2", "f16c", "bmi2"))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        name = "liboracle_v3.so" if _cpu_has_v3() else "liboracle_base.so"
        path = HERE / name
        if not path.exists():
            build()
        _lib = ctypes.CDLL(str(path))
    return _lib


def _f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(x):
    return None if x is None else x.ctypes.data_as(ctypes.c_void_p)


def set_threads(n: int) -> None:
    os.environ["OMP_NUM_THREADS"] = str(n)


def w4_unpack_i8(wq: np.ndarray) -> np.ndarray:
    k, n = wq.shape[0] * 2, wq.shape[1]
    out = np.empty((k, n), dtype=np.int8)
    load().oracle_w4_unpack_i8(_p(np.ascontiguousarray(wq)), _p(out), k, n)
    return out


def w4_dequant(wq: np.ndarray, scale: np.ndarray, dtype: str, group: int = 32) -> np.ndarray:
    k, n = wq.shape[0] * 2, wq.shape[1]
    out = np.empty((k, n), dtype=np.float32)
    load().oracle_w4_dequant(_p(np.ascontiguousarray(wq)), _p(_f32(scale)), _p(out), k, n, group,
                             _DT[dtype])
    return out


def w4a16_gemm(a, wq, scale, bias, dtype: str, group: int = 32, scratch=None) -> np.ndarray:
    a = _f32(a)
    m, k = a.shape
    n = wq.shape[1]
    c = np.empty((m, n), dtype=np.float32)
    if scratch is None:
        scratch = np.empty((k, n), dtype=np.float32)
    bias = None if bias is None else _f32(bias)
    load().oracle_w4a16_gemm(_p(a), _p(np.ascontiguousarray(wq)), _p(_f32(scale)), _p(bias), _p(c),
                             m, n, k, group, _DT[dtype], _p(scratch))
    return c


def w8a16_gemm(a, wq_nk, scale, bias, dtype: str, scratch=None) -> np.ndarray:
    a = _f32(a)
    m, k = a.shape
    n = wq_nk.shape[0]
    c = np.empty((m, n), dtype=np.float32)
    if scratch is None:
        scratch = np.empty((k, n), dtype=np.float32)
    bias = None if bias is None else _f32(bias)
    load().oracle_w8a16_gemm(_p(a), _p(np.ascontiguousarray(wq_nk)), _p(_f32(scale)), _p(bias), _p(c),
                             m, n, k, _DT[dtype], _p(scratch))
    return c
