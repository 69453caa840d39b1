"""ORACLE — test infrastructure only.  ctypes front-end for oracle/ggml_quants.c.

Restates what ``llama-quantize`` computes for the reference's GGUF path
(ref/src/quantool/methods/llama_cpp/llama_cpp.py:165-178); see the header of ggml_quants.c
for provenance and the parity-pin status of each block type.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libggml_oracle.so")

# GGUFPY/constants.py GGMLQuantizationType
GGML_TYPES = {"Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "IQ4_NL": 20, "Q2_K": 10, "Q3_K": 11, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ggml_quants.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.oracle_block_elems.argtypes = [ctypes.c_int]
        L.oracle_block_bytes.argtypes = [ctypes.c_int]
        for f in (L.oracle_quantize, L.oracle_dequantize):
            f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64]
            f.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_round_f16.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        L.oracle_f16_to_f32.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _lib = L
    return _lib


def _tid(t):
    return GGML_TYPES[t] if isinstance(t, str) else int(t)


def block_elems(t) -> int:
    return lib().oracle_block_elems(_tid(t))


def block_bytes(t) -> int:
    return lib().oracle_block_bytes(_tid(t))


def quantize(x: np.ndarray, t) -> np.ndarray:
    """x: [nrows, ncols] fp32 -> uint8 [nrows, ncols/block_elems*block_bytes]."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    nrows, ncols = x.shape
    be, bb = block_elems(t), block_bytes(t)
    assert ncols % be == 0, f"ncols {ncols} not a multiple of {be}"
    y = np.empty((nrows, ncols // be * bb), dtype=np.uint8)
    rc = lib().oracle_quantize(_tid(t), x.ctypes.data, y.ctypes.data, nrows, ncols)
    assert rc == 0
    return y


def dequantize(y: np.ndarray, t, ncols: int) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.uint8)
    nrows = y.shape[0]
    x = np.empty((nrows, ncols), dtype=np.float32)
    rc = lib().oracle_dequantize(_tid(t), y.ctypes.data, x.ctypes.data, nrows, ncols)
    assert rc == 0
    return x


def round_f16(x: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 (RNE) -> fp32, the reference's f16-GGUF intermediate."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().oracle_round_f16(x.ctypes.data, y.ctypes.data, x.size)
    return y


def set_threads(n: int) -> None:
    """0 = all online cores (llama-quantize's default nthread = hardware_concurrency)."""
    lib().oracle_set_threads(int(n))


def get_threads() -> int:
    return lib().oracle_get_threads()
