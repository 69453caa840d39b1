"""ctypes binding of the C-ABI library (include/quantool_b200.h).

Host glue only: unwraps torch tensors into device pointers + sizes and passes torch's current
CUDA stream.  There is NO CPU fallback: if the library is missing or the tensor is not on a
CUDA device the call raises.
"""
import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libquantool_b200.so")

QT_F32, QT_F16, QT_BF16 = 0, 1, 2
_DT = {torch.float32: QT_F32, torch.float16: QT_F16, torch.bfloat16: QT_BF16}

GGML = {"Q4_0": 2, "Q4_1": 3, "Q5_0": 6, "Q5_1": 7, "Q8_0": 8, "Q4_K": 12, "Q5_K": 13, "Q6_K": 14}

_i64, _i32, _vp, _f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPE
_SIGS = {
    "qt_abi_version": [],
    "qt_device_sm_count": [],
    "qt_gguf_block_elems": [_i32],
    "qt_gguf_block_bytes": [_i32],
    "qt_gguf_quantize": [_i32, _vp, _i32, _i32, _i64, _i64, _vp, _vp],
    "qt_gguf_dequantize": [_i32, _vp, _i64, _i64, _vp, _vp],
}
_RESTYPE = {"qt_last_error": ctypes.c_char_p}

_lib: Optional[ctypes.CDLL] = None


class QtError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the C-ABI library; fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QtError(f"{LIB_PATH} not found: build it with `python -m quantool_b200.csrc.build` "
                          "(there is no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        L.qt_last_error.restype = ctypes.c_char_p
        L.qt_last_error.argtypes = []
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, ctypes.c_int)
        _lib = L
    return _lib


def exported_symbols():
    return ["qt_last_error"] + list(_SIGS)


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().qt_last_error()
        raise QtError(f"{what} failed: rc={rc} ({msg.decode() if msg else ''})")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise QtError(f"{name} must be a CUDA tensor (no CPU fallback)")
    if not t.is_contiguous():
        raise QtError(f"{name} must be contiguous")
    return t


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


# ---------------------------------------------------------------------------------------
# GGUF
# ---------------------------------------------------------------------------------------
def gguf_block_elems(t: str) -> int:
    return lib().qt_gguf_block_elems(GGML[t])


def gguf_block_bytes(t: str) -> int:
    return lib().qt_gguf_block_bytes(GGML[t])


def gguf_quantize(x: torch.Tensor, qtype: str, round_via_f16: bool = True,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [nrows, ncols] (f32/f16/bf16, CUDA) -> uint8 [nrows, ncols/be*bb] packed blocks."""
    _dev(x, "x")
    assert x.dim() == 2
    nrows, ncols = x.shape
    be, bb = gguf_block_elems(qtype), gguf_block_bytes(qtype)
    if ncols % be:
        raise QtError(f"ncols={ncols} is not a multiple of the {qtype} block size {be}")
    if out is None:
        out = torch.empty((nrows, ncols // be * bb), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().qt_gguf_quantize(GGML[qtype], _p(x), _DT[x.dtype], int(round_via_f16), nrows, ncols,
                                      _p(out), _stream()), "qt_gguf_quantize")
    return out


def gguf_dequantize(y: torch.Tensor, qtype: str, ncols: int) -> torch.Tensor:
    _dev(y, "y")
    nrows = y.shape[0]
    out = torch.empty((nrows, ncols), dtype=torch.float32, device=y.device)
    with torch.cuda.device(y.device):
        _check(lib().qt_gguf_dequantize(GGML[qtype], _p(y), nrows, ncols, _p(out), _stream()),
               "qt_gguf_dequantize")
    return out
