"""The C-ABI library loads without a GPU and exports every symbol include/quantool_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "quantool_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qt_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from quantool_b200.csrc import build
    so = build.build()
    lib = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/quantool_b200.h but not exported"


def test_binding_covers_the_header():
    from quantool_b200 import cabi
    assert sorted(cabi.exported_symbols()) == _declared()


def test_pure_host_entry_points_work_without_gpu():
    from quantool_b200 import cabi
    L = cabi.lib()
    assert L.qt_abi_version() == 2
    assert L.qt_gguf_block_bytes(cabi.GGML["Q4_K"]) == 144 and L.qt_gguf_block_elems(cabi.GGML["Q4_K"]) == 256
    assert L.qt_gguf_block_bytes(cabi.GGML["Q8_0"]) == 34 and L.qt_gguf_block_bytes(99) == -1


def test_product_fails_loudly_without_cuda_tensor():
    import pytest
    import torch
    from quantool_b200 import cabi
    with pytest.raises(cabi.QtError):
        cabi.gguf_quantize(torch.zeros((1, 32)), "Q8_0")
    with pytest.raises(cabi.QtError):
        cabi.hessian_accumulate(torch.zeros((8, 8), dtype=torch.bfloat16), torch.zeros((8, 8)))
