"""The product path may not import, call, link or execute anything under oracle/."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_package_never_references_oracle():
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "quantool_b200")):
        for fn in files:
            if not fn.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            src = open(os.path.join(dp, fn), errors="ignore").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M) or "libggml_oracle" in src:
                bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_no_cpu_fallback_or_compat_layers():
    for dp, _, files in os.walk(os.path.join(ROOT, "quantool_b200")):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert "import triton" not in src and "torch.compile" not in src, fn


def test_product_path_fails_loudly_without_cuda_or_library(tmp_path, monkeypatch):
    """No CPU fallback anywhere on the path: host tensors, a missing CUDA device and a missing library all raise."""
    import pytest
    import torch
    from quantool_b200 import cabi
    from quantool_b200.engine import gguf_file
    from quantool_b200.methods.llm_compressor.base import Modifier
    from quantool_b200.methods.llm_compressor.gptq import GPTQ
    x = torch.zeros((4, 256))
    with pytest.raises(cabi.QtError, match="CUDA tensor"):
        cabi.gguf_quantize(x, "Q8_0")
    with pytest.raises(cabi.QtError, match="CUDA tensor"):
        cabi.pack_int32(torch.zeros((4, 8), dtype=torch.int8), 4)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            gguf_file.quantize_gguf(str(tmp_path / "in.gguf"), str(tmp_path / "out.gguf"), "Q8_0")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            GPTQ(model_id="m")._oneshot(model=str(tmp_path), recipe=Modifier(kind="gptq", scheme="W4A16"),
                                        output_dir=str(tmp_path / "o"), dataset=torch.zeros((1, 4), dtype=torch.long))
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", str(tmp_path / "libquantool_b200.so"))
    with pytest.raises(cabi.QtError, match="not found"):
        cabi.lib()
