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
