"""The reference's OWN unit tests for the path's boundary, run against quantool_b200.

`/root/reference/tests/quantool/methods/test_llama_cpp.py` (plugin registration, multi-level GGUF flow) and
`/root/reference/tests/quantool/utils/test_dataset_textifier.py` (chat-template rendering of calibration rows) are
executed unmodified in a child pytest whose conftest aliases the `quantool.*` module names they import to the
quantool_b200 modules that replace them - the maintainer-side shim of INTEGRATION.md section 1 in test form.
The test files are read from /root/reference at run time and are never copied into this repository, so the test
skips where the reference is absent (the GPU box)."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests/quantool"

CONFTEST = '''
import sys
import types

sys.path.insert(0, {root!r})
import quantool_b200
import quantool_b200.methods
from quantool_b200.core import registry
from quantool_b200.methods.llama_cpp import llama_cpp
from quantool_b200.methods.llm_compressor import chat


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


textifier = _module("quantool.utils.dataset_textifier", convert_row=chat.render_chat_row,
                    has_chat_template=chat.has_chat_template, is_conversational=chat.is_conversational)
_module("quantool")
_module("quantool.core")
_module("quantool.methods")
_module("quantool.methods.llama_cpp")
_module("quantool.utils", dataset_textifier=textifier)
sys.modules["quantool.core.registry"] = registry
sys.modules["quantool.methods.llama_cpp.llama_cpp"] = llama_cpp
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not on this machine")
def test_reference_unit_tests_pass_against_quantool_b200(tmp_path):
    work = tmp_path / "ref_tests"
    work.mkdir()
    (work / "conftest.py").write_text(CONFTEST.format(root=ROOT))
    for rel in ("methods/test_llama_cpp.py", "utils/test_dataset_textifier.py"):
        shutil.copy(os.path.join(REF_TESTS, rel), work / os.path.basename(rel))       # scratch copy, outside the repo
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", str(work)], capture_output=True,
                       text=True, cwd=str(work), env=env, timeout=600)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    summary = [l for l in r.stdout.splitlines() if " passed" in l][-1]
    assert "failed" not in summary and "error" not in summary, tail
    assert int(summary.split(" passed")[0].split()[-1]) >= 23, summary           # 2 plugin tests + 21 textifier cases
