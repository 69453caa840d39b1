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


SHIM_SCRIPT = '''
import json, os, sys, types
sys.path.insert(0, {root!r})
import datasets, transformers                      # before the accelerate stub: both probe for the real package
sys.path.insert(0, {refsrc!r})
for name in ("accelerate", "accelerate.commands", "accelerate.commands.config"):
    sys.modules[name] = types.ModuleType(name)
acc = types.ModuleType("accelerate.commands.config.config_args")
acc.cache_dir = "/tmp"                              # the reference's only use of accelerate: a stray import (cli.py:4)
sys.modules["accelerate.commands.config.config_args"] = acc

import quantool.methods        # the REFERENCE package (scratch copy + the shim package quantool/methods/zz_b200)
from quantool.core.registry import QuantizerRegistry
import quantool.entrypoints.cli as rcli             # the REFERENCE CLI steps, unmodified
from quantool.args import (CalibrationArguments, CommonArguments, EvaluationArguments, ExportArguments,
                           LoggingArguments, ModelArguments, QuantizationArguments)
from transformers import HfArgumentParser
from unittest.mock import patch
from quantool_b200.engine import gguf_file
from quantool_b200.methods.llm_compressor.base import LLMCompressorQuantizer

work = {work!r}
rows = [dict(text="row %d" % i) for i in range(10)]
open(os.path.join(work, "calib.jsonl"), "w").write("\\n".join(json.dumps(r) for r in rows))
parser = HfArgumentParser((ModelArguments, QuantizationArguments, CalibrationArguments, EvaluationArguments,
                           ExportArguments, CommonArguments, LoggingArguments))
names = ("model_args", "quant_args", "calibration_args", "evaluation_args", "export_args", "common_args", "logging_args")
seen, steps = [], []


class Model:
    def save_pretrained(self, dest, save_compressed=False, **_):
        open(os.path.join(dest, "model.safetensors"), "w").write("weights")


def engine(self, **kw):
    seen.append(kw)
    return Model()


def convert(model_path, out_file, outtype="f16", require_tokenizer=True):
    steps.append(("convert", outtype)); open(out_file, "w").write("gguf"); return out_file


def quantize(input_gguf, out_file, ftype, devices=None):
    steps.append(("quantize", ftype)); open(out_file, "w").write("gguf"); return out_file


def run(cfg):
    st = dict(zip(names, parser.parse_dict(cfg, allow_extra_keys=False)))
    st["model_path"], st["tokenizer"] = "/models/m", None
    rcli.validate_args_step(st)
    st = rcli.quantize_step(st)
    st = rcli.model_card_step(st)
    return rcli.save_step(st)


with patch.object(LLMCompressorQuantizer, "_oneshot", engine), \\
        patch.object(gguf_file, "convert_hf_to_f16_gguf", convert), patch.object(gguf_file, "quantize_gguf", quantize):
    st = run(dict(model_id="org/m", method="gptq", quant_level="W4A16", output_path=os.path.join(work, "save1"),
                  quantization_config=dict(method_kwargs=dict(actorder="group"), output_dir=os.path.join(work, "o1")),
                  dataset_path=os.path.join(work, "calib.jsonl"), load_in_pipeline=True, sample_size=4, shuffle=False))
    kw = seen[-1]
    out = dict(gptq_class=type(st["quantizer"]).__module__, gptq_output=st["quantized_output"],
               gptq_rows=list(kw["dataset"]["text"]), gptq_actorder=kw["recipe"].actorder, gptq_scheme=kw["recipe"].scheme,
               gptq_saved=sorted(os.listdir(os.path.join(work, "save1"))))
    st = run(dict(model_id="org/m", method="gguf", quant_level=["Q8_0", "Q4_K_M"], output_path=os.path.join(work, "save2"),
                  quantization_config=dict(llama_cpp_path=None)))
    out.update(gguf_class=type(st["quantizer"]).__module__, gguf_output=[os.path.basename(p) for p in st["quantized_output"]],
               gguf_steps=steps, gguf_saved=sorted(os.listdir(os.path.join(work, "save2"))))
print("SHIM_JSON " + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="the reference checkout is not on this machine")
def test_integration_shim_lets_the_reference_cli_drive_our_plugins(tmp_path):
    """INTEGRATION.md section 1 executed for real: a scratch copy of the REFERENCE package gets the documented shim
    sub-package, the reference's own loader imports it, and the reference's unmodified CLI steps
    (validate_args -> quantize -> generate_readme -> save_model) then run a GPTQ and a GGUF configuration.  Only the
    engine entry points are recorders here (no GPU in this container)."""
    import json
    # a scratch copy of the reference package (outside this repository) with the shim added the way INTEGRATION.md
    # tells a maintainer to: one new sub-package of quantool/methods, imported last by the reference's own loader
    refsrc = tmp_path / "refsrc"
    shutil.copytree("/root/reference/src/quantool", refsrc / "quantool")
    shim = refsrc / "quantool" / "methods" / "zz_b200"
    shim.mkdir()
    (shim / "__init__.py").write_text(
        "from quantool.core.registry import QuantizerRegistry\n"
        "import quantool_b200.methods\n"
        "from quantool_b200 import QuantizerRegistry as B200\n\n"
        "for name in (\"gptq\", \"awq\", \"smoothquant\", \"gguf\"):\n"
        "    QuantizerRegistry._plugins[name] = B200._plugins[name]\n")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    r = subprocess.run([sys.executable, "-c", SHIM_SCRIPT.format(root=ROOT, work=str(tmp_path), refsrc=str(refsrc))],
                       capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    got = [l for l in r.stdout.splitlines() if l.startswith("SHIM_JSON ")]
    assert got, (r.stdout + r.stderr)[-3000:]
    o = json.loads(got[-1][len("SHIM_JSON "):])
    assert o["gptq_class"] == "quantool_b200.methods.llm_compressor.gptq"
    assert o["gptq_output"] == str(tmp_path / "o1") and o["gptq_rows"] == [f"row {i}" for i in range(4)]
    assert o["gptq_actorder"] == "group" and o["gptq_scheme"] == "W4A16"
    assert o["gptq_saved"] == ["README.md", "model.safetensors"]
    assert o["gguf_class"] == "quantool_b200.methods.llama_cpp.llama_cpp"
    assert o["gguf_output"] == ["m-Q8_0.gguf", "m-Q4_K_M.gguf"]
    assert o["gguf_steps"] == [["convert", "f16"], ["quantize", "Q8_0"], ["quantize", "Q4_K_M"]]
    assert o["gguf_saved"] == ["README.md", "m-Q4_K_M.gguf", "m-Q8_0.gguf"]
