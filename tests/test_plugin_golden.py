"""Drop-in boundary pinned against the reference ITSELF: tests/golden/plugin_traces.json holds what
the reference's own GPTQ / AWQ / SmoothQuant / GGUF plugin classes did for a fixed list of calls when run in the
build container with recording stubs in place of llm-compressor and llama.cpp
(tests/golden/make_plugin_golden.py).  The same calls are replayed on quantool_b200's plugins with the engine
entry points (`_oneshot`, `gguf_file.convert_hf_to_f16_gguf`, `gguf_file.quantize_gguf`) replaced by recorders:
return values, the keyword arguments that reach the engine, the recipe, the conversion / quantization steps and
their file names, `last_*` attributes, saved files and exceptions (type and message) must be the reference's."""
import dataclasses
import json
import os
from unittest.mock import patch

import pytest

import quantool_b200.methods  # noqa: F401
from quantool_b200 import QuantizerRegistry
from quantool_b200.methods.llama_cpp.llama_cpp import QuantType
from quantool_b200.methods.llm_compressor.base import ONESHOT_PARAMS, Modifier

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "plugin_traces.json")))
KIND = {"gptq": "GPTQModifier", "awq": "AWQModifier", "smoothquant": "SmoothQuantModifier"}


def _untag(v, out, cwd):
    if isinstance(v, str):
        return v.replace("<OUT>", out).replace("<CWD>", cwd)
    if isinstance(v, list):
        return [_untag(x, out, cwd) for x in v]
    if isinstance(v, dict):
        return {k: _untag(x, out, cwd) for k, x in v.items()}
    return v


def _recipe_as_golden(recipe):
    """Our Modifier dataclass -> the reference's (modifier class, constructor kwargs): every field that differs from
    the dataclass default must have been passed explicitly, and vice versa."""
    if isinstance(recipe, (list, tuple)):
        return [_recipe_as_golden(m) for m in recipe]
    assert isinstance(recipe, Modifier)
    return {"modifier": KIND[recipe.kind], "fields": dataclasses.asdict(recipe)}


def _check_recipe(ours, gold):
    if isinstance(gold, list):
        assert isinstance(ours, list) and len(ours) == len(gold)
        for o, g in zip(ours, gold):
            _check_recipe(o, g)
        return
    assert ours["modifier"] == gold["modifier"]
    defaults = dataclasses.asdict(Modifier(kind="x"))
    for k, v in gold["kwargs"].items():
        assert ours["fields"][k] == v, (k, ours["fields"][k], v)
    for k, v in ours["fields"].items():
        if k not in gold["kwargs"] and k != "kind":
            assert v == defaults[k], f"field {k}={v!r} was not passed by the reference"


def test_oneshot_parameter_list_is_the_golden_one():
    assert sorted(ONESHOT_PARAMS) == sorted(GOLD["oneshot_params"])


def test_class_attributes_equal_the_reference():
    assert sorted(QuantizerRegistry.list()) == sorted(GOLD["classes"])
    for name, g in GOLD["classes"].items():
        cls = QuantizerRegistry._plugins[name]
        assert cls.__name__ == g["class"] and cls.name == name
        assert [str(x) for x in cls.supported_levels] == g["supported_levels"]
        assert bool(cls.supports_multiple_levels) == g["supports_multiple_levels"]
        assert cls.template_card.title == g["card_title"]
        assert cls.template_card.hyperparameters == g["card_hyperparameters"]


@pytest.mark.parametrize("case", GOLD["llm_compressor"], ids=lambda c: c["id"])
def test_llm_compressor_family_replays_the_reference(case, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    out, cwd = str(tmp_path / "out"), str(tmp_path)
    q = QuantizerRegistry.create(case["method"], model_id=case["model_id"])
    seen = []

    def engine(**kw):
        seen.append(kw)
        return "MODEL"
    call = _untag(case["call"], out, cwd)
    with patch.object(q, "_oneshot", side_effect=engine):
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                q.quantize(**call)
            assert type(ei.value).__name__ == case["raises"] and str(ei.value) == case["message"]
            assert bool(seen) == case["oneshot_called"]
            return
        ret = q.quantize(**call)
    assert ret == _untag(case["returns"], out, cwd)
    assert str(q.last_output_dir) == _untag(case["last_output_dir"], out, cwd)
    assert os.path.isdir(ret) == case["output_dir_exists"] and q.last_model == case["last_model"]
    got, want = dict(seen[-1]), _untag(case["oneshot_kwargs"], out, cwd)
    _check_recipe(_recipe_as_golden(got.pop("recipe")), want.pop("recipe"))
    assert got == want


@pytest.mark.parametrize("case", GOLD["gguf"], ids=lambda c: c["id"])
def test_gguf_plugin_replays_the_reference(case, tmp_path):
    from quantool_b200.engine import gguf_file
    out, cwd = str(tmp_path / "out"), str(tmp_path)
    q = QuantizerRegistry.create("gguf", model_id=case["model_id"])
    steps = []

    def convert(model_path, out_file, outtype="f16", require_tokenizer=True):
        steps.append(["convert", model_path, out_file, outtype])
        open(out_file, "w").write("gguf")
        return out_file

    def quantize(input_gguf, out_file, ftype, devices=None):
        steps.append(["quantize", input_gguf, out_file, ftype])
        open(out_file, "w").write("gguf")
        return out_file
    call = _untag(case["call"], out, cwd)
    conv = lambda x: QuantType[x[5:]] if isinstance(x, str) and x.startswith("ENUM:") else x
    if "level" in call:
        call["level"] = [conv(x) for x in call["level"]] if isinstance(call["level"], list) else conv(call["level"])
    with patch.object(gguf_file, "convert_hf_to_f16_gguf", convert), patch.object(gguf_file, "quantize_gguf", quantize):
        ret = q.quantize(**call)
    assert ret == _untag(case["returns"], out, cwd) and q.last_gguf == _untag(case["last_gguf"], out, cwd)
    want = []
    for cmd in _untag(case["commands"], out, cwd):      # the reference's child-process command lines
        if cmd[0] == "<PYTHON>":
            assert cmd[1].endswith("convert_hf_to_gguf.py") and cmd[3] == "--outfile" and cmd[5] == "--outtype"
            want.append(["convert", cmd[2], cmd[4], cmd[6]])
        else:
            assert cmd[0].endswith("llama-quantize")
            want.append(["quantize", cmd[1], cmd[2], cmd[3]])
    assert steps == want
    saved = tmp_path / "saved"
    saved.mkdir()
    q.save_pretrained(str(saved))
    assert sorted(f for f in os.listdir(saved) if f.endswith(".gguf")) == case["saved_files"]


def test_gguf_constructor_fails_like_the_reference_when_the_engine_is_missing():
    """Reference: RuntimeError when llama-quantize is not found; here: RuntimeError when the CUDA library is."""
    from quantool_b200 import cabi
    assert GOLD["gguf_ctor_without_engine"]["raises"] == "RuntimeError"
    with patch.object(cabi, "lib", side_effect=OSError("libquantool_b200.so: cannot open shared object file")):
        with pytest.raises(RuntimeError):
            QuantizerRegistry.create("gguf", model_id="m")
