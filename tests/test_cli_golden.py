"""CLI parity pinned against the reference ITSELF: tests/golden/cli_traces.json holds what the
reference's own `validate_args_step` / `quantize_step` / `model_card_step` / `save_step`
(ref/src/quantool/entrypoints/cli.py:163-444) did for a list of configurations, run in the build container with
recording stubs for llm-compressor, llama.cpp and accelerate (tests/golden/make_cli_golden.py).  The same
configurations go through quantool_b200's CLI steps with the engine entry points replaced by recorders; what
reaches the engine (calibration rows in order, routed keys, recipe), the state keys, saved files and exceptions
must be the reference's.  Two reference defects are NOT reproduced and are asserted as such below."""
import hashlib
import json
import os
import sys
from unittest.mock import patch

import pytest

from quantool_b200.entrypoints import cli
from quantool_b200.methods.llm_compressor.base import LLMCompressorQuantizer

from test_plugin_golden import _check_recipe, _recipe_as_golden

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cli_traces.json")))
BASE = {"model_id": "org/My-Model", "method": "gptq", "quant_level": "W4A16", "quantization_config": {},
        "output_path": "<OUT>/saved"}
NAMES = ("model_args", "quant_args", "calibration_args", "evaluation_args", "export_args", "common_args", "logging_args")

# the reference fails on these for reasons that are defects of its CLI, not contracts (see the test bodies)
REFERENCE_DEFECTS = {
    "descriptor_dataset_id": "got multiple values for keyword argument 'dataset'",
    "descriptor_dataset_id_no_sample_size": "got multiple values for keyword argument 'dataset'",
    "gguf_levels_and_ignored_calibration": "object.__init__() takes exactly one argument",
}


def _subst(v, m):
    if isinstance(v, str):
        for tag, real in m.items():
            v = v.replace(tag, real)
        return v
    if isinstance(v, list):
        return [_subst(x, m) for x in v]
    if isinstance(v, dict):
        return {k: _subst(x, m) for k, x in v.items()}
    return v


@pytest.fixture
def workdir(tmp_path, monkeypatch):
    data = tmp_path / "data"
    data.mkdir()
    for fn, rows in GOLD["datasets"].items():
        (data / fn).write_text("\n".join(json.dumps(r) for r in rows))
    (tmp_path / "cli_golden_fns.py").write_text(GOLD["fns"])
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.chdir(tmp_path)
    sys.modules.pop("cli_golden_fns", None)
    return tmp_path


def _state(cfg, tags):
    from transformers import HfArgumentParser
    from quantool_b200.args import ALL
    full = _subst({**BASE, **cfg}, tags)
    return dict(zip(NAMES, HfArgumentParser(ALL).parse_dict(full, allow_extra_keys=False)))


def _tokenizer(tmp_path):
    from transformers import AutoTokenizer
    from _tiny import write_tiny_tokenizer
    d = tmp_path / "tok"
    d.mkdir(exist_ok=True)
    write_tiny_tokenizer(str(d))
    tok = AutoTokenizer.from_pretrained(str(d))
    tok.chat_template = GOLD["template"]
    return tok


def _dataset_view(v):
    cols = getattr(v, "column_names", None)
    if cols is None or isinstance(v, (str, list)):
        return v
    out = {"__dataset__": True, "num_rows": len(v), "columns": sorted(cols)}
    for c in ("text", "input_ids"):
        if c in cols:
            out[c] = list(v[c])
    return out


@pytest.mark.parametrize("case", GOLD["validate"], ids=lambda c: c["id"])
def test_validate_args_step_replays_the_reference(case):
    st = _state(case["cfg"], {})
    if "raises" in case:
        with pytest.raises(Exception) as ei:
            cli.validate_args_step(st)
        assert type(ei.value).__name__ == case["raises"]
        head = case["message"].split("Available methods:")[0]       # registration order differs, the set does not
        assert str(ei.value).startswith(head)
        if "Available methods:" in case["message"]:
            assert sorted(eval(str(ei.value).split("Available methods:")[1])) == \
                sorted(eval(case["message"].split("Available methods:")[1]))
        else:
            assert str(ei.value) == case["message"]
    else:
        assert cli.validate_args_step(st) is st


@pytest.mark.parametrize("case", GOLD["quantize"], ids=lambda c: c["id"])
def test_quantize_and_save_steps_replay_the_reference(case, workdir):
    from quantool_b200.engine import gguf_file
    out = workdir / "case"
    out.mkdir()
    tags = {"<DATA>": str(workdir / "data"), "<OUT>": str(out), "<LLAMA_CPP>": str(workdir / "llama.cpp"), "<CWD>": str(workdir)}
    st = _state(case["cfg"], tags)
    st["model_path"] = "/models/My-Model"
    st["tokenizer"] = _tokenizer(workdir) if case["tokenizer"] else None
    seen, steps = [], []

    class Model:
        def save_pretrained(self, dest, save_compressed=False, **_):
            assert save_compressed is True
            open(os.path.join(dest, "model.safetensors"), "w").write("weights")

    def engine(self, **kw):
        seen.append(kw)
        return Model()

    def convert(model_path, out_file, outtype="f16", require_tokenizer=True):
        steps.append(["convert", model_path, out_file, outtype])
        open(out_file, "w").write("gguf")
        return out_file

    def quantize(input_gguf, out_file, ftype, devices=None):
        steps.append(["quantize", input_gguf, out_file, ftype])
        open(out_file, "w").write("gguf")
        return out_file

    import huggingface_hub
    hub = []

    def create_repo(**kw):
        hub.append({"call": "create_repo", "kwargs": kw})
        return "https://huggingface.co/" + kw["repo_id"]

    def upload_folder(**kw):
        kw = dict(kw)
        kw["folder_files"] = sorted(os.listdir(kw.pop("folder_path")))
        hub.append({"call": "upload_folder", "kwargs": kw})
        return "commit-url"
    os.environ.pop("HF_TOKEN", None)
    with patch.object(LLMCompressorQuantizer, "_oneshot", engine), \
            patch.object(huggingface_hub, "create_repo", create_repo), patch.object(huggingface_hub, "upload_folder", upload_folder), \
            patch.object(gguf_file, "convert_hf_to_f16_gguf", convert), patch.object(gguf_file, "quantize_gguf", quantize):
        if case["id"] in REFERENCE_DEFECTS:
            assert REFERENCE_DEFECTS[case["id"]] in case["message"]          # what the reference does: crash
            st = cli.quantize_step(st)                                       # here: the evident intent
            if case["id"].startswith("descriptor_dataset_id"):
                kw = seen[-1]
                assert kw["dataset"] == case["cfg"]["dataset_id"] and "dataset_path" not in kw
                assert kw.get("num_calibration_samples") == case["cfg"]["sample_size"]
            else:
                assert [os.path.basename(p) for p in st["quantized_output"]] == ["My-Model-Q4_K_M.gguf", "My-Model-Q8_0.gguf"]
                assert os.path.dirname(st["quantized_output"][0]) == str(out / "gg")
            return
        if "raises" in case:
            with pytest.raises(Exception) as ei:
                cli.quantize_step(st)
            assert type(ei.value).__name__ == case["raises"] and str(ei.value) == _subst(case["message"], tags)
            return
        st = cli.quantize_step(st)
        assert sorted(k for k in st if k not in NAMES and k not in ("model_path", "tokenizer")) == case["state_keys"]
        want_out = _subst(case["quantized_output"], tags)
        if case["cfg"].get("method") == "gguf":      # default output directory is a fresh mkdtemp on both sides
            assert os.path.basename(st["quantized_output"]) == os.path.basename(want_out)
            assert [s[0] for s in steps] == ["convert" if c[0] == "<PYTHON>" else "quantize" for c in case["commands"]]
            assert [s[-1] for s in steps] == [c[-1] for c in case["commands"]]
        else:
            assert st["quantized_output"] == want_out
            got = {k: _dataset_view(v) for k, v in seen[-1].items()}
            want = _subst(case["oneshot_kwargs"], tags)
            _check_recipe(_recipe_as_golden(got.pop("recipe")), want.pop("recipe"))
            assert got == want
        st = cli.model_card_step(st)
        st = cli.save_step(st)
        assert sorted(os.listdir(st["export_args"].output_path)) == case["saved"]
        assert hub == case["hub_calls"]          # push_to_hub: same create_repo / upload_folder calls, same uploaded files
        # the model card: byte for byte the README the reference wrote (huggingface_hub template + front matter)
        readme = open(os.path.join(st["export_args"].output_path, "README.md")).read()
        assert readme == GOLD["readme"][st["quant_args"].method]
        assert hashlib.sha256(readme.encode()).hexdigest() == case["readme_sha256"]
