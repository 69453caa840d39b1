"""CLI / args / step-pipeline parity: the reference's YAML schema parses
unchanged and the steps route arguments like ref/src/quantool/entrypoints/cli.py."""
import json
import os
from unittest.mock import patch

import pytest
import torch

from quantool_b200.entrypoints import cli

YAML = """
model_id: "{model}"
tokenizer_name: null
cache_dir: null
use_auth_token: false
revision: null
method: "{method}"
quant_level: {level}
quantization_config:
{qcfg}
dataset_path: {dataset}
sample_size: 6
shuffle: true
dataset_seed: 42
output_path: "{out}"
push_to_hub: false
repo_id: null
private: true
enable_evaluation: false
eval_dataset: null
metrics: ["perplexity"]
seed: 42
verbose: true
report_to: null
log_level: "INFO"
save_logs: true
log_dir: "./logs"
"""


def _cfg(tmp_path, **kw):
    p = tmp_path / "cfg.yaml"
    p.write_text(YAML.format(**kw))
    return str(p)


def test_yaml_schema_of_the_reference_parses(tmp_path):
    p = _cfg(tmp_path, model="m", method="gptq", level='"W8A8"', qcfg='  targets: "Linear"\n  ignore: ["lm_head"]',
             dataset="null", out="o")
    margs, qargs, cargs, evargs, eargs, comargs, largs = cli.parse([p])
    assert qargs.method == "gptq" and qargs.quant_level == "W8A8"
    assert qargs.quantization_config == {"targets": "Linear", "ignore": ["lm_head"]}
    assert cargs.sample_size == 6 and cargs.split == "train" and comargs.seed == 42 and largs.log_level == "INFO"
    bad = tmp_path / "bad.yaml"
    bad.write_text('model_id: "m"\nnot_a_field: 1\n')
    with pytest.raises(ValueError):
        cli.parse([str(bad)])                      # allow_extra_keys=False, as in the reference


def test_validate_and_calibration_errors(tmp_path):
    p = _cfg(tmp_path, model="m", method="nope", level="null", qcfg="  {}", dataset="null", out="o")
    st = dict(zip(("model_args", "quant_args", "calibration_args", "eval", "export_args", "common_args", "logging_args"),
                  cli.parse([p])))
    with pytest.raises(ValueError, match="Quantization method 'nope' not available. Available methods:"):
        cli.validate_args_step(st)
    st["quant_args"].method = "awq"
    st["quant_args"].quant_level = ["W4A16", "W8A16"]
    with pytest.raises(ValueError, match="does not support multiple"):
        cli.validate_args_step(st)
    st["quant_args"].quant_level = "W4A16"
    st["model_path"] = "m"
    with pytest.raises(RuntimeError, match="requires calibration data"):
        cli.quantize_step(st)                      # rewrapped as RuntimeError like cli.py:362-364


def test_quantize_step_routes_config_twice_like_the_reference(tmp_path):
    ds = tmp_path / "calib.jsonl"
    ds.write_text("\n".join(json.dumps({"input_ids": list(range(i, i + 12))}) for i in range(10)))
    p = _cfg(tmp_path, model="m", method="gptq", level='"W4A16"',
             qcfg="  method_kwargs:\n    actorder: group\n  max_seq_length: 8", dataset=f'"{ds}"', out=str(tmp_path / "o"))
    st = dict(zip(("model_args", "quant_args", "calibration_args", "eval", "export_args", "common_args", "logging_args"),
                  cli.parse([p])))
    st["model_path"] = "local/dir"
    seen = {}
    from quantool_b200.methods.llm_compressor.gptq import GPTQ

    def fake(self, model, level=None, **kw):
        seen.update(kw, model=model, level=level)
        return "outdir"
    with patch.object(GPTQ, "quantize", fake):
        st = cli.quantize_step(st)
        assert st["quantized_output"] == "outdir" and seen["model"] == "local/dir" and seen["level"] == "W4A16"
        assert seen["method_kwargs"] == {"actorder": "group"} and seen["max_seq_length"] == 8
        # without `load_in_pipeline` the plugin gets the descriptor and resolves it itself (ref cli.py:330-340)
        assert seen["dataset"] is None and seen["dataset_path"] == str(ds) and "num_calibration_samples" not in seen
        # with it, the rows are loaded, shuffled (dataset_seed) and cut to sample_size here
        st["calibration_args"].load_in_pipeline = True
        seen.clear()
        cli.quantize_step(st)
    assert "dataset_path" not in seen and len(seen["dataset"]) == 6
    assert len(seen["dataset"][0]["input_ids"]) == 12            # dict rows travel as a datasets.Dataset


@pytest.mark.gpu
def test_cli_end_to_end_gptq_and_gguf(tmp_path):
    from safetensors.torch import load_file, save_file
    from quantool_b200.engine import llama
    shape = llama.LlamaShape(256, 512, 2, 4, 2, 512, rope_theta=10000.0, tie_word_embeddings=True)
    sd = llama.random_state_dict(shape, seed=1)
    mdir = tmp_path / "tiny"
    mdir.mkdir()
    save_file(sd, str(mdir / "model.safetensors"), metadata={"format": "pt"})
    json.dump(shape.to_hf_config(), open(mdir / "config.json", "w"))
    from _tiny import write_tiny_tokenizer
    write_tiny_tokenizer(str(mdir))
    ids = torch.randint(0, 512, (10, 64), generator=torch.Generator().manual_seed(0))
    torch.save(ids, tmp_path / "calib.pt")
    out = tmp_path / "out_gptq"
    p = _cfg(tmp_path, model=str(mdir), method="gptq", level='"W4A16"',
             qcfg=f'  output_dir: "{tmp_path / "work"}"', dataset=f'"{tmp_path / "calib.pt"}"', out=str(out))
    cli.main([p])
    assert os.path.exists(out / "model.safetensors") and os.path.exists(out / "README.md")
    cfg = json.load(open(out / "config.json"))
    assert cfg["quantization_config"]["format"] == "pack-quantized"
    assert "model.layers.0.self_attn.q_proj.weight_packed" in load_file(str(out / "model.safetensors"))
    out2 = tmp_path / "out_gguf"
    (tmp_path / "g").mkdir()
    p2 = tmp_path / "g" / "cfg.yaml"
    p2.write_text(YAML.format(model=str(mdir), method="gguf", level='["Q8_0", "Q4_0"]', qcfg="  llama_cpp_path: null",
                              dataset="null", out=str(out2)))
    cli.main([str(p2)])
    assert sorted(f for f in os.listdir(out2) if f.endswith(".gguf")) == ["tiny-Q4_0.gguf", "tiny-Q8_0.gguf"]


# ---- calibration rows: chat templates, text fallback, preprocess_fn (ref base.py:257-345, cli.py:284-323) ----------
TEMPLATE = ("{% for m in messages %}<|{{ m['role'] }}|>{{ m['content'] }}<|end|>{% endfor %}"
            "{% if add_generation_prompt %}<|assistant|>{% endif %}")


def _chat_tokenizer(tmp_path):
    from transformers import AutoTokenizer
    from _tiny import write_tiny_tokenizer
    d = tmp_path / "tok"
    d.mkdir(exist_ok=True)
    write_tiny_tokenizer(str(d))
    tok = AutoTokenizer.from_pretrained(str(d))
    tok.chat_template = TEMPLATE
    return tok


def test_chat_rows_are_rendered_like_the_reference(tmp_path):
    from quantool_b200.methods.llm_compressor.chat import has_chat_template, is_conversational, render_chat_row
    tok = _chat_tokenizer(tmp_path)
    u, a = {"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}
    assert has_chat_template(tok) and not has_chat_template(object())
    assert is_conversational({"messages": [u]}) and not is_conversational({"text": "x"}) and not is_conversational({"prompt": "x"})
    assert render_chat_row({"messages": [u, a]}, tok) == {"text": "<|user|>hi<|end|><|assistant|>yo<|end|>"}
    # a prompt that ends on a user turn gets the generation prompt; the completion is what the template adds after it
    r = render_chat_row({"prompt": [u], "completion": [a]}, tok)
    assert r == {"prompt": "<|user|>hi<|end|><|assistant|>", "completion": "yo<|end|>"}
    r = render_chat_row({"prompt": [u], "chosen": [a], "rejected": [{"role": "assistant", "content": "no"}]}, tok)
    assert r["prompt"] == "<|user|>hi<|end|><|assistant|>" and r["chosen"] == "yo<|end|>" and r["rejected"] == "no<|end|>"
    r = render_chat_row({"chosen": [u, a], "rejected": [u]}, tok)
    assert r == {"chosen": "<|user|>hi<|end|><|assistant|>yo<|end|>", "rejected": "<|user|>hi<|end|>"}
    assert render_chat_row({"prompt": [u], "completion": [a], "label": True}, tok)["label"] is True
    # untouched: plain rows, tokenizers without a template; rejected: unsupported key combinations
    assert render_chat_row({"text": "plain"}, tok) == {"text": "plain"}
    tok.chat_template = None
    assert render_chat_row({"messages": [u]}, tok) == {"messages": [u]}
    tok.chat_template = TEMPLATE
    with pytest.raises(KeyError):
        render_chat_row({"messages": [u], "prompt": [u]}, tok)
    # a prompt ending on a system turn cannot be rendered: the row comes back as it was
    bad = {"prompt": [{"role": "system", "content": "s"}]}
    assert render_chat_row(bad, tok) == bad


def test_prepare_calibration_data_text_column(tmp_path):
    import datasets
    from quantool_b200.methods.llm_compressor.gptq import GPTQ
    q = GPTQ(model_id="org/m")
    tok = _chat_tokenizer(tmp_path)
    u, a = {"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}
    ds = q.prepare_calibration_data([{"messages": [u, a]}, {"messages": [u]}], tokenizer=tok)
    assert ds["text"] == ["<|user|>hi<|end|><|assistant|>yo<|end|>", "<|user|>hi<|end|>"]
    # prompt / completion rows: `text` is copied from the first fallback column (the rendered prompt), as upstream
    ds = q.prepare_calibration_data(datasets.Dataset.from_list([{"prompt": [u], "completion": [a]}]), tokenizer=tok)
    assert ds["text"] == ["<|user|>hi<|end|><|assistant|>"] and ds["completion"] == ["yo<|end|>"]
    # no tokenizer: no rendering, plain fallback column
    ds = q.prepare_calibration_data(datasets.Dataset.from_list([{"completion": "abc"}]))
    assert ds["text"] == ["abc"]
    # every split of a DatasetDict
    dd = datasets.DatasetDict({"train": datasets.Dataset.from_list([{"prompt": "p"}]),
                               "test": datasets.Dataset.from_list([{"text": "t"}])})
    dd = q.prepare_calibration_data(dd)
    assert dd["train"]["text"] == ["p"] and dd["test"]["text"] == ["t"]
    # token ids pass through
    ids = torch.zeros((2, 4), dtype=torch.long)
    assert q.prepare_calibration_data(ids) is ids and q.prepare_calibration_data([[1, 2], [3]]) == [[1, 2], [3]]
    # and the rendered rows tokenize at their own length
    q.last_tokenizer = tok
    rows = q._token_ids({"dataset": q.prepare_calibration_data([{"messages": [u, a]}, {"messages": [u]}], tokenizer=tok),
                         "shuffle_calibration_samples": False}, None)
    assert isinstance(rows, list) and len(rows) == 2 and rows[0].numel() > rows[1].numel() > 0


def upper_text(example, suffix=""):
    return {"text": example["text"].upper() + suffix}


def with_tokenizer(example, tokenizer, suffix=""):
    return {"text": example["text"] + tokenizer.eos_token + suffix}


def test_cli_chat_rows_and_preprocess_fn(tmp_path):
    u, a = {"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}
    ds = tmp_path / "chat.jsonl"
    ds.write_text("\n".join(json.dumps({"messages": [u, a]}) for _ in range(8)))
    p = _cfg(tmp_path, model="m", method="awq", level='"W4A16"', qcfg="  {}", dataset=f'"{ds}"', out=str(tmp_path / "o"))
    names = ("model_args", "quant_args", "calibration_args", "eval", "export_args", "common_args", "logging_args")
    st = dict(zip(names, cli.parse([p])))
    st["model_path"], st["tokenizer"] = "local/dir", _chat_tokenizer(tmp_path)
    st["calibration_args"].load_in_pipeline = True
    st["calibration_args"].shuffle = False
    seen = {}
    from quantool_b200.methods.llm_compressor.awq import AWQ

    def fake(self, model, level=None, **kw):
        seen.clear()
        seen.update(kw)
        return "outdir"
    with patch.object(AWQ, "quantize", fake):
        cli.quantize_step(st)
        assert seen["dataset"]["text"] == ["<|user|>hi<|end|><|assistant|>yo<|end|>"] * 6
        # preprocess_fn = "module.func", calibration_config as keyword arguments, tokenizer injected when asked for
        plain = tmp_path / "plain.jsonl"
        plain.write_text("\n".join(json.dumps({"text": f"row {i}"}) for i in range(8)))
        st["calibration_args"].dataset_path = str(plain)
        st["calibration_args"].shuffle = False
        st["calibration_args"].preprocess_fn = "test_cli.upper_text"
        st["calibration_args"].calibration_config = {"suffix": "!"}
        cli.quantize_step(st)
        assert seen["dataset"]["text"][:2] == ["ROW 0!", "ROW 1!"]
        st["calibration_args"].preprocess_fn = "test_cli.with_tokenizer"
        cli.quantize_step(st)
        assert seen["dataset"]["text"][0] == "row 0<|end_of_text|>!"
        # a failing preprocess step is logged and the rows go on unchanged (reference behaviour)
        st["calibration_args"].preprocess_fn = "test_cli.does_not_exist"
        cli.quantize_step(st)
        assert seen["dataset"]["text"][0] == "row 0"
        # rows with neither text nor token ids
        junk = tmp_path / "junk.jsonl"
        junk.write_text(json.dumps({"foo": 1}))
        st["calibration_args"].dataset_path, st["calibration_args"].preprocess_fn = str(junk), None
        cli.quantize_step(st)
        assert seen["dataset"].column_names == ["foo"]      # passed on as it is; the plugin rejects it (below)
    with pytest.raises(ValueError, match="unsupported calibration dataset"):
        AWQ(model_id="m")._token_ids({"dataset": seen["dataset"]}, None)


def test_chat_rows_equal_the_reference_golden(tmp_path):
    """tests/golden/chat_rows.json holds what the reference's OWN `convert_row` returned for these rows
    (tests/golden/make_chat_golden.py, generated in the build container where /root/reference exists)."""
    from quantool_b200.methods.llm_compressor.chat import has_chat_template, render_chat_row
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chat_rows.json")))
    tok = _chat_tokenizer(tmp_path)
    tok.chat_template = gold["template"]
    assert has_chat_template(tok) == gold["has_chat_template"]["with_template"]
    assert len(gold["cases"]) >= 12
    for case in gold["cases"]:
        assert render_chat_row(dict(case["row"]), tok) == case["rendered"], case["row"]
    with pytest.raises(KeyError):
        assert gold["invalid"]["raises"] == "KeyError"
        render_chat_row(dict(gold["invalid"]["row"]), tok)
    tok.chat_template = None
    assert has_chat_template(tok) == gold["has_chat_template"]["without_template"]
    assert render_chat_row(dict(gold["no_template_row"]["row"]), tok) == gold["no_template_row"]["rendered"]


def test_text_column_and_tokenizer_keys_are_honoured(tmp_path):
    """`text_column` and `tokenizer` are `oneshot` parameters the reference forwards (ref base.py:118-124)."""
    import datasets
    from quantool_b200.methods.llm_compressor.gptq import GPTQ
    tok = _chat_tokenizer(tmp_path)
    ds = datasets.Dataset.from_list([{"body": "the thin thing"}, {"body": "in the inn"}])
    q = GPTQ(model_id="org/m")
    rows = q._token_ids({"dataset": ds, "text_column": "body", "tokenizer": tok, "shuffle_calibration_samples": False}, None)
    assert [r.tolist() for r in rows] == [tok("the thin thing")["input_ids"], tok("in the inn")["input_ids"]]
    assert q.last_tokenizer is tok
    with pytest.raises(ValueError, match="unsupported calibration dataset"):
        GPTQ(model_id="org/m")._token_ids({"dataset": ds, "tokenizer": tok}, None)      # no `text` column


def test_preprocessing_func_runs_before_tokenization(tmp_path):
    import datasets
    from quantool_b200.methods.llm_compressor.gptq import GPTQ
    tok = _chat_tokenizer(tmp_path)
    ds = datasets.Dataset.from_list([{"q": "the", "a": "inn"}])
    rows = GPTQ(model_id="m")._token_ids({"dataset": ds, "tokenizer": tok, "shuffle_calibration_samples": False,
                                          "preprocessing_func": lambda ex: {"text": ex["q"] + " " + ex["a"]}}, None)
    assert rows.tolist() == [tok("the inn")["input_ids"]]
