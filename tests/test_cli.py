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
    with pytest.raises(ValueError, match="Unknown quantization method"):
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
    assert st["quantized_artifact"] == "outdir" and seen["model"] == "local/dir" and seen["level"] == "W4A16"
    assert seen["method_kwargs"] == {"actorder": "group"} and seen["max_seq_length"] == 8
    assert seen["num_calibration_samples"] == 6 and len(seen["dataset"]) == 6 and len(seen["dataset"][0]) == 12


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
