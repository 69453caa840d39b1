"""Writes tests/golden/cli_traces.json: what the REFERENCE's own CLI steps
(/root/reference/src/quantool/entrypoints/cli.py: validate_args_step :163, quantize_step :192, save_step :369,
model_card_step :435) do for a fixed list of YAML configurations - run here, in the build container, with the
absent third-party packages replaced by recording stubs (the llm-compressor / llama.cpp stubs of
make_plugin_golden.py plus `accelerate`, whose only use is a stray import at cli.py:4).

Recorded per case: the keyword arguments that reach `llmcompressor.oneshot` (datasets as their `text` /
`input_ids` columns, in order) or the llama.cpp command lines, the state keys the step sets, and exceptions
(type and message).  tests/test_cli_golden.py replays the same configurations on quantool_b200's CLI.
Run:  python tests/golden/make_cli_golden.py"""
import hashlib
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import make_plugin_golden as mpg  # noqa: E402

TEMPLATE = ("{% for m in messages %}<|{{ m['role'] }}|>{{ m['content'] }}<|end|>{% endfor %}"
            "{% if add_generation_prompt %}<|assistant|>{% endif %}")
U, A = {"role": "user", "content": "hi"}, {"role": "assistant", "content": "yo"}

# dataset files written next to the YAML; "<DATA>" in a config is replaced by their directory
DATASETS = {
    "text.jsonl": [{"text": f"row number {i}"} for i in range(20)],
    "chat.jsonl": [{"messages": [U, {"role": "assistant", "content": f"answer {i}"}]} for i in range(12)],
    "prompt.jsonl": [{"prompt": f"question {i}", "completion": f"reply {i}"} for i in range(6)],
}

BASE = {"model_id": "org/My-Model", "method": "gptq", "quant_level": "W4A16", "quantization_config": {},
        "output_path": "<OUT>/saved"}
CASES = [
    {"id": "descriptor_dataset_path", "cfg": {"dataset_path": "<DATA>/text.jsonl", "sample_size": 6}},
    {"id": "descriptor_dataset_id", "cfg": {"dataset_id": "org/some-dataset", "sample_size": 64,
                                            "quantization_config": {"max_seq_length": 128, "method_kwargs": {"dampening_frac": 0.02}}}},
    {"id": "descriptor_dataset_id_no_sample_size", "cfg": {"dataset_id": "org/some-dataset", "sample_size": None}},
    {"id": "pipeline_text_rows_shuffled", "cfg": {"dataset_path": "<DATA>/text.jsonl", "load_in_pipeline": True,
                                                  "sample_size": 5, "shuffle": True, "dataset_seed": 7}},
    {"id": "pipeline_text_rows_in_order", "cfg": {"dataset_path": "<DATA>/text.jsonl", "load_in_pipeline": True,
                                                  "sample_size": 4, "shuffle": False}},
    {"id": "pipeline_sample_size_larger_than_dataset", "cfg": {"dataset_path": "<DATA>/prompt.jsonl", "load_in_pipeline": True,
                                                               "sample_size": 100, "shuffle": False}},
    {"id": "pipeline_chat_rows_rendered", "tokenizer": True,
     "cfg": {"method": "awq", "dataset_path": "<DATA>/chat.jsonl", "load_in_pipeline": True, "sample_size": 3, "shuffle": False}},
    {"id": "pipeline_chat_rows_without_tokenizer", "tokenizer": False,
     "cfg": {"dataset_path": "<DATA>/chat.jsonl", "load_in_pipeline": True, "sample_size": 2, "shuffle": False}},
    {"id": "pipeline_preprocess_fn", "cfg": {"dataset_path": "<DATA>/text.jsonl", "load_in_pipeline": True, "sample_size": 3,
                                             "shuffle": False, "preprocess_fn": "cli_golden_fns.shout",
                                             "calibration_config": {"suffix": "!"}}},
    {"id": "pipeline_preprocess_fn_with_tokenizer", "tokenizer": True,
     "cfg": {"dataset_path": "<DATA>/text.jsonl", "load_in_pipeline": True, "sample_size": 2, "shuffle": False,
             "preprocess_fn": "cli_golden_fns.add_eos"}},
    {"id": "pipeline_preprocess_fn_missing", "cfg": {"dataset_path": "<DATA>/text.jsonl", "load_in_pipeline": True, "sample_size": 2,
                                                     "shuffle": False, "preprocess_fn": "cli_golden_fns.nope"}},
    {"id": "smoothquant_config_routing", "cfg": {"method": "smoothquant", "quant_level": "W8A8", "dataset_path": "<DATA>/text.jsonl",
                                                 "quantization_config": {"method_kwargs": {"smoothing_strength": 0.7},
                                                                         "num_calibration_samples": 9, "output_dir": "<OUT>/work"}}},
    {"id": "missing_calibration", "cfg": {}},
    {"id": "invalid_scheme_is_rewrapped", "cfg": {"quant_level": "W3A16", "dataset_path": "<DATA>/text.jsonl"}},
    {"id": "gguf_single_level", "cfg": {"method": "gguf", "quant_level": "Q8_0", "quantization_config": {"llama_cpp_path": "<LLAMA_CPP>"}}},
    {"id": "gguf_levels_and_ignored_calibration", "cfg": {"method": "gguf", "quant_level": ["Q4_K_M", "Q8_0"], "dataset_path": "<DATA>/text.jsonl",
                                                          "quantization_config": {"llama_cpp_path": "<LLAMA_CPP>", "output_dir": "<OUT>/gg"}}},
    {"id": "gguf_no_level", "cfg": {"method": "gguf", "quant_level": None, "quantization_config": {"llama_cpp_path": "<LLAMA_CPP>"}}},
    {"id": "push_to_hub", "cfg": {"dataset_path": "<DATA>/text.jsonl", "push_to_hub": True, "repo_id": "org/My-Model-W4A16", "private": True}},
    {"id": "push_to_hub_without_repo_id_saves_locally", "cfg": {"dataset_path": "<DATA>/text.jsonl", "push_to_hub": True, "repo_id": None}},
    {"id": "gguf_push_to_hub", "cfg": {"method": "gguf", "quant_level": "Q4_0", "quantization_config": {"llama_cpp_path": "<LLAMA_CPP>"},
                                       "push_to_hub": True, "repo_id": "org/My-Model-GGUF", "private": False}},
]
VALIDATE_CASES = [
    {"id": "unknown_method", "cfg": {"method": "nope"}},
    {"id": "list_level_on_single_level_method", "cfg": {"method": "awq", "quant_level": ["W4A16", "W8A16"]}},
    {"id": "list_level_on_gguf", "cfg": {"method": "gguf", "quant_level": ["Q4_0", "Q8_0"]}},
    {"id": "no_level", "cfg": {"method": "gptq", "quant_level": None}},
]

FNS = '''
def shout(example, suffix=""):
    return {"text": example["text"].upper() + suffix}


def add_eos(example, tokenizer):
    return {"text": example["text"] + tokenizer.eos_token}
'''


def write_inputs(work):
    data = os.path.join(work, "data")
    os.makedirs(data, exist_ok=True)
    for fn, rows in DATASETS.items():
        with open(os.path.join(data, fn), "w") as f:
            f.write("\n".join(json.dumps(r) for r in rows))
    with open(os.path.join(work, "cli_golden_fns.py"), "w") as f:
        f.write(FNS)
    if work not in sys.path:
        sys.path.insert(0, work)
    return data


def chat_tokenizer(work):
    from transformers import AutoTokenizer
    from _tiny import write_tiny_tokenizer
    d = os.path.join(work, "tok")
    os.makedirs(d, exist_ok=True)
    write_tiny_tokenizer(d)
    tok = AutoTokenizer.from_pretrained(d)
    tok.chat_template = TEMPLATE
    return tok


def subst(v, m):
    if isinstance(v, str):
        for tag, real in m.items():
            v = v.replace(tag, real)
        return v
    if isinstance(v, list):
        return [subst(x, m) for x in v]
    if isinstance(v, dict):
        return {k: subst(x, m) for k, x in v.items()}
    return v


def dataset_view(v):
    """A datasets.Dataset as the columns calibration uses, in row order."""
    cols = getattr(v, "column_names", None)
    if cols is None or isinstance(v, (str, list)):
        return v
    out = {"__dataset__": True, "num_rows": len(v), "columns": sorted(cols)}
    for c in ("text", "input_ids"):
        if c in cols:
            out[c] = list(v[c])
    return out


def main():
    sys.path.insert(0, "/root/reference/src")
    work = tempfile.mkdtemp(prefix="cli_golden_")
    os.chdir(work)
    import datasets  # noqa: F401  (imported before the accelerate stub exists: both probe for the real package)
    import transformers  # noqa: F401
    from transformers import AutoTokenizer, HfArgumentParser  # noqa: F401
    calls = mpg.install_stubs()
    acc = types.ModuleType("accelerate.commands.config.config_args")
    acc.cache_dir = "/tmp"
    for name in ("accelerate", "accelerate.commands", "accelerate.commands.config"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["accelerate.commands.config.config_args"] = acc
    import quantool.entrypoints.cli as rcli
    import quantool.methods.llama_cpp.llama_cpp as ref_gguf
    from quantool.args import (CalibrationArguments, CommonArguments, EvaluationArguments, ExportArguments,
                               LoggingArguments, ModelArguments, QuantizationArguments)
    data = write_inputs(work)
    lcp = os.path.join(work, "llama.cpp")
    os.makedirs(lcp)
    for fn in ("convert_hf_to_gguf.py", "llama-quantize"):
        open(os.path.join(lcp, fn), "w").write("#!/bin/sh\n")
    cmds = []

    def run_command(logger, cmd, *a, **k):
        cmds.append(list(cmd))
        target = cmd[cmd.index("--outfile") + 1] if "--outfile" in cmd else cmd[2]
        open(target, "w").write("gguf")
    ref_gguf.run_command = run_command
    import quantool.core.helpers.export_mixin as ref_export
    hub = []

    def create_repo(**kw):
        hub.append({"call": "create_repo", "kwargs": kw})
        return "https://huggingface.co/" + kw["repo_id"]

    def upload_folder(**kw):
        kw = dict(kw)
        kw["folder_files"] = sorted(os.listdir(kw.pop("folder_path")))
        hub.append({"call": "upload_folder", "kwargs": kw})
        return "commit-url"
    ref_export.create_repo, ref_export.upload_folder = create_repo, upload_folder
    os.environ.pop("HF_TOKEN", None)
    parser = HfArgumentParser((ModelArguments, QuantizationArguments, CalibrationArguments, EvaluationArguments,
                               ExportArguments, CommonArguments, LoggingArguments))
    names = ("model_args", "quant_args", "calibration_args", "evaluation_args", "export_args", "common_args", "logging_args")
    tok = chat_tokenizer(work)
    gold = {"template": TEMPLATE, "datasets": DATASETS, "fns": FNS, "quantize": [], "validate": [], "readme": {}}

    def state_for(cfg, out):
        full = subst({**BASE, **cfg}, {"<DATA>": data, "<OUT>": out, "<LLAMA_CPP>": lcp})
        return dict(zip(names, parser.parse_dict(full, allow_extra_keys=False)))

    for i, case in enumerate(VALIDATE_CASES):
        rec = {"id": case["id"], "cfg": case["cfg"]}
        try:
            rcli.validate_args_step(state_for(case["cfg"], work))
            rec["ok"] = True
        except Exception as e:
            rec["raises"], rec["message"] = type(e).__name__, str(e)
        gold["validate"].append(rec)

    for i, case in enumerate(CASES):
        out = os.path.join(work, f"case{i}")
        os.makedirs(out)
        tags = [(out, "<OUT>"), (data, "<DATA>"), (lcp, "<LLAMA_CPP>"), (sys.executable, "<PYTHON>"), (work, "<CWD>")]
        st = state_for(case["cfg"], out)
        st["model_path"] = "/models/My-Model"
        st["tokenizer"] = tok if case.get("tokenizer") else None
        del calls[:], cmds[:], hub[:]
        rec = {"id": case["id"], "cfg": case["cfg"], "tokenizer": bool(case.get("tokenizer"))}
        try:
            st = rcli.quantize_step(st)
            rec["state_keys"] = sorted(k for k in st if k not in names and k not in ("model_path", "tokenizer"))
            rec["quantized_output"] = mpg.jsonable(st["quantized_output"], tags)
            if calls:
                kw = {k: dataset_view(v) for k, v in calls[-1].items()}
                rec["oneshot_kwargs"] = mpg.jsonable(kw, tags)
            if cmds:
                rec["commands"] = mpg.jsonable(cmds, tags)
        except Exception as e:
            rec["raises"], rec["message"] = type(e).__name__, mpg.jsonable(str(e), tags)
            gold["quantize"].append(rec)
            continue
        try:
            st = rcli.model_card_step(st)
            st = rcli.save_step(st)
            rec["saved"] = sorted(os.listdir(st["export_args"].output_path))
            rec["hub_calls"] = mpg.jsonable(hub, tags)
            readme = open(os.path.join(st["export_args"].output_path, "README.md")).read()
            rec["readme_sha256"] = hashlib.sha256(readme.encode()).hexdigest()
            gold["readme"].setdefault(st["quant_args"].method, readme)
        except Exception as e:
            rec["save_raises"], rec["save_message"] = type(e).__name__, mpg.jsonable(str(e), tags)
        gold["quantize"].append(rec)
    with open(os.path.join(HERE, "cli_traces.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold["quantize"]), "+", len(gold["validate"]), "traces")


if __name__ == "__main__":
    main()
