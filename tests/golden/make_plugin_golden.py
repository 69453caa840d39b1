"""Writes tests/golden/plugin_traces.json: what the REFERENCE's own plugin classes do at the drop-in boundary
 for a fixed list of calls - run here, in the build container, from /root/reference/src with the
absent third-party engines replaced by recording stubs:

  * `llmcompressor.oneshot` -> records the keyword arguments it receives (its signature lists upstream's
    parameter names, which the reference discovers with inspect.signature, ref base.py:46-72);
    `GPTQModifier` / `AWQModifier` / `SmoothQuantModifier` -> record their constructor arguments;
  * `llama-quantize` / `convert_hf_to_gguf.py` -> dummy files in a temporary llama.cpp directory; `run_command`
    (ref llama_cpp.py:12) records each command line and creates the output file.

tests/test_plugin_golden.py replays the same calls on quantool_b200's plugins and compares.
Run:  python tests/golden/make_plugin_golden.py      (the reference is not on the GPU box)"""
import copy
import inspect
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))

# upstream llmcompressor.oneshot's parameter names (llm-compressor 0.8.x; the same list as
# quantool_b200/methods/llm_compressor/base.py::ONESHOT_PARAMS)
ONESHOT_PARAMS = [
    "model", "tokenizer", "processor", "recipe", "recipe_args", "dataset", "dataset_path", "dataset_config_name",
    "splits", "num_calibration_samples", "shuffle_calibration_samples", "max_seq_length", "batch_size", "pad_to_max_length",
    "text_column", "concatenate_data", "streaming", "preprocessing_func", "data_collator", "output_dir",
    "save_compressed", "trust_remote_code_model", "precision", "pipeline", "sequential_targets", "calibration_dataloader",
    "clear_sparse_session", "log_dir", "stage", "oneshot_device", "model_revision", "tie_word_embeddings",
]

TOK = [[1, 2, 3, 4]]
# "<OUT>" is replaced by a per-case temporary directory; an absent output_dir exercises the default path
LLM_CASES = [
    {"id": "gptq_top_level_and_prefixed_keys", "method": "gptq",
     "call": {"model": "some/path", "level": "W4A16", "dataset": TOK, "num_calibration_samples": 2, "max_seq_length": 8,
              "output_dir": "<OUT>", "method_kwargs__dampening_frac": 0.05, "targets": "top-level-is-dropped",
              "llama_cpp_path": "not-an-oneshot-key"}},
    {"id": "gptq_method_kwargs_passthrough", "method": "gptq",
     "call": {"model": "p", "level": None, "dataset": TOK, "output_dir": "<OUT>",
              "method_kwargs": {"scheme": "W8A8", "block_size": 128, "sequential_targets": ["LlamaDecoderLayer"],
                                "ignore": ["lm_head", "re:.*gate_proj"], "targets": "Linear", "unknown_key": 1}}},
    {"id": "gptq_default_output_dir", "method": "gptq", "model_id": "org/My-Model",
     "call": {"model": "p", "level": "W8A16", "dataset": TOK}},
    {"id": "gptq_default_output_dir_no_level", "method": "gptq", "model_id": "org/My-Model",
     "call": {"model": "p", "dataset": TOK}},
    {"id": "gptq_oneshot_kwargs_win_over_top_level", "method": "gptq",
     "call": {"model": "p", "level": "W4A16", "dataset_path": "calib.json", "output_dir": "<OUT>", "max_seq_length": 16,
              "oneshot_kwargs": {"max_seq_length": 32, "save_compressed": False, "model": "other/model"}}},
    {"id": "gptq_dataset_argument_overrides_oneshot_kwargs", "method": "gptq",
     "call": {"model": "p", "level": "W4A16", "dataset": TOK, "output_dir": "<OUT>",
              "oneshot_kwargs": {"dataset": "named-dataset"}}},
    {"id": "gptq_unsupported_but_valid_preset", "method": "gptq",
     "call": {"model": "p", "level": "FP8", "dataset": TOK, "output_dir": "<OUT>"}},
    {"id": "gptq_no_calibration", "method": "gptq", "call": {"model": "p", "level": "W4A16", "output_dir": "<OUT>"}},
    {"id": "gptq_empty_dataset_is_no_calibration", "method": "gptq",
     "call": {"model": "p", "level": "W4A16", "dataset": [], "output_dir": "<OUT>"}},
    {"id": "gptq_invalid_scheme", "method": "gptq", "call": {"model": "p", "level": "W3A16", "dataset": TOK, "output_dir": "<OUT>"}},
    {"id": "awq_default", "method": "awq", "call": {"model": "p", "dataset": TOK, "output_dir": "<OUT>"}},
    {"id": "awq_asym_with_ignore", "method": "awq",
     "call": {"model": "p", "level": "W4A16_ASYM", "dataset": TOK, "output_dir": "<OUT>",
              "method_kwargs": {"ignore": ["lm_head"], "targets": "Linear", "block_size": 64}}},
    {"id": "awq_invalid_scheme", "method": "awq", "call": {"model": "p", "level": "int4", "dataset": TOK, "output_dir": "<OUT>"}},
    {"id": "smoothquant_default", "method": "smoothquant", "call": {"model": "p", "dataset": TOK, "output_dir": "<OUT>"}},
    {"id": "smoothquant_strength_and_level", "method": "smoothquant",
     "call": {"model": "p", "level": "INT8", "dataset": TOK, "output_dir": "<OUT>",
              "method_kwargs__smoothing_strength": 0.8, "method_kwargs": {"ignore": ["lm_head"], "dampening_frac": 0.1}}},
    {"id": "smoothquant_default_output_dir", "method": "smoothquant", "model_id": "m", "call": {"model": "p", "dataset": TOK}},
]

GGUF_CASES = [
    {"id": "gguf_default_level", "model_id": "org/My-Model", "call": {"model": "hf/dir", "output_dir": "<OUT>"}},
    {"id": "gguf_q8_0_goes_through_f16", "model_id": "org/My-Model", "call": {"model": "hf/dir", "level": "Q8_0", "output_dir": "<OUT>"}},
    {"id": "gguf_enum_member", "model_id": "My-Model", "call": {"model": "hf/dir", "level": "ENUM:Q5_K_M", "output_dir": "<OUT>"}},
    {"id": "gguf_f16_direct", "model_id": "org/My-Model", "call": {"model": "hf/dir", "level": "f16", "output_dir": "<OUT>"}},
    {"id": "gguf_F16_by_name", "model_id": "org/My-Model", "call": {"model": "hf/dir", "level": "F16", "output_dir": "<OUT>"}},
    {"id": "gguf_invalid_level_defaults", "model_id": "org/My-Model", "call": {"model": "hf/dir", "level": "Q9_X", "output_dir": "<OUT>"}},
    {"id": "gguf_lowercase_level_is_invalid", "model_id": "org/My-Model", "call": {"model": "hf/dir", "level": "q4_k_m", "output_dir": "<OUT>"}},
    {"id": "gguf_multiple_levels", "model_id": "org/My-Model",
     "call": {"model": "hf/dir", "level": ["Q4_K_M", "Q8_0", "ENUM:Q3_K_S", "bogus"], "output_dir": "<OUT>"}},
    {"id": "gguf_multiple_levels_with_f32", "model_id": "org/My-Model",
     "call": {"model": "hf/dir", "level": ["f32", "Q4_0"], "output_dir": "<OUT>"}},
    {"id": "gguf_extra_kwargs_ignored", "model_id": "org/My-Model",
     "call": {"model": "hf/dir", "level": "Q6_K", "output_dir": "<OUT>", "dataset": TOK, "num_calibration_samples": 4}},
]


def install_stubs():
    import loguru
    calls = []

    class RecordedModel:
        """What `oneshot` returns: the plugins only ever call `save_pretrained(dest, save_compressed=True)` on it."""

        def save_pretrained(self, dest, save_compressed=False, **_):
            assert save_compressed is True
            open(os.path.join(dest, "model.safetensors"), "w").write("weights")

    def oneshot(**kw):
        calls.append(kw)
        return RecordedModel()
    oneshot.__signature__ = inspect.Signature([inspect.Parameter(n, inspect.Parameter.KEYWORD_ONLY, default=None)
                                               for n in ONESHOT_PARAMS])

    class _Recorded:
        def __init__(self, **kw):
            self.kw = kw
    loguru.logger.remove()
    llm = types.ModuleType("llmcompressor")
    llm.oneshot = oneshot
    llm.logger = copy.deepcopy(loguru.logger)
    mq = types.ModuleType("llmcompressor.modifiers.quantization")
    mq.GPTQModifier = type("GPTQModifier", (_Recorded,), {})
    ma = types.ModuleType("llmcompressor.modifiers.awq")
    ma.AWQModifier = type("AWQModifier", (_Recorded,), {})
    ms = types.ModuleType("llmcompressor.modifiers.smoothquant")
    ms.SmoothQuantModifier = type("SmoothQuantModifier", (_Recorded,), {})
    for name, mod in (("llmcompressor", llm), ("llmcompressor.modifiers", types.ModuleType("llmcompressor.modifiers")),
                      ("llmcompressor.modifiers.quantization", mq), ("llmcompressor.modifiers.awq", ma),
                      ("llmcompressor.modifiers.smoothquant", ms)):
        sys.modules[name] = mod
    return calls


def jsonable(v, out_dir):
    if isinstance(v, dict):
        return {k: jsonable(x, out_dir) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [jsonable(x, out_dir) for x in v]
    if hasattr(v, "kw"):
        return {"modifier": type(v).__name__, "kwargs": jsonable(v.kw, out_dir)}
    if type(v).__name__ == "RecordedModel":
        return "MODEL"
    if isinstance(v, (str, os.PathLike)):
        s = os.fspath(v)
        for real, tag in out_dir:
            s = s.replace(real, tag)
        return s
    if v is None or isinstance(v, (bool, int, float)):
        return v
    return repr(v)


def materialise(call, out):
    c = {}
    for k, v in call.items():
        if v == "<OUT>":
            v = out
        c[k] = v
    return c


def main():
    sys.path.insert(0, "/root/reference/src")
    work = tempfile.mkdtemp(prefix="plugin_golden_")
    os.chdir(work)                                   # the reference writes ./logs and the default ./output here
    calls = install_stubs()
    import quantool.methods  # noqa: F401
    from quantool.core.registry import QuantizerRegistry
    assert sorted(QuantizerRegistry.list()) == ["awq", "gguf", "gptq", "smoothquant"], QuantizerRegistry.list()
    gold = {"oneshot_params": ONESHOT_PARAMS, "llm_compressor": [], "gguf": [], "classes": {}}
    for name in QuantizerRegistry.list():
        cls = QuantizerRegistry._plugins[name]
        gold["classes"][name] = {"class": cls.__name__, "supported_levels": [str(x) for x in cls.supported_levels],
                                 "supports_multiple_levels": bool(cls.supports_multiple_levels),
                                 "card_title": cls.template_card.title,
                                 "card_hyperparameters": cls.template_card.hyperparameters}

    for i, case in enumerate(LLM_CASES):
        out = os.path.join(work, f"out{i}")
        tags = [(out, "<OUT>"), (work, "<CWD>")]
        q = QuantizerRegistry.create(case["method"], model_id=case.get("model_id", "org/model"))
        del calls[:]
        rec = {"id": case["id"], "method": case["method"], "model_id": case.get("model_id", "org/model"), "call": case["call"]}
        try:
            ret = q.quantize(**materialise(case["call"], out))
            rec["returns"] = jsonable(ret, tags)
            rec["oneshot_kwargs"] = jsonable(calls[-1], tags)
            rec["last_output_dir"] = jsonable(str(q.last_output_dir), tags)
            rec["output_dir_exists"] = os.path.isdir(ret)
            rec["last_model"] = jsonable(q.last_model, tags)
        except Exception as e:
            rec["raises"] = type(e).__name__
            rec["message"] = str(e)
            rec["oneshot_called"] = bool(calls)
        gold["llm_compressor"].append(rec)

    # ---- gguf: a fake llama.cpp directory; run_command records and creates the output file ----
    import quantool.methods.llama_cpp.llama_cpp as ref_gguf
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
    for i, case in enumerate(GGUF_CASES):
        out = os.path.join(work, f"gg{i}")
        tags = [(out, "<OUT>"), (lcp, "<LLAMA_CPP>"), (sys.executable, "<PYTHON>"), (work, "<CWD>")]
        q = QuantizerRegistry.create("gguf", model_id=case["model_id"], llama_cpp_path=lcp)
        call = materialise(case["call"], out)
        lv = call.get("level")
        conv = lambda x: ref_gguf.QuantType[x[5:]] if isinstance(x, str) and x.startswith("ENUM:") else x
        if isinstance(lv, list):
            call["level"] = [conv(x) for x in lv]
        elif lv is not None:
            call["level"] = conv(lv)
        del cmds[:]
        rec = {"id": case["id"], "model_id": case["model_id"], "call": case["call"]}
        try:
            ret = q.quantize(**call)
            rec["returns"] = jsonable(ret, tags)
            rec["last_gguf"] = jsonable(q.last_gguf, tags)
            rec["commands"] = jsonable(cmds, tags)
            saved = os.path.join(work, f"saved{i}")
            os.makedirs(saved)
            q.save_pretrained(saved)
            rec["saved_files"] = sorted(f for f in os.listdir(saved) if f.endswith(".gguf"))
        except Exception as e:
            rec["raises"] = type(e).__name__
            rec["message"] = str(e)
        gold["gguf"].append(rec)
    # without llama.cpp the reference's constructor fails (SURVEY 8b error convention)
    try:
        QuantizerRegistry.create("gguf", model_id="m")
        gold["gguf_ctor_without_engine"] = None
    except Exception as e:
        gold["gguf_ctor_without_engine"] = {"raises": type(e).__name__, "message": str(e)}
    with open(os.path.join(HERE, "plugin_traces.json"), "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
    print("wrote", len(gold["llm_compressor"]), "+", len(gold["gguf"]), "traces")


if __name__ == "__main__":
    main()
