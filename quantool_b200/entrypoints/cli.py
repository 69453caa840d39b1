"""Console entry point: `python -m quantool_b200.entrypoints.cli cfg.yaml` (or argv flags).

Same step sequence as the reference's CLI (ref/src/quantool/entrypoints/cli.py:22-98, 102-444;
SURVEY.md §3.1): setup_logging -> validate_args -> load_model -> quantize -> generate_readme ->
save_model, driven by a dict threaded through the steps.  The reference's YAML configs
(`ref/test_{gptq,awq,gguf}_config.yaml`) parse unchanged.  Differences, all forced by this being
an offline, single-node engine: the model must resolve to a local directory (or the HF cache) and
calibration data comes from `dataset_path` (a `.pt` tensor of token ids, a `.json`/`.jsonl` file of
`{"input_ids": [...]}` / `{"text": ...}` rows) or from `datasets.load_dataset` when it is reachable.
"""
import json
import logging
import os
import sys
import time

import torch

from .. import methods  # noqa: F401  (registers the plugins)
from ..args import ALL
from ..core import LoggerFactory, QuantizerRegistry

logger = LoggerFactory.get_logger(__name__)


class PipelineBase:
    """Ordered (name, fn) steps over a state dict (ref core/helpers/pipeline.py:13-27)."""

    def __init__(self):
        self.steps = []

    def add_step(self, fn, name=None):
        self.steps.append((name or fn.__name__, fn))
        return self

    def run(self, state):
        for name, fn in self.steps:
            t0 = time.perf_counter()
            logger.info(f"START {name}")
            out = fn(state)
            if isinstance(out, dict):
                state = out
            logger.info(f"END {name} (duration={time.perf_counter() - t0:.3f}s)")
        return state


def setup_logging_step(state):
    largs = state["logging_args"]
    logging.getLogger("quantool_b200").setLevel(getattr(logging, str(largs.log_level).upper(), logging.INFO))
    state["loggers"] = {}
    return state


def validate_args_step(state):
    qargs = state["quant_args"]
    if qargs.method not in QuantizerRegistry.list():
        raise ValueError(f"Unknown quantization method '{qargs.method}'. Available: {QuantizerRegistry.list()}")
    cls = QuantizerRegistry._plugins[qargs.method]
    if isinstance(qargs.quant_level, list) and not cls.supports_multiple_levels:
        raise ValueError(f"Method '{qargs.method}' does not support multiple quantization levels. "
                         f"Please specify a single level.")
    return state


def load_model_step(state):
    margs = state["model_args"]
    path = margs.model_id
    if not os.path.isdir(path):
        try:
            from huggingface_hub import snapshot_download
            path = snapshot_download(margs.model_id, cache_dir=margs.cache_dir, revision=margs.revision,
                                     local_files_only=True)
        except Exception as e:
            raise RuntimeError(f"model '{margs.model_id}' is neither a local directory nor in the local HF cache "
                               f"(this build does not download): {e}")
    state["model_path"] = path
    try:
        from transformers import AutoTokenizer
        state["tokenizer"] = AutoTokenizer.from_pretrained(margs.tokenizer_name or path)
    except Exception as e:
        logger.warning(f"no tokenizer loaded: {e}")
        state["tokenizer"] = None
    return state


def _load_local_dataset(path: str, sample_size, seed: int, shuffle: bool):
    if path.endswith(".pt"):
        ids = torch.load(path)
        rows = [list(map(int, r)) for r in ids]
    else:
        rows = []
        with open(path) as f:
            data = json.load(f) if path.endswith(".json") else [json.loads(l) for l in f if l.strip()]
        for r in data:
            rows.append(r)
    if shuffle:
        g = torch.Generator().manual_seed(seed)
        order = torch.randperm(len(rows), generator=g).tolist()
        rows = [rows[i] for i in order]
    if sample_size:
        rows = rows[: int(sample_size)]
    return rows


def _apply_preprocess_fn(dataset, cargs, tokenizer):
    """`preprocess_fn: "module.func"` is mapped over the rows, with `calibration_config` as keyword arguments and the
    tokenizer as second positional argument when the function has a `tokenizer` parameter (ref cli.py:284-313).  As in
    the reference a failing preprocess step is logged and the rows go on unchanged."""
    import importlib
    import inspect
    try:
        if not hasattr(dataset, "map"):
            raise TypeError("rows of token ids cannot be preprocessed (use a .json / .jsonl dataset)")
        module_name, fn_name = cargs.preprocess_fn.rsplit(".", 1)
        fn = getattr(importlib.import_module(module_name), fn_name)
        kw = dict(cargs.calibration_config or {})
        if "tokenizer" in inspect.signature(fn).parameters:
            if tokenizer is None:
                raise ValueError(f"Preprocessing function '{cargs.preprocess_fn}' requires a tokenizer, but none was loaded")
            return dataset.map(lambda ex: fn(ex, tokenizer, **kw), batched=False)
        return dataset.map(lambda ex: fn(ex, **kw), batched=False)
    except Exception as e:
        logger.warning(f"Failed to run preprocess_fn '{cargs.preprocess_fn}': {e}")
        return dataset


def quantize_step(state):
    qargs, margs = state["quant_args"], state["model_args"]
    source = state.get("model_path", margs.model_id)
    try:
        quantizer = QuantizerRegistry.create(qargs.method, model_id=margs.model_id, **qargs.quantization_config)
        cargs = state.get("calibration_args")
        requires = bool(quantizer.require_calibration())
        has_desc = bool(cargs and (cargs.dataset_id or cargs.dataset_path))
        if requires and not has_desc:
            raise ValueError(f"Quantization method '{qargs.method}' requires calibration data, but none was provided. "
                             f"Specify 'dataset_id' or 'dataset_path' in calibration_args.")
        if not requires and has_desc:
            logger.warning(f"Quantization method '{qargs.method}' does not require calibration data; it is ignored.")
        extra = {}
        dataset = None
        if requires:
            tok = state.get("tokenizer")
            if tok is not None:
                quantizer.last_tokenizer = tok             # the plugin tokenizes `text` rows and saves it with the model
            if cargs.dataset_path:
                dataset = _load_local_dataset(cargs.dataset_path, cargs.sample_size, cargs.dataset_seed, cargs.shuffle)
            else:
                from datasets import load_dataset
                dataset = load_dataset(cargs.dataset_id, cargs.dataset_config or None, split=cargs.split,
                                       cache_dir=cargs.dataset_cache_dir)
                if cargs.shuffle:
                    dataset = dataset.shuffle(seed=cargs.dataset_seed)
                if cargs.sample_size:
                    dataset = dataset.select(range(min(int(cargs.sample_size), len(dataset))))
            if dataset and isinstance(dataset, list) and isinstance(dataset[0], dict):
                import datasets
                dataset = datasets.Dataset.from_list(dataset)
            if cargs.preprocess_fn:
                dataset = _apply_preprocess_fn(dataset, cargs, tok)
            # chat-template rendering + `text` column (ref cli.py:315-323); token-id rows pass through unchanged
            dataset = quantizer.prepare_calibration_data(dataset, tokenizer=tok)
            cols = set(getattr(dataset, "column_names", []) or [])
            if cols and not ({"text", "input_ids"} & cols):
                raise ValueError(f"calibration rows carry neither `text` nor `input_ids` (columns: {sorted(cols)})")
            if "text" in cols and "input_ids" not in cols and tok is None:
                raise ValueError("text calibration rows need a tokenizer")
            extra["shuffle_calibration_samples"] = False     # already shuffled above with dataset_seed
            if cargs.sample_size:
                extra["num_calibration_samples"] = int(cargs.sample_size)
        kwargs = dict(qargs.quantization_config)
        kwargs.update(extra)
        if dataset is not None:
            kwargs["dataset"] = dataset
        state["quantized_artifact"] = quantizer.quantize(model=source, level=qargs.quant_level, **kwargs)
        state["quantizer"] = quantizer
    except Exception as e:
        raise RuntimeError(f"Quantization failed: {e}") from e
    return state


def model_card_step(state):
    state["quantizer"].save_model_card(state["export_args"].output_path)
    return state


def save_step(state):
    eargs = state["export_args"]
    q = state["quantizer"]
    if eargs.push_to_hub:
        q.push_to_hub(repo_id=eargs.repo_id, private=eargs.private,
                      commit_message=f"Upload {state['quant_args'].method} quantized model")
    else:
        q.save_pretrained(eargs.output_path)
    return state


def parse(argv):
    from transformers import HfArgumentParser
    parser = HfArgumentParser(ALL)
    if len(argv) == 1 and argv[0].endswith((".yaml", ".yml")):
        return parser.parse_yaml_file(argv[0], allow_extra_keys=False)
    return parser.parse_args_into_dataclasses(args=argv)


def main(argv=None):
    try:
        margs, qargs, cargs, evargs, eargs, comargs, largs = parse(sys.argv[1:] if argv is None else argv)
        pipeline = (PipelineBase().add_step(setup_logging_step, "setup_logging").add_step(validate_args_step, "validate_args")
                    .add_step(load_model_step, "load_model").add_step(quantize_step, "quantize")
                    .add_step(model_card_step, "generate_readme").add_step(save_step, "save_model"))
        state = {"model_args": margs, "quant_args": qargs, "calibration_args": cargs, "export_args": eargs,
                 "common_args": comargs, "logging_args": largs}
        pipeline.run(state)
    except KeyboardInterrupt:
        sys.exit(1)
    except Exception as e:
        import traceback
        logger.error(f"Pipeline execution failed: {e}")
        logger.error(traceback.format_exc())
        sys.exit(1)


if __name__ == "__main__":
    main()
