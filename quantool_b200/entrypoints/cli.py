"""Console entry point: `python -m quantool_b200.entrypoints.cli cfg.yaml` (or argv flags).

Same step sequence as the reference's CLI (ref/src/quantool/entrypoints/cli.py:22-98, 102-444;
SURVEY.md §3.1): setup_logging -> validate_args -> load_model -> quantize -> generate_readme ->
save_model, driven by a dict threaded through the steps.  The reference's YAML configs
(`ref/test_{gptq,awq,gguf}_config.yaml`) parse unchanged, and the steps are replayed against traces of
the reference's own steps (tests/test_cli_golden.py): both calibration modes (`load_in_pipeline: true` - load,
shuffle, sample, `preprocess_fn`, chat templates here - or a `dataset_path` / `dataset_id` descriptor the
plugin resolves), messages, state keys, model card, save / push rules.  Differences, all forced by this being
an offline engine: the model must resolve to a local directory (or the local HF cache), hub datasets must be in
the local `datasets` cache; additive: a `.pt` file of token ids as `dataset_path`.
"""
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
    """Method availability and level shape (ref cli.py:163-189, same messages)."""
    qargs = state["quant_args"]
    available = QuantizerRegistry.list()
    if qargs.method not in available:
        raise ValueError(f"Quantization method '{qargs.method}' not available. Available methods: {available}")
    cls = QuantizerRegistry._plugins[qargs.method]
    if getattr(qargs, "quant_level", None):
        if isinstance(qargs.quant_level, list):
            if not getattr(cls, "supports_multiple_levels", False):
                raise ValueError(f"Quantization method '{qargs.method}' does not support multiple quantization levels. "
                                 f"Please specify a single level or choose a method that supports multiple levels.")
            logger.info(f"Using multiple quantization levels: {qargs.quant_level}")
        else:
            logger.info(f"Using quantization level: {qargs.quant_level}")
    else:
        logger.warning("No quantization level specified, using method defaults")
    return state


def load_model_step(state):
    """Resolve the model to a local directory (ref cli.py:102-136 downloads; this build is offline: a local
    directory or the local HF cache) and load the tokenizer the calibration preprocessing may need."""
    margs = state["model_args"]
    try:
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
            state["tokenizer"] = AutoTokenizer.from_pretrained(margs.tokenizer_name or path, trust_remote_code=True)
        except Exception as e:
            logger.warning(f"Could not load tokenizer: {e}")
            state["tokenizer"] = None
    except Exception as e:
        raise RuntimeError(f"Model loading failed: {e}") from e
    return state


def _load_calibration_dataset(cargs):
    """`load_in_pipeline: true` (ref cli.py:251-282): the dataset is loaded here - a hub id through
    `datasets.load_dataset(id, config, split=...)`, a local path as JSON / JSON-lines - then shuffled with
    `dataset_seed` and cut to `sample_size`, with the same `datasets` calls as the reference so that the same rows
    arrive in the same order.  Additive: a `.pt` file holding a [n, seq] tensor of token ids."""
    if cargs.dataset_path and str(cargs.dataset_path).endswith(".pt"):
        ids = torch.load(cargs.dataset_path)
        rows = [r for r in torch.as_tensor(ids).long()]
        if cargs.shuffle:
            order = torch.randperm(len(rows), generator=torch.Generator().manual_seed(int(cargs.dataset_seed))).tolist()
            rows = [rows[i] for i in order]
        if cargs.sample_size:
            rows = rows[: int(cargs.sample_size)]
        return [r.tolist() for r in rows]
    from datasets import load_dataset
    if cargs.dataset_id:
        ds = load_dataset(cargs.dataset_id, cargs.dataset_config or None, split=cargs.split,
                          cache_dir=cargs.dataset_cache_dir)
    else:
        ds = load_dataset("json", data_files=cargs.dataset_path, split=cargs.split)
    if cargs.shuffle:
        ds = ds.shuffle(seed=cargs.dataset_seed)
    if cargs.sample_size:
        ds = ds.select(range(min(len(ds), int(cargs.sample_size))))
    return ds


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
    """ref cli.py:192-364.  Calibration data reaches the plugin in one of the reference's two ways: loaded and
    prepared here (`load_in_pipeline: true` -> `dataset=<rows>`), or as a descriptor the plugin resolves itself
    (`dataset_path=<path>`, or `dataset=<hub id>` + `num_calibration_samples=sample_size`)."""
    qargs, margs = state["quant_args"], state["model_args"]
    source = state.get("model_path", margs.model_id)
    try:
        quantizer = QuantizerRegistry.create(qargs.method, model_id=margs.model_id, **qargs.quantization_config)
        cargs = state.get("calibration_args")
        try:
            requires = bool(quantizer.require_calibration())
        except Exception:
            requires = False
        has_desc = bool(cargs and (getattr(cargs, "dataset_id", None) or getattr(cargs, "dataset_path", None)))
        if requires and not has_desc:
            raise ValueError(f"Quantization method '{qargs.method}' requires calibration data, but none was provided. "
                             f"Specify 'dataset_id' or 'dataset_path' in calibration_args.")
        if not requires and has_desc:
            logger.warning(f"Quantization method '{qargs.method}' does not require calibration data, but "
                           f"calibration_args were provided. They may be ignored by the method.")
        extra = {}
        dataset = None
        if has_desc and getattr(cargs, "load_in_pipeline", False):
            tok = state.get("tokenizer")
            dataset = _load_calibration_dataset(cargs)
            if cargs.preprocess_fn:
                dataset = _apply_preprocess_fn(dataset, cargs, tok)
            try:
                # chat-template rendering + `text` column (ref cli.py:315-328); token-id rows pass through unchanged
                dataset = (quantizer.prepare_calibration_data(dataset, tokenizer=tok) if tok is not None
                           else quantizer.prepare_calibration_data(dataset))
            except Exception as e:
                logger.warning(f"Quantizer-specific calibration preparation failed: {e}")
        elif has_desc:
            if getattr(cargs, "dataset_id", None):
                # the reference also sets this key and then passes `dataset=None` explicitly, which makes its own
                # call fail ("got multiple values for keyword argument 'dataset'"); the evident intent is kept
                dataset = cargs.dataset_id
                if getattr(cargs, "sample_size", None):
                    extra["num_calibration_samples"] = cargs.sample_size
            else:
                extra["dataset_path"] = cargs.dataset_path
        kwargs = dict(qargs.quantization_config or {})
        kwargs.update(extra)
        state["quantized_output"] = quantizer.quantize(model=source, level=qargs.quant_level, dataset=dataset, **kwargs)
        state["quantizer"] = quantizer
    except Exception as e:
        logger.error(f"Quantization failed: {e}")
        raise RuntimeError(f"Failed to quantize model using {qargs.method}: {e}") from e
    return state


def model_card_step(state):
    """README from the quantizer's template card; a failure is logged, not fatal (ref cli.py:435-444)."""
    try:
        state["quantizer"].save_model_card(save_directory=state["export_args"].output_path)
    except Exception as e:
        logger.warning(f"Failed to check README generation capability: {e}")
    return state


def save_step(state):
    """ref cli.py:369-432: push when `push_to_hub` and `repo_id` are both set, else save locally."""
    eargs = state["export_args"]
    q = state["quantizer"]
    try:
        if getattr(eargs, "push_to_hub", False) and getattr(eargs, "repo_id", None):
            q.push_to_hub(repo_id=eargs.repo_id, private=getattr(eargs, "private", None),
                          commit_message=f"Upload quantized model using {state['quant_args'].method}")
        else:
            output_path = eargs.output_path
            if not output_path:
                output_path = f"./output/{state['quant_args'].method}_{state['model_args'].model_id.replace('/', '_')}"
            q.save_pretrained(save_directory=output_path)
    except Exception as e:
        logger.error(f"Failed to save model: {e}")
        raise RuntimeError(f"Model saving failed: {e}") from e
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
