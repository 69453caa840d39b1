"""Shared plugin logic of the calibration-based methods (gptq / awq / smoothquant).

Drop-in for ref/src/quantool/methods/llm_compressor/base.py: same class name, `quantize()`
signature, kwargs routing (`oneshot_kwargs`, `method_kwargs`, `method_kwargs__<name>`, top-level
keys that are `oneshot` parameters), default output directory, calibration-presence check and
`_save_model_files` behaviour.  Where the reference calls `llmcompressor.oneshot(**kwargs)`
(base.py:159-161) this runs the sm_100a engine in quantool_b200.engine.pipeline.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Dict, List, Optional, Tuple, Union

import torch

from ...core.base import BaseQuantizer
from ...core.logger import LoggerFactory

logger = LoggerFactory.get_logger(__name__)

RecipeType = Union[Any, List[Any]]

# parameter names of llmcompressor.oneshot that the reference forwards when they appear as
# top-level keys (ref base.py:118-124 discovers them with inspect.signature at run time)
ONESHOT_PARAMS = {
    "model", "tokenizer", "processor", "recipe", "recipe_args", "dataset", "dataset_path", "dataset_config_name",
    "splits", "num_calibration_samples", "shuffle_calibration_samples", "max_seq_length", "batch_size", "pad_to_max_length",
    "text_column", "concatenate_data", "streaming", "preprocessing_func", "data_collator", "output_dir",
    "save_compressed", "trust_remote_code_model", "precision", "pipeline", "sequential_targets", "calibration_dataloader",
    "clear_sparse_session", "log_dir", "stage", "oneshot_device", "model_revision", "tie_word_embeddings",
}


@dataclass
class Modifier:
    """Recipe entry: the engine-side counterpart of an llm-compressor modifier object."""
    kind: str                                   # "gptq" | "awq" | "smoothquant"
    scheme: Optional[str] = None
    targets: Any = "Linear"
    ignore: List[str] = field(default_factory=lambda: ["lm_head"])
    block_size: int = 128
    dampening_frac: float = 0.01
    actorder: Optional[str] = None              # additive key: None | "group" | "weight"
    sequential_targets: Any = None
    smoothing_strength: float = 0.5
    mappings: Any = None
    n_grid: int = 20
    duo_scaling: bool = True


_MODIFIER_KINDS = {"GPTQModifier": "gptq", "AWQModifier": "awq", "SmoothQuantModifier": "smoothquant"}


def parse_recipe(recipe) -> List[Modifier]:
    """An explicit `recipe=` (ref base.py:81,132-135 hands it to `oneshot` untouched) as a list of `Modifier`s.
    Accepted: `Modifier` objects (one or a list), and llm-compressor's YAML recipe layout - a path, a YAML string or
    the parsed dict `{<stage>: {<group>_modifiers: {GPTQModifier | AWQModifier | SmoothQuantModifier: {fields}}}}`,
    which is also what this engine writes as `recipe.yaml`.  Unknown modifiers or fields raise: nothing is dropped."""
    if isinstance(recipe, Modifier):
        return [recipe]
    if isinstance(recipe, (list, tuple)) and all(isinstance(m, Modifier) for m in recipe):
        return list(recipe)
    if isinstance(recipe, (str, os.PathLike)):
        import yaml
        text = os.fspath(recipe)
        if os.path.exists(text):
            with open(text) as f:
                text = f.read()
        recipe = yaml.safe_load(text)
    if not isinstance(recipe, dict):
        raise TypeError("recipe must be quantool_b200 Modifier objects or an llm-compressor style YAML recipe "
                        "(llm-compressor modifier instances cannot be interpreted without llm-compressor)")
    known = {f.name for f in Modifier.__dataclass_fields__.values()} - {"kind"}
    mods: List[Modifier] = []
    for stage, groups in recipe.items():
        if not isinstance(groups, dict):
            raise ValueError(f"recipe stage {stage!r} holds no modifier groups")
        for group, entries in groups.items():
            for cls, fields in (entries or {}).items():
                if cls not in _MODIFIER_KINDS:
                    raise ValueError(f"recipe modifier {cls!r} is not implemented ({sorted(_MODIFIER_KINDS)} are)")
                fields = dict(fields or {})
                extra = sorted(set(fields) - known)
                if extra:
                    raise ValueError(f"{cls}: unsupported recipe fields {extra}")
                if isinstance(fields.get("targets"), list) and len(fields["targets"]) == 1:
                    fields["targets"] = fields["targets"][0]
                mods.append(Modifier(kind=_MODIFIER_KINDS[cls], **fields))
    if not mods:
        raise ValueError("the recipe holds no modifiers")
    return mods


class LLMCompressorQuantizer(BaseQuantizer):
    """Shared logic for the calibration-based quantizers."""

    def __init__(self, model_id, *args, **kwargs):
        super().__init__(model_id)
        self.last_output_dir: Optional[Path] = None
        self.last_model = None
        self.last_tokenizer = None
        self.source_model = None
        self._last_recipe: Optional[RecipeType] = None

    @classmethod
    def _get_oneshot_params(cls) -> set:
        return set(ONESHOT_PARAMS)

    def require_calibration(self):
        return True

    def quantize(self, model: Union[str, Path, Any], level: Optional[str] = None, recipe: Optional[RecipeType] = None,
                 oneshot_kwargs: Optional[Dict[str, Any]] = None, method_kwargs: Optional[Dict[str, Any]] = None,
                 dataset: Optional[Any] = None, **kwargs) -> str:
        if isinstance(level, list):
            raise ValueError(f"Method '{self.name}' does not support multiple quantization levels. "
                             f"Please specify a single level.")
        oneshot_kwargs = dict(oneshot_kwargs or {})
        method_kwargs = dict(method_kwargs or {})
        if dataset is not None:
            oneshot_kwargs["dataset"] = dataset
        valid = self._get_oneshot_params()
        for key in list(kwargs.keys()):
            if key in valid:
                oneshot_kwargs.setdefault(key, kwargs.pop(key))
        for key in list(kwargs.keys()):
            if key.startswith("method_kwargs__"):
                method_kwargs[key.split("__", 1)[1]] = kwargs.pop(key)

        if recipe is None:
            recipe, inferred_level = self._build_recipe(level, method_kwargs)
        else:
            inferred_level = level or getattr(self, "default_level", "default")
        self._last_recipe = recipe

        oneshot_kwargs = self._prepare_oneshot_kwargs(model, oneshot_kwargs, inferred_level)
        oneshot_kwargs.setdefault("recipe", recipe)
        if not self._has_calibration_source(oneshot_kwargs):
            raise ValueError("llm-compressor integrations require calibration data. "
                             "Provide `dataset`, `dataset_path`, or a custom `calibration_dataloader` "
                             "through `oneshot_kwargs`.")
        self.logger.info(f"Running sm_100a oneshot with output_dir={oneshot_kwargs.get('output_dir')}")
        self.source_model = model
        try:
            self.last_model = self._oneshot(**oneshot_kwargs)
        except Exception as e:
            self.logger.error(f"oneshot failed: {e}")
            raise e
        self.last_output_dir = Path(oneshot_kwargs["output_dir"]).resolve()
        self.logger.info(f"Quantization complete. Model ready at: {self.last_output_dir}")
        return str(self.last_output_dir)

    # ExportMixin hook (ref base.py:175-207)
    def _save_model_files(self, save_directory: Union[str, Path]):
        if not self.last_model:
            raise RuntimeError("No quantized model available. Call `quantize()` before saving.")
        dest = Path(save_directory)
        dest.mkdir(parents=True, exist_ok=True)
        self.last_model.save_pretrained(str(dest), save_compressed=True)
        if self.last_tokenizer is not None:
            self.last_tokenizer.save_pretrained(str(dest))

    # ------------------------------------------------------------------
    def _build_recipe(self, level: Optional[str], method_kwargs: Dict[str, Any]) -> Tuple[RecipeType, str]:  # pragma: no cover
        raise NotImplementedError

    def _default_output_dir(self, level_hint: Optional[str]) -> Path:
        model_name = str(self.model_id).replace("/", "_") if self.model_id else "model"
        level_fragment = (level_hint or "default").replace("/", "_")
        return Path("./output") / f"{self.name}_{model_name}_{level_fragment}"

    def _prepare_oneshot_kwargs(self, model, oneshot_kwargs: Dict[str, Any], level_hint: Optional[str]) -> Dict[str, Any]:
        prepared = dict(oneshot_kwargs)
        prepared.setdefault("model", model)
        prepared.setdefault("save_compressed", True)
        prepared.setdefault("trust_remote_code_model", True)
        output_dir = prepared.get("output_dir") or self._default_output_dir(level_hint)
        prepared["output_dir"] = str(output_dir)
        Path(prepared["output_dir"]).mkdir(parents=True, exist_ok=True)
        return prepared

    def _has_calibration_source(self, oneshot_kwargs: Dict[str, Any]) -> bool:
        for key in ("dataset", "dataset_path", "calibration_dataloader"):
            v = oneshot_kwargs.get(key)
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                if v.numel() > 0:
                    return True
            elif v is not False and (not hasattr(v, "__len__") or len(v) > 0):
                return True
        return False

    def prepare_calibration_data(self, dataset, tokenizer=None):
        """Calibration rows -> something `quantize(dataset=...)` can tokenize (ref base.py:257-345): with a
        tokenizer, conversational rows are rendered through its chat template (`chat.render_chat_row`); then every
        split gets a `text` column, copied from the first of prompt / completion / chosen / rejected / label when it
        has none.  Token-id tensors and lists of token-id rows pass through; a list of dict rows becomes a
        `datasets.Dataset` first."""
        from .chat import render_chat_row
        if isinstance(dataset, torch.Tensor):
            return dataset
        if isinstance(dataset, (list, tuple)):
            if not dataset or not isinstance(dataset[0], dict):
                return dataset
            import datasets
            dataset = datasets.Dataset.from_list([dict(r) for r in dataset])
        if tokenizer is not None:
            try:
                dataset = dataset.map(lambda ex: render_chat_row(ex, tokenizer), batched=False)
                self.logger.info("Applied chat template processing to calibration dataset")
            except Exception as e:
                self.logger.warning(f"Failed to apply chat template processing: {e}")
        try:
            if isinstance(dataset, dict):                      # DatasetDict: every split
                for split in list(dataset.keys()):
                    dataset[split] = self._ensure_text_column(dataset[split])
            else:
                dataset = self._ensure_text_column(dataset)
        except Exception as e:  # pragma: no cover
            self.logger.warning(f"Error while ensuring text column for calibration dataset: {e}")
        return dataset

    def _ensure_text_column(self, ds):
        cols = set(getattr(ds, "column_names", []) or [])
        if "text" in cols or "text_target" in cols or "input_ids" in cols:
            return ds
        for c in ("prompt", "completion", "chosen", "rejected", "label"):
            if c in cols:
                try:
                    ds = ds.map(lambda ex, _c=c: {"text": ex.get(_c)}, batched=False)
                    self.logger.info(f"Created 'text' column from fallback '{c}'")
                except Exception as e:  # pragma: no cover
                    self.logger.warning(f"Failed to create 'text' fallback column from '{c}': {e}")
                break
        return ds

    # ------------------------------------------------------------------
    # engine
    # ------------------------------------------------------------------
    def _load_model(self, model):
        """-> (hf_config dict, host state dict, source dir or None).  `model` is a local HF directory
        (the reference passes the resolved path, cli.py:345) or an in-memory (config, state_dict) pair."""
        from ...engine import gguf_file
        if isinstance(model, (tuple, list)) and len(model) == 2:
            cfg, sd = model
            cfg = cfg.to_hf_config() if hasattr(cfg, "to_hf_config") else dict(cfg)
            return cfg, sd, None
        path = getattr(model, "name_or_path", str(model))
        if not os.path.isdir(path):
            raise FileNotFoundError(f"model path {path!r} is not a local directory; resolve/download the model first "
                                    f"(the reference's load_model_step does this before calling quantize)")
        cfg, sd = gguf_file.load_hf_model(path)
        return cfg, sd, path

    def _token_ids(self, oneshot_kwargs: Dict[str, Any], source_dir: Optional[str]):
        """Calibration samples as token ids: a [n, seq] tensor when every sample has the same length, else a list
        of 1-D tensors.  Like llm-compressor's oneshot every sample is calibrated at ITS OWN length (truncated to
        `max_seq_length`, never to the shortest sample), `num_calibration_samples` defaults to 512 and the samples
        are shuffled before the first `num_calibration_samples` are taken (`shuffle_calibration_samples`, default
        True; seeded here so runs repeat)."""
        ds = oneshot_kwargs.get("dataset")
        n = oneshot_kwargs.get("num_calibration_samples") or 512
        max_len = oneshot_kwargs.get("max_seq_length")     # explicit: truncates everything; default: tokenizer only
        tok_len = max_len or 384                           # llm-compressor DatasetArguments default for tokenization
        shuffle = oneshot_kwargs.get("shuffle_calibration_samples", True)
        text_col = oneshot_kwargs.get("text_column") or "text"      # oneshot's `text_column`, `tokenizer` are honoured
        if ds is None and oneshot_kwargs.get("dataset_path"):
            ds = self._load_dataset_path(oneshot_kwargs["dataset_path"], oneshot_kwargs)
        if isinstance(ds, str):
            ds = self._load_dataset_id(ds, oneshot_kwargs)
        if oneshot_kwargs.get("preprocessing_func") is not None and hasattr(ds, "map"):
            ds = ds.map(oneshot_kwargs["preprocessing_func"])       # oneshot's per-row hook, applied before tokenization
        if ds is None:
            raise ValueError("no calibration data: pass `dataset` (token-id tensor, list of token lists, or a "
                             "datasets.Dataset with `input_ids` or `text`) or a local `dataset_path`")
        if isinstance(ds, torch.Tensor):
            rows = list(ds.long().unbind(0))
        else:
            cols = set(getattr(ds, "column_names", []) or [])
            if "input_ids" in cols:
                rows = [list(r) for r in ds["input_ids"]]
            elif text_col in cols:
                tok = oneshot_kwargs.get("tokenizer") or self.last_tokenizer
                if tok is None or isinstance(tok, str):
                    from transformers import AutoTokenizer
                    tok = AutoTokenizer.from_pretrained(tok or source_dir or self.model_id)
                self.last_tokenizer = tok
                rows = [tok(t, truncation=True, max_length=tok_len)["input_ids"] for t in ds[text_col]]
            elif isinstance(ds, (list, tuple)):
                rows = [list(r) for r in ds]
            else:
                raise ValueError("unsupported calibration dataset type")
            rows = [torch.as_tensor(r, dtype=torch.long).reshape(-1) for r in rows]
        empty = sum(1 for r in rows if r.numel() == 0)
        if empty:
            self.logger.warning(f"dropping {empty} empty calibration samples")
        rows = [r[:max_len] if max_len else r for r in rows if r.numel() > 0]
        if not rows:
            raise ValueError("the calibration dataset holds no non-empty samples")
        if shuffle and len(rows) > 1:
            order = torch.randperm(len(rows), generator=torch.Generator().manual_seed(42)).tolist()
            rows = [rows[i] for i in order]
        rows = rows[: int(n)]
        if len({int(r.numel()) for r in rows}) == 1:
            return torch.stack(rows).contiguous()
        return rows

    def _load_dataset_id(self, name: str, kw: Dict[str, Any]):
        """`dataset=<hub id>` (what the reference's CLI passes without `load_in_pipeline`, ref cli.py:334-338, and
        llm-compressor resolves through `datasets`): served from the local `datasets` cache when it is there."""
        import datasets
        split = kw.get("splits") or "train"
        split = split if isinstance(split, str) else next(iter(split.values() if isinstance(split, dict) else split))
        try:
            ds = datasets.load_dataset(name, kw.get("dataset_config_name") or None, split=split)
        except Exception as e:
            raise ValueError(f"calibration dataset {name!r} could not be loaded (hub datasets need network access or "
                             f"a populated local cache; pass `dataset_path` or a loaded dataset instead): {e}") from e
        return self.prepare_calibration_data(ds)

    def _load_dataset_path(self, path: str, kw: Dict[str, Any]):
        """Local calibration files (`dataset_path`; the reference hands this key to oneshot, ref base.py:118-124):
        .pt (token-id tensor), .json / .jsonl / .txt / .parquet / a `datasets.save_to_disk` directory with a `text`
        or `input_ids` column.  Hub dataset names need network access and raise."""
        if not os.path.exists(path):
            raise ValueError(f"dataset_path {path!r} does not exist locally (hub datasets cannot be fetched here)")
        if path.endswith(".pt"):
            return torch.load(path)
        import datasets
        if os.path.isdir(path):
            ds = datasets.load_from_disk(path)
        else:
            ext = path.rsplit(".", 1)[-1].lower()
            kind = {"jsonl": "json", "json": "json", "txt": "text", "parquet": "parquet", "csv": "csv"}.get(ext)
            if kind is None:
                raise ValueError(f"dataset_path: unsupported file type .{ext}")
            ds = datasets.load_dataset(kind, data_files=path)
        if isinstance(ds, dict) or hasattr(ds, "keys"):
            split = (kw.get("splits") or "train")
            split = split if isinstance(split, str) else next(iter(split))
            ds = ds[split if split in ds else next(iter(ds.keys()))]
        return self.prepare_calibration_data(ds)

    def _oneshot(self, model, recipe, output_dir, save_compressed=True, **kw):
        from ...engine import artifacts, llama, pipeline, schemes
        if not torch.cuda.is_available():
            raise RuntimeError("the sm_100a quantization engine needs a CUDA device (there is no CPU fallback)")
        mods = parse_recipe(recipe)
        self._check_supported(mods, kw)
        cfg, sd, src = self._load_model(model)
        shape = llama.LlamaShape.from_hf_config(cfg)
        ids = self._token_ids({**kw, "dataset": kw.get("dataset")}, src)
        quant = [m for m in mods if m.kind in ("gptq", "awq")]
        if len(quant) != 1:
            raise ValueError("a recipe needs exactly one quantizing modifier (gptq or awq)")
        q = quant[0]
        smooth = [m for m in mods if m.kind == "smoothquant"]
        args = schemes.resolve(q.scheme, q.actorder)
        fmt = artifacts.artifact_format(args.num_bits, q.scheme)
        dev = torch.device("cuda", torch.cuda.current_device())
        if q.kind == "gptq":
            if q.block_size != 128:
                raise ValueError("block_size other than 128 is not supported by the sm_100a GPTQ kernel")
            res = pipeline.quantize_model_gptq(shape, sd, ids, args, dev, fmt=fmt, percdamp=q.dampening_frac,
                                               smooth_strength=smooth[0].smoothing_strength if smooth else None,
                                               ignore=tuple(q.ignore))
        else:
            res = pipeline.quantize_model_awq(shape, sd, ids, args, dev, fmt=fmt, n_grid=q.n_grid,
                                              duo_scaling=q.duo_scaling)
        qcfg = artifacts.quantization_config(q.scheme, args.actorder, fmt, ignore=q.ignore)
        d = pipeline.Dist()
        qm = artifacts.QuantizedModel(cfg, res.tensors, qcfg, source_dir=src, writer=(d.rank == 0))
        qm.stats = res
        qm.recipe = list(mods)              # written as recipe.yaml next to the weights, as oneshot does
        if save_compressed:
            qm.save_pretrained(output_dir, save_compressed=True)      # rank 0 writes; the other ranks hold no tensors
        if d.on:
            d.dist.barrier()
        return qm

    @staticmethod
    def _check_supported(mods, kw: Dict[str, Any]) -> None:
        """Every accepted key is either honoured or rejected - nothing is silently dropped (the reference routes
        these keys into the modifier / oneshot, ref base.py:110-130, gptq.py:75-84, awq.py:77-79)."""
        def _as_list(v):
            return [v] if isinstance(v, str) else list(v or [])
        for m in mods:
            if _as_list(m.targets) != ["Linear"]:
                raise ValueError(f"targets={m.targets!r}: the sm_100a engine quantizes the decoder-layer Linears "
                                 f"(targets='Linear'); use `ignore` to exclude modules")
            if m.sequential_targets is not None and _as_list(m.sequential_targets) != ["LlamaDecoderLayer"]:
                raise ValueError(f"sequential_targets={m.sequential_targets!r}: calibration is sequential per "
                                 f"LlamaDecoderLayer (the llm-compressor default for Llama); other values are not supported")
            if m.mappings is not None:
                raise ValueError("custom `mappings` are not supported: the engine uses llm-compressor's Llama default "
                                 "mappings")
            if m.kind in ("awq", "smoothquant") and _as_list(m.ignore) not in ([], ["lm_head"]):
                raise ValueError(f"`ignore` beyond ['lm_head'] is honoured by the gptq pass only (got {m.ignore!r} on "
                                 f"{m.kind})")
        st = kw.get("sequential_targets")
        if st is not None and _as_list(st) != ["LlamaDecoderLayer"]:
            raise ValueError(f"sequential_targets={st!r} is not supported (LlamaDecoderLayer only)")
        for key in ("pipeline",):
            if kw.get(key) not in (None, "sequential", "independent"):
                raise ValueError(f"{key}={kw[key]!r} is not supported")
        if kw.get("calibration_dataloader") is not None:
            raise ValueError("`calibration_dataloader` is not supported: pass `dataset` or `dataset_path`")
        for key in ("data_collator", "recipe_args", "stage", "processor"):
            if kw.get(key) is not None:
                raise ValueError(f"`{key}` is not supported by the sm_100a engine")
        if kw.get("streaming"):
            raise ValueError("`streaming` datasets are not supported: calibration rows are materialised")
        if kw.get("precision") not in (None, "auto"):
            raise ValueError(f"precision={kw['precision']!r} is not supported: the checkpoint's own dtype is used ('auto')")
        if kw.get("pad_to_max_length") or kw.get("concatenate_data"):
            raise ValueError("pad_to_max_length / concatenate_data are not supported: samples are calibrated at their "
                             "own length")
