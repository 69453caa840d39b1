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
        """Ensure a `text` column exists (ref base.py:257-345); token-id tensors pass through."""
        if isinstance(dataset, torch.Tensor):
            return dataset
        cols = set(getattr(dataset, "column_names", []) or [])
        if "text" in cols or "text_target" in cols or "input_ids" in cols:
            return dataset
        for c in ("prompt", "completion", "chosen", "rejected", "label"):
            if c in cols:
                try:
                    dataset = dataset.map(lambda ex, _c=c: {"text": ex.get(_c)}, batched=False)
                    self.logger.info(f"Created 'text' column from fallback '{c}'")
                except Exception as e:  # pragma: no cover
                    self.logger.warning(f"Failed to create 'text' fallback column from '{c}': {e}")
                break
        return dataset

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

    def _token_ids(self, oneshot_kwargs: Dict[str, Any], source_dir: Optional[str]) -> torch.Tensor:
        ds = oneshot_kwargs.get("dataset")
        n = oneshot_kwargs.get("num_calibration_samples")
        max_len = oneshot_kwargs.get("max_seq_length") or 2048
        if ds is None:
            raise ValueError("this build calibrates from an in-memory `dataset` (token-id tensor, list of token lists, "
                             "or a datasets.Dataset with `input_ids` or `text`); `dataset_path` loading needs the "
                             "reference's dataset loader")
        if isinstance(ds, torch.Tensor):
            ids = ds.long()
        else:
            rows = None
            cols = set(getattr(ds, "column_names", []) or [])
            if "input_ids" in cols:
                rows = [list(r) for r in ds["input_ids"]]
            elif "text" in cols:
                from transformers import AutoTokenizer
                tok = self.last_tokenizer or AutoTokenizer.from_pretrained(source_dir or self.model_id)
                self.last_tokenizer = tok
                rows = [tok(t, truncation=True, max_length=max_len)["input_ids"] for t in ds["text"]]
            elif isinstance(ds, (list, tuple)):
                rows = [list(r) for r in ds]
            else:
                raise ValueError("unsupported calibration dataset type")
            rows = [r[:max_len] for r in rows if len(r) > 0]
            seq = min(len(r) for r in rows)
            ids = torch.tensor([r[:seq] for r in rows], dtype=torch.long)   # equal-length batch (no padding tokens)
        if n:
            ids = ids[: int(n)]
        if ids.shape[1] > max_len:
            ids = ids[:, :max_len]
        return ids.contiguous()

    def _oneshot(self, model, recipe, output_dir, save_compressed=True, **kw):
        from ...engine import artifacts, llama, pipeline, schemes
        if not torch.cuda.is_available():
            raise RuntimeError("the sm_100a quantization engine needs a CUDA device (there is no CPU fallback)")
        mods = recipe if isinstance(recipe, (list, tuple)) else [recipe]
        for m in mods:
            if not isinstance(m, Modifier):
                raise TypeError("recipe entries must be quantool_b200 Modifier objects (llm-compressor modifier "
                                "instances / YAML recipe paths are not interpretable without llm-compressor)")
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
                                               smooth_strength=smooth[0].smoothing_strength if smooth else None)
        else:
            res = pipeline.quantize_model_awq(shape, sd, ids, args, dev, fmt=fmt, n_grid=q.n_grid,
                                              duo_scaling=q.duo_scaling)
        qcfg = artifacts.quantization_config(q.scheme, args.actorder, fmt, ignore=q.ignore)
        qm = artifacts.QuantizedModel(cfg, res.tensors, qcfg, source_dir=src)
        qm.stats = res
        if save_compressed:
            qm.save_pretrained(output_dir, save_compressed=True)
        return qm
