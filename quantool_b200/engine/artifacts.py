"""compressed-tensors artifact writer: model.safetensors + config.json["quantization_config"].

What `last_model.save_pretrained(dest, save_compressed=True)` produces for the reference
(ref/src/quantool/methods/llm_compressor/base.py:188; CT/compressors/model_compressors/model_compressor.py:177-213;
SURVEY.md §8b "Artifact contract", §8f row 1).  The metadata block is generated with the installed
compressed-tensors pydantic models so that key names / defaults are the real ones; tensors come
from the CUDA kernels.  Also offers the AutoGPTQ / AutoAWQ `qweight/qzeros/scales/g_idx` view as a
pure repack of the same codes (north_star naming; layouts per
VLLM/model_executor/layers/quantization/utils/quant_utils.py:704-813).
"""
import json
import os
import shutil
from typing import Dict, Optional

import torch


def quantization_config(level: str, actorder: Optional[str], fmt: str, ignore=("lm_head",)) -> dict:
    """The `quantization_config` block as the installed compressed-tensors writes it
    (`ModelCompressor.update_config`, CT/compressors/model_compressors/model_compressor.py:177-213): the pydantic dump
    of the config - the group carries its own `format` - plus `version`, `quant_method` and the (empty)
    `sparsity_config` / `transform_config` entries of this CT version.
    tests/test_artifact_consumers.py compares it with that writer's own output."""
    from compressed_tensors import __version__ as ct_version
    from compressed_tensors.quantization import QuantizationConfig, preset_name_to_scheme
    scheme = preset_name_to_scheme(level, ["Linear"])
    if actorder is not None and scheme.weights is not None:
        scheme.weights.actorder = "static" if actorder == "weight" else actorder
    if "format" in type(scheme).model_fields:
        scheme.format = fmt
    qc = QuantizationConfig(config_groups={"group_0": scheme}, quant_method="compressed-tensors", format=fmt,
                            quantization_status="compressed", ignore=list(ignore))
    d = qc.model_dump(mode="json")
    d["version"] = ct_version
    try:
        from compressed_tensors.base import SPARSITY_CONFIG_NAME, TRANSFORM_CONFIG_NAME
        d.setdefault(SPARSITY_CONFIG_NAME, {})
        d.setdefault(TRANSFORM_CONFIG_NAME, {})
    except ImportError:       # older compressed-tensors: no such entries
        pass
    return d


def artifact_format(num_bits: int, level: str) -> str:
    """The per-module format compressed-tensors infers (CT/compressors/format.py infer_module_format ->
    PackedQuantizationCompressor.can_compress / IntQuantizationCompressor.can_compress), restated for integer
    Linear schemes: weight-only INT with 4 or 8 bits -> "pack-quantized" (W4A16, W4A16_ASYM, W8A16: `weight_packed`
    int32); INT weights WITH quantized input activations -> "int-quantized" (W8A8/INT8, W4A8: int8 `weight`).
    tests/test_plugin_api.py checks this table against the installed compressed-tensors."""
    from compressed_tensors.quantization import preset_name_to_scheme
    scheme = preset_name_to_scheme(level, ["Linear"])
    if scheme.input_activations is None and num_bits in (4, 8):
        return "pack-quantized"
    if scheme.input_activations is not None:
        return "int-quantized"
    raise ValueError(f"no compressed-tensors integer format for scheme {level!r} ({num_bits}-bit weight-only)")


class QuantizedModel:
    """Holder returned as `plugin.last_model`: host tensors in artifact key names + HF config."""

    def __init__(self, hf_config: dict, tensors: Dict[str, torch.Tensor], qconfig: dict, source_dir: Optional[str] = None,
                 writer: bool = True):
        self.hf_config = dict(hf_config)
        self.tensors = tensors
        self.qconfig = qconfig
        self.source_dir = source_dir
        self.writer = writer        # multi-process runs: rank 0 holds the tensors and is the only one that writes

    def save_pretrained(self, dest: str, save_compressed: bool = True, max_shard_size="5GB", **_):
        """config.json + weights laid out as `PreTrainedModel.save_pretrained` does for the reference
        (transformers 4.56.2 as pinned at ref/pyproject.toml:26-30: `max_shard_size="5GB"`): one `model.safetensors`
        when everything fits into a shard, else `model-0000i-of-0000n.safetensors` + `model.safetensors.index.json`,
        split by huggingface_hub's `split_torch_state_dict_into_shards` (the function transformers itself calls)."""
        from huggingface_hub import split_torch_state_dict_into_shards
        from safetensors.torch import save_file
        if not self.writer:
            return
        os.makedirs(dest, exist_ok=True)
        cfg = dict(self.hf_config)
        cfg["quantization_config"] = self.qconfig
        with open(os.path.join(dest, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)            # transformers and CT both write sorted keys
        tensors = {k: v.contiguous() for k, v in self.tensors.items()}
        split = split_torch_state_dict_into_shards(tensors, filename_pattern="model{suffix}.safetensors",
                                                   max_shard_size=max_shard_size)
        for fn, keys in split.filename_to_tensors.items():
            save_file({k: tensors[k] for k in keys}, os.path.join(dest, fn), metadata={"format": "pt"})
        if split.is_sharded:
            with open(os.path.join(dest, "model.safetensors.index.json"), "w") as f:
                json.dump({"metadata": split.metadata, "weight_map": split.tensor_to_filename}, f, indent=2, sort_keys=True)
        if getattr(self, "recipe", None):
            with open(os.path.join(dest, "recipe.yaml"), "w") as f:
                f.write(recipe_yaml(self.recipe))
        if self.source_dir and os.path.isdir(self.source_dir):
            for fn in os.listdir(self.source_dir):
                if fn.startswith("tokenizer") or fn in ("special_tokens_map.json", "generation_config.json", "vocab.json",
                                                        "merges.txt"):
                    shutil.copy(os.path.join(self.source_dir, fn), dest)


_MODIFIER_NAMES = {"gptq": "GPTQModifier", "awq": "AWQModifier", "smoothquant": "SmoothQuantModifier"}
_MODIFIER_FIELDS = {"gptq": ("targets", "ignore", "scheme", "block_size", "dampening_frac", "actorder"),
                    "awq": ("targets", "ignore", "scheme", "duo_scaling"),
                    "smoothquant": ("smoothing_strength",)}


def recipe_yaml(modifiers) -> str:
    """`recipe.yaml`, the record of the applied recipe llm-compressor's `oneshot` leaves next to the weights
    (`default_stage: default_modifiers: <ModifierClass>: {fields}`; layout restated from published llm-compressor
    output directories - informational, no consumer parses it, unpinned)."""
    import yaml
    mods = {}
    for m in modifiers:
        fields = {}
        for k in _MODIFIER_FIELDS[m.kind]:
            v = getattr(m, k)
            if v is None:
                continue
            fields[k] = [v] if k == "targets" and isinstance(v, str) else (list(v) if isinstance(v, (list, tuple)) else v)
        if m.kind == "awq" and m.n_grid != 20:
            fields["n_grid"] = m.n_grid                 # additive key; upstream's grid is the constant 20
        mods[_MODIFIER_NAMES[m.kind]] = fields
    return yaml.safe_dump({"default_stage": {"default_modifiers": mods}}, sort_keys=False)


def autogptq_view(weight_packed: torch.Tensor, weight_scale: torch.Tensor, weight_zero_point: Optional[torch.Tensor],
                  g_idx: Optional[torch.Tensor], num_bits: int, K: int, group_size: int) -> Dict[str, torch.Tensor]:
    """AutoGPTQ layout: qweight [K/pf, N] packed along K, scales [G, N], qzeros [G, N/pf], g_idx [K].
    Pure index shuffling of the compressed-tensors codes (device or host tensors)."""
    from .. import cabi
    N = weight_packed.shape[0]
    pf = 32 // num_bits
    dev = weight_packed.device
    if dev.type == "cuda":
        codes = cabi.unpack_int32(weight_packed.contiguous(), num_bits, K)
    else:
        shifts = torch.arange(pf, dtype=torch.int32) * num_bits
        u = (weight_packed.unsqueeze(-1) >> shifts) & ((1 << num_bits) - 1)
        codes = (u.reshape(N, -1)[:, :K] - (1 << (num_bits - 1))).to(torch.int8)
    u = (codes.to(torch.int32) + (1 << (num_bits - 1))).t().contiguous()          # [K, N] unsigned
    shifts = (torch.arange(pf, dtype=torch.int32, device=u.device) * num_bits).view(1, pf, 1)
    qweight = (u.view(K // pf, pf, N) << shifts).sum(dim=1, dtype=torch.int32)
    G = weight_scale.shape[1]
    scales = weight_scale.t().contiguous()
    if weight_zero_point is None:
        zeros = torch.full((G, N), 1 << (num_bits - 1), dtype=torch.int32, device=u.device)
    else:
        zeros = (weight_zero_point.to(torch.int32) + (1 << (num_bits - 1))).t().contiguous()
    shifts2 = (torch.arange(pf, dtype=torch.int32, device=u.device) * num_bits).view(1, 1, pf)
    qzeros = (zeros.view(G, N // pf, pf) << shifts2).sum(dim=2, dtype=torch.int32)
    if g_idx is None:
        g_idx = (torch.arange(K, dtype=torch.int32, device=u.device) // (group_size or K))
    return {"qweight": qweight, "qzeros": qzeros, "scales": scales, "g_idx": g_idx.to(torch.int32)}


AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)   # AutoAWQ nibble order inside one int32 (vLLM reverses it with 0,4,1,5,2,6,3,7)


def autoawq_view(weight_packed: torch.Tensor, weight_scale: torch.Tensor, weight_zero_point: Optional[torch.Tensor],
                 K: int) -> Dict[str, torch.Tensor]:
    """AutoAWQ GEMM layout of a 4-bit group-quantized Linear: qweight [K, N/8] packed along N (output channels)
    with the interleave AWQ_ORDER, qzeros [G, N/8] packed the same way, scales [G, N].  Pure index shuffling of
    the compressed-tensors codes (consumer: vLLM quant_utils awq_pack)."""
    from .. import cabi
    N = weight_packed.shape[0]
    dev = weight_packed.device
    if N % 8:
        raise ValueError("AutoAWQ packing needs the number of output channels to be a multiple of 8")
    if dev.type == "cuda":
        codes = cabi.unpack_int32(weight_packed.contiguous(), 4, K)
    else:
        shifts = torch.arange(8, dtype=torch.int32) * 4
        u = (weight_packed.unsqueeze(-1) >> shifts) & 0xF
        codes = (u.reshape(N, -1)[:, :K] - 8).to(torch.int8)
    u = (codes.to(torch.int32) + 8).t().contiguous()                       # [K, N] unsigned codes
    order = torch.tensor(AWQ_ORDER, dtype=torch.long, device=u.device)
    shifts = (torch.arange(8, dtype=torch.int32, device=u.device) * 4).view(1, 1, 8)
    qweight = (u.view(K, N // 8, 8)[:, :, order] << shifts).sum(dim=2, dtype=torch.int32)
    G = weight_scale.shape[1]
    if weight_zero_point is None:
        zeros = torch.full((G, N), 8, dtype=torch.int32, device=u.device)
    else:
        zeros = (weight_zero_point.to(torch.int32) + 8).t().contiguous()
    qzeros = (zeros.view(G, N // 8, 8)[:, :, order] << shifts).sum(dim=2, dtype=torch.int32)
    return {"qweight": qweight, "qzeros": qzeros, "scales": weight_scale.t().contiguous()}
