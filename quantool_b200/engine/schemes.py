"""Quantization scheme resolution.

`quant_level` is a compressed-tensors preset scheme name, validated exactly as the reference
does with `compressed_tensors.is_preset_scheme` (ref/src/quantool/methods/llm_compressor/gptq/gptq.py:62;
table at CT/quantization/quant_scheme.py:406-428).  Only the scheme *description* comes from
compressed-tensors; all arithmetic runs in the CUDA kernels.
"""
from dataclasses import dataclass
from typing import Optional


@dataclass
class WeightArgs:
    num_bits: int
    symmetric: bool
    strategy: str                 # "group" | "channel"
    group_size: Optional[int]     # None for channel
    actorder: Optional[str] = None  # None | "group" | "weight"

    @property
    def qmin(self) -> int:
        return -(1 << (self.num_bits - 1))

    @property
    def qmax(self) -> int:
        return (1 << (self.num_bits - 1)) - 1


def is_preset_scheme(name: str) -> bool:
    from compressed_tensors.quantization import is_preset_scheme as _is
    return _is(name)


def resolve(level: str, actorder: Optional[str] = None) -> WeightArgs:
    from compressed_tensors.quantization import preset_name_to_scheme
    scheme = preset_name_to_scheme(level, ["Linear"])
    w = scheme.weights
    if w is None or str(getattr(w.type, "value", w.type)) != "int":
        raise ValueError(f"Scheme '{level}' is not an integer weight scheme; the sm_100a path implements "
                         f"W4A16, W4A16_ASYM, W8A16, W8A8/INT8 and W4A8")
    strategy = str(getattr(w.strategy, "value", w.strategy))
    if strategy not in ("group", "channel"):
        raise ValueError(f"weight strategy '{strategy}' of scheme '{level}' is not supported")
    ao = actorder if actorder is not None else (str(getattr(w.actorder, "value", w.actorder)) if w.actorder else None)
    if ao in ("static",):
        ao = "weight"
    if ao not in (None, "group", "weight"):
        raise ValueError(f"actorder must be None, 'group' or 'weight' (got {ao!r})")
    if strategy != "group":
        ao = None
    return WeightArgs(num_bits=w.num_bits, symmetric=bool(w.symmetric), strategy=strategy,
                      group_size=w.group_size if strategy == "group" else None, actorder=ao)


def scheme_config_groups(level: str, actorder: Optional[str]):
    """The `config_groups` block of config.json["quantization_config"] for this preset."""
    from compressed_tensors.quantization import preset_name_to_scheme
    scheme = preset_name_to_scheme(level, ["Linear"])
    d = scheme.model_dump(mode="json")
    if actorder is not None and d.get("weights") is not None:
        d["weights"]["actorder"] = "static" if actorder == "weight" else actorder
    return d
