"""GPTQ plugin (drop-in for ref/src/quantool/methods/llm_compressor/gptq/gptq.py:12-91)."""
from typing import Any, Dict, Optional, Tuple

from ...core.meta import TemplateQuantizationCard
from ...core.registry import QuantizerRegistry
from .base import LLMCompressorQuantizer, Modifier, RecipeType

_PRESETS = ("Valid schemes include: W8A16, W4A16, W4A16_ASYM, W8A8, INT8, W4A8, "
            "FP8, FP8_DYNAMIC, FP8_BLOCK, NVFP4A16, NVFP4, UNQUANTIZED")


@QuantizerRegistry.register
class GPTQ(LLMCompressorQuantizer):
    name = "gptq"
    supported_levels = ["W4A16", "W8A8", "INT8", "W8A16", "W4A16_ASYM", "W4A8"]
    template_card = TemplateQuantizationCard(
        title="GPTQ Quantization",
        description="Post-training quantization using GPTQ algorithm with calibration data",
        hyperparameters={"method": "gptq", "scheme": "W4A16", "targets": "Linear", "ignore": ["lm_head"],
                         "num_calibration_samples": 512},
        intended_use="Efficient inference for LLMs with minimal accuracy loss",
        limitations="Requires calibration dataset; quantization time scales with model size",
        citations=["https://arxiv.org/abs/2210.17323"],
    )

    def _build_recipe(self, level: Optional[str], method_kwargs: Dict[str, Any]) -> Tuple[RecipeType, str]:
        from ...engine.schemes import is_preset_scheme
        scheme = level or method_kwargs.get("scheme", "W4A16")
        if not is_preset_scheme(scheme):
            raise ValueError(f"Scheme '{scheme}' is not a valid compressed-tensors preset scheme. {_PRESETS}")
        if scheme not in self.supported_levels:
            self.logger.warning(f"Level '{scheme}' not in supported list, using anyway: {self.supported_levels}")
        kw = {"scheme": scheme, "targets": method_kwargs.get("targets", "Linear"),
              "ignore": method_kwargs.get("ignore", ["lm_head"])}
        for key in ["block_size", "dampening_frac", "sequential_targets"]:   # ref gptq.py:82-84
            if key in method_kwargs:
                kw[key] = method_kwargs[key]
        if "actorder" in method_kwargs:      # additive key: the reference cannot reach actorder
            kw["actorder"] = method_kwargs["actorder"]
        recipe = Modifier(kind="gptq", **kw)
        self.logger.info(f"Built GPTQ recipe with scheme={scheme}, targets={kw['targets']}")
        return recipe, scheme
