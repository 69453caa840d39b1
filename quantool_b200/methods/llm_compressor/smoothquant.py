"""SmoothQuant plugin (drop-in for ref/src/quantool/methods/llm_compressor/smoothquant/smoothquant.py:11-90):
a two-stage recipe, SmoothQuant smoothing followed by GPTQ."""
from typing import Any, Dict, List, Optional, Tuple

from ...core.meta import TemplateQuantizationCard
from ...core.registry import QuantizerRegistry
from .base import LLMCompressorQuantizer, Modifier, RecipeType
from .gptq import _PRESETS


@QuantizerRegistry.register
class SmoothQuant(LLMCompressorQuantizer):
    name = "smoothquant"
    supported_levels = ["W8A8", "INT8", "W4A8"]
    template_card = TemplateQuantizationCard(
        title="SmoothQuant",
        description="Smoothing-based activation quantization for W8A8",
        hyperparameters={"method": "smoothquant", "scheme": "W8A8", "smoothing_strength": 0.5, "targets": "Linear",
                         "ignore": ["lm_head"], "num_calibration_samples": 512},
        intended_use="W8A8 quantization with activation smoothing for better accuracy",
        limitations="Requires calibration dataset; best for W8A8 schemes",
        citations=["https://arxiv.org/abs/2211.10438"],
    )

    def _build_recipe(self, level: Optional[str], method_kwargs: Dict[str, Any]) -> Tuple[RecipeType, str]:
        from ...engine.schemes import is_preset_scheme
        scheme = level or method_kwargs.get("scheme", "W8A8")
        if not is_preset_scheme(scheme):
            raise ValueError(f"Scheme '{scheme}' is not a valid compressed-tensors preset scheme. {_PRESETS}")
        smoothing_strength = method_kwargs.get("smoothing_strength", 0.5)
        recipe: List[Any] = [
            Modifier(kind="smoothquant", smoothing_strength=smoothing_strength),
            Modifier(kind="gptq", scheme=scheme, targets=method_kwargs.get("targets", "Linear"),
                     ignore=method_kwargs.get("ignore", ["lm_head"])),
        ]
        self.logger.info(f"Built SmoothQuant recipe with scheme={scheme}, smoothing_strength={smoothing_strength}")
        return recipe, scheme
