"""AWQ plugin (drop-in for ref/src/quantool/methods/llm_compressor/awq/awq.py:11-84)."""
from typing import Any, Dict, Optional, Tuple

from ...core.meta import TemplateQuantizationCard
from ...core.registry import QuantizerRegistry
from .base import LLMCompressorQuantizer, Modifier, RecipeType
from .gptq import _PRESETS


@QuantizerRegistry.register
class AWQ(LLMCompressorQuantizer):
    name = "awq"
    supported_levels = ["W4A16", "W4A16_ASYM", "W8A16"]
    template_card = TemplateQuantizationCard(
        title="AWQ Quantization",
        description="Activation-aware weight quantization preserving salient weights",
        hyperparameters={"method": "awq", "scheme": "W4A16", "targets": "Linear", "ignore": ["lm_head"],
                         "num_calibration_samples": 512},
        intended_use="Weight-only quantization with better accuracy than naive PTQ",
        limitations="Requires calibration dataset; weight-only (activations remain fp16)",
        citations=["https://arxiv.org/abs/2306.00978"],
    )

    def _build_recipe(self, level: Optional[str], method_kwargs: Dict[str, Any]) -> Tuple[RecipeType, str]:
        from ...engine.schemes import is_preset_scheme
        scheme = level or method_kwargs.get("scheme", "W4A16")
        if not is_preset_scheme(scheme):
            raise ValueError(f"Scheme '{scheme}' is not a valid compressed-tensors preset scheme. {_PRESETS}")
        if scheme not in self.supported_levels:
            self.logger.warning(f"AWQ only supports weight-only quantization with 16-bit activations. "
                                f"Scheme '{scheme}' may not be compatible. Supported: {self.supported_levels}")
        kw = {"scheme": scheme, "targets": method_kwargs.get("targets", "Linear"),
              "ignore": method_kwargs.get("ignore", ["lm_head"])}
        for key in ["mappings", "smoothing_strength"]:      # ref awq.py:77-79
            if key in method_kwargs:
                kw[key] = method_kwargs[key]
        for key in ["n_grid", "duo_scaling"]:               # upstream constants, exposed as additive keys
            if key in method_kwargs:
                kw[key] = method_kwargs[key]
        if kw.get("mappings") is not None:
            raise ValueError("custom AWQ mappings are not supported: the engine uses the Llama default mappings")
        recipe = Modifier(kind="awq", **kw)
        self.logger.info(f"Built AWQ recipe with scheme={scheme}")
        return recipe, scheme
