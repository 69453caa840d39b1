"""quantool_b200 — B200-native implementation of quantool's per-layer quantization math
behind quantool's method-registry plugin API."""
from .core import BaseQuantizer, QuantizerRegistry, TemplateQuantizationCard

__all__ = ["BaseQuantizer", "QuantizerRegistry", "TemplateQuantizationCard"]
__version__ = "0.1.0"
