from .base import BaseQuantizer
from .logger import LoggerFactory
from .meta import TemplateQuantizationCard
from .mixins import CalibrationMixin, ExportMixin
from .registry import QuantizerRegistry, Registry

__all__ = ["BaseQuantizer", "LoggerFactory", "TemplateQuantizationCard", "CalibrationMixin",
           "ExportMixin", "QuantizerRegistry", "Registry"]
