"""BaseQuantizer: the plugin base class (mirrors ref/src/quantool/core/base.py:7-33)."""
from abc import ABC, abstractmethod
from typing import List, Union

from .logger import LoggerFactory
from .meta import TemplateQuantizationCard
from .mixins import CalibrationMixin, ExportMixin


class BaseQuantizer(ABC, ExportMixin, CalibrationMixin):
    name: str
    supported_levels: list
    supports_multiple_levels: bool = False
    template_card: TemplateQuantizationCard

    def __init__(self, model_id, *args, **kwargs):
        self.model_id = model_id
        self.logger = LoggerFactory.get_logger(self.__class__.__name__)
        super().__init__(*args, **kwargs)

    @abstractmethod
    def quantize(self, model, level: Union[str, List[str]], **kwargs) -> Union[str, List[str]]:
        if isinstance(level, list) and not self.supports_multiple_levels:
            raise ValueError(f"Method '{self.name}' does not support multiple quantization levels. "
                             f"Please specify a single level.")
