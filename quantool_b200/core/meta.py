"""Model-card template dataclass (mirrors ref/src/quantool/core/meta.py:5-21)."""
from dataclasses import dataclass, field
from typing import Any, Dict, List


@dataclass
class TemplateQuantizationCard:
    title: str
    description: str
    hyperparameters: Dict[str, Any] = field(default_factory=dict)
    intended_use: str = ""
    limitations: str = ""
    citations: List[str] = field(default_factory=list)
