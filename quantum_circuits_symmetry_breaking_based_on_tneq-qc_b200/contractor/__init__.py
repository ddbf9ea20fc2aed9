from .base import ContractionStrategy
from .compiler import StrategyCompiler
from .b200_strategy import B200Strategy

__all__ = ["ContractionStrategy", "StrategyCompiler", "B200Strategy"]
