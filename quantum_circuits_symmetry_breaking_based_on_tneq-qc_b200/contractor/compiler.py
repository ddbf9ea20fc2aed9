"""Strategy registry and selection (reference: tneq_qc/contractor/compiler.py:14-136).

`compile` asks every strategy registered for the current mode whether it is
compatible, and returns the compute function of the cheapest one.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Tuple

from .base import ContractionStrategy


class StrategyCompiler:
    MODES: Dict[str, List[str]] = {"fast": [], "balanced": [], "full": []}
    _strategies: Dict[str, ContractionStrategy] = {}
    verbose = False

    def __init__(self, mode: str = "fast"):
        if mode not in self.MODES:
            raise ValueError(f"Invalid mode '{mode}'. Must be one of {list(self.MODES.keys())}")
        self.mode = mode

    @classmethod
    def register_strategy(cls, strategy: ContractionStrategy, modes: List[str] = None):
        cls._strategies[strategy.name] = strategy
        for mode in modes or ():
            if mode in cls.MODES and strategy.name not in cls.MODES[mode]:
                cls.MODES[mode].append(strategy.name)

    def register_custom_strategy(self, strategy: ContractionStrategy, modes: List[str]):
        self.register_strategy(strategy, modes)

    @classmethod
    def get_registered_strategies(cls) -> Dict[str, ContractionStrategy]:
        return cls._strategies.copy()

    @property
    def strategies(self) -> Dict[str, ContractionStrategy]:
        return self._strategies

    def compile(self, qctn, shapes_info: Dict[str, Any], backend, **kwargs) -> Tuple[Callable, str, float]:
        best = None
        for name in self.MODES[self.mode]:
            strategy = self._strategies.get(name)
            if strategy is None or not strategy.check_compatibility(qctn, shapes_info):
                continue
            cost = strategy.estimate_cost(qctn, shapes_info)
            if best is None or cost < best[2]:
                best = (strategy, name, cost)
        if best is None:
            raise RuntimeError("No compatible strategy found!")
        if self.verbose:
            print(f"[Compiler] Selected strategy: {best[1]} (cost: {best[2]:.2e})")
        return best[0].get_compute_function(qctn, shapes_info, backend, **kwargs), best[1], best[2]
