"""Contraction-strategy plug-in interface (reference: tneq_qc/contractor/base.py:12-62)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Callable, Dict


class ContractionStrategy(ABC):
    """A way of contracting one QCTN against circuit states and measurements.

    get_compute_function returns
        compute_fn(cores_dict, circuit_states, measure_matrices, right_cores_dict=None)
    """

    @abstractmethod
    def check_compatibility(self, qctn, shapes_info: Dict[str, Any]) -> bool: ...

    @abstractmethod
    def get_compute_function(self, qctn, shapes_info: Dict[str, Any], backend, **kwargs) -> Callable: ...

    @abstractmethod
    def estimate_cost(self, qctn, shapes_info: Dict[str, Any]) -> float: ...

    @property
    @abstractmethod
    def name(self) -> str: ...
