from .backend_interface import BackendInfo, ComputeBackend
from .backend_factory import BackendFactory
from .backend_b200 import B200Backend

__all__ = ["BackendInfo", "ComputeBackend", "BackendFactory", "B200Backend"]
