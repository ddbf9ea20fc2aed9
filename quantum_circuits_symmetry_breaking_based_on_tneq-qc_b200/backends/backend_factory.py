"""name -> backend class registry (reference: tneq_qc/backends/backend_factory.py:17-100).

Only one backend ships: 'b200' (hand-written sm_100a CUDA behind a C ABI).
There is deliberately no CPU / multi-vendor fallback: creating the backend
without a CUDA device or without the built library raises.
"""
from __future__ import annotations

from typing import Optional, Type

from .backend_interface import ComputeBackend


class BackendFactory:
    _backends = {}
    _default_backend: Optional[str] = None
    _backend_instance: Optional[ComputeBackend] = None

    @classmethod
    def create_backend(cls, backend_name: str, device: Optional[str] = None,
                       tensor_type: Optional[str] = None, **kwargs) -> ComputeBackend:
        key = backend_name.lower()
        if key not in cls._backends:
            raise ValueError(f"Unknown backend: {key}. Available backends: {list(cls._backends.keys())}")
        return cls._backends[key](device=device, tensor_type=tensor_type, **kwargs)

    @classmethod
    def set_default_backend(cls, backend_name: str, device: Optional[str] = None,
                            tensor_type: Optional[str] = None, **kwargs):
        cls._default_backend = backend_name.lower()
        cls._backend_instance = cls.create_backend(backend_name, device=device, tensor_type=tensor_type, **kwargs)

    @classmethod
    def get_default_backend(cls) -> ComputeBackend:
        if cls._backend_instance is None:
            cls.set_default_backend("b200", "cuda")
        return cls._backend_instance

    @classmethod
    def register_backend(cls, name: str, backend_class: Type[ComputeBackend]):
        cls._backends[name.lower()] = backend_class
