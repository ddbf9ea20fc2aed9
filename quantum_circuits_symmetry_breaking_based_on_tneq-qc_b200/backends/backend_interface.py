"""Plug-in interface for compute backends.

Mirror of the reference ABC (tneq_qc/backends/backend_interface.py:14-518) so
that a backend written against the reference can be registered here and vice
versa: same method names, same argument meaning.  Only the interface lives
here; the one concrete implementation shipped is `B200Backend`.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Optional


class BackendInfo:
    """What a backend is: its registry name, device string and logical dtype."""

    def __init__(self, backend_type: str, device: Optional[str] = None, dtype: Optional[str] = None, **kwargs):
        self.backend_type = backend_type.lower()
        self.device = device
        self.dtype = dtype
        self.config = kwargs

    def __repr__(self):
        return (f"BackendInfo(backend_type='{self.backend_type}', device='{self.device}', "
                f"dtype='{self.dtype}', config={self.config})")


class ComputeBackend(ABC):
    """Tensor ops + autograd driver + optimizer update used by the engine."""

    def __init__(self, tensor_type: Optional[str] = None):
        self.backend_info: Optional[BackendInfo] = None
        self._tensor_type_name = tensor_type

    # -- TNTensor helpers (backend_interface.py:70-100) -----------------------
    @property
    def use_tn_tensor(self) -> bool:
        return self._tensor_type_name == "TNTensor"

    def wrap_tensor(self, tensor):
        if self.use_tn_tensor:
            from ..core.tn_tensor import TNTensor
            return tensor if isinstance(tensor, TNTensor) else TNTensor(tensor)
        return tensor

    def unwrap_tensor(self, tensor):
        from ..core.tn_tensor import TNTensor
        return tensor.tensor if isinstance(tensor, TNTensor) else tensor

    def get_tensor_type(self):
        if self.use_tn_tensor:
            from ..core.tn_tensor import TNTensor
            return TNTensor
        return self._get_raw_tensor_type()

    def get_backend_info(self) -> BackendInfo:
        if self.backend_info is None:
            self.backend_info = BackendInfo(self.get_backend_name())
        return self.backend_info

    def set_backend_info(self, backend_info: BackendInfo):
        if backend_info.backend_type != self.get_backend_name():
            raise ValueError(f"BackendInfo type '{backend_info.backend_type}' does not match "
                             f"backend '{self.get_backend_name()}'")
        self.backend_info = backend_info

    # -- what every backend must provide ----------------------------------------
    @abstractmethod
    def execute_expression(self, expression, *tensors): ...

    @abstractmethod
    def compute_value_and_grad(self, loss_fn, argnums): ...

    @abstractmethod
    def jit_compile(self, func): ...

    @abstractmethod
    def convert_to_tensor(self, array): ...

    @abstractmethod
    def optimizer_update(self, params, grads, state, method: str, hyperparams: dict): ...

    @abstractmethod
    def get_backend_name(self) -> str: ...

    @abstractmethod
    def init_random_core(self, shape): ...

    @abstractmethod
    def _get_raw_tensor_type(self): ...

    @abstractmethod
    def tensor_to_numpy(self, tensor): ...

    @abstractmethod
    def set_random_seed(self, seed: int): ...

    @abstractmethod
    def reshape(self, tensor, shape): ...

    @abstractmethod
    def eye(self, n: int, dtype=None): ...

    @abstractmethod
    def zeros(self, shape, dtype=None): ...

    @abstractmethod
    def ones(self, shape, dtype=None): ...

    @abstractmethod
    def clone(self, tensor): ...

    @abstractmethod
    def unsqueeze(self, tensor, dim): ...

    @abstractmethod
    def expand(self, tensor, *sizes): ...

    @abstractmethod
    def clamp(self, tensor, min=None, max=None): ...

    @abstractmethod
    def diagonal(self, tensor, dim1=-2, dim2=-1): ...

    @abstractmethod
    def sum(self, tensor, dim=None, keepdim=False): ...

    @abstractmethod
    def multinomial(self, probs, num_samples): ...

    @abstractmethod
    def arange(self, *args, dtype=None): ...

    @abstractmethod
    def stack(self, tensors, dim=0): ...

    @abstractmethod
    def log(self, tensor): ...

    @abstractmethod
    def mean(self, tensor, dim=None, keepdim=False): ...

    @abstractmethod
    def squeeze(self, tensor, dim=None): ...

    @abstractmethod
    def einsum(self, equation, *operands): ...
